// Fused tri-plane ray-march, backward (sm_100a).
//
// Nothing is saved by the forward: each CTA re-runs passes A-C for its rays (raymarch_block.cuh), then
//   D'. per ray: forward transmittance scan over the depth-merged 2N samples, reverse scan producing
//       dL/d(rgb_i, sigma_i) for every sample ("Backward"),
//   E.  per sample (lane == sample): MLP backward (dW1,db1,dW2,db2 reduced in shared memory, flushed once per CTA),
//       d(feature) -> scatter-add into the plane gradients with red.global.add.v4.f32 (4 taps x 3 planes x 32 ch),
//       and d(sample position) -> d(ray_o), d(ray_d) through the bilinear weights.
// The importance-sampled depths are constants (tri_plane_renderer.py:241,254), so are the coarse depths.
// Persistent grid (multiple of the SM count), grid-stride over ray blocks.
#include "raymarch_block.cuh"

namespace rm {

constexpr int kTRB = 8;            // rays per block iteration in the backward
constexpr int kStageStride = 68;   // sample-major staging rows of 64 floats, 16-byte aligned, conflict-free

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct BwdSmem {
    float4* g_co; float4* g_fi;   // [TR][N+1]
    float* Tbuf;                  // [TR][2N+1]
    float* stage;                 // [kWarps][32][kStageStride]   (v / g_pre staging; aliased by gdot [32][12])
    float* gro; float* grd;       // [TR][3]
    float* gw1; float* gb1; float* gw2; float* gb2;   // [kC][kH], [kH], [kH][4], [4]
    unsigned char* ord;           // [TR][2N]
    static __host__ __device__ size_t bytes(int N) {
        size_t b = 2 * (size_t)kTRB * (N + 1) * 16;
        b += (size_t)kTRB * (2 * N + 1) * 4;
        b += (size_t)kWarps * 32 * kStageStride * 4;
        b += (size_t)kTRB * 6 * 4;
        b += (size_t)(kC * kH + kH + kH * 4 + 4) * 4;
        b += (size_t)kTRB * 2 * N;
        return (b + 15) & ~(size_t)15;
    }
    __device__ void carve(unsigned char* raw, int N) {
        g_co = reinterpret_cast<float4*>(raw);
        g_fi = g_co + kTRB * (N + 1);
        Tbuf = reinterpret_cast<float*>(g_fi + kTRB * (N + 1));
        stage = Tbuf + kTRB * (2 * N + 1);
        // keep `stage` 16-byte aligned
        stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(stage) + 15) & ~(uintptr_t)15);
        gro = stage + kWarps * 32 * kStageStride;
        grd = gro + kTRB * 3;
        gw1 = grd + kTRB * 3;
        gb1 = gw1 + kC * kH;
        gw2 = gb1 + kH;
        gb2 = gw2 + kH * 4;
        ord = reinterpret_cast<unsigned char*>(gb2 + 4);
    }
};

template <class PT>
__global__ void __launch_bounds__(kThreads, 2) raymarch_bwd_kernel(Params p, int nblocks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TR = kTRB;
    Block<TR> s;
    s.carve(smem_raw, p.o.N);
    BwdSmem g;
    g.carve(smem_raw + ((Block<TR>::bytes(p.o.N) + 15) & ~(size_t)15), p.o.N);
    const int N = p.o.N, NP = s.NP, R = p.o.R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blocks_per_img = (R + TR - 1) / TR;
    float* featw = s.feat + warp * kC * 33;
    float* stg = g.stage + warp * 32 * kStageStride;

    stage_mlp<TR>(s, p);
    for (int t = tid; t < kC * kH + kH + kH * 4 + 4; t += kThreads) g.gw1[t] = 0.f;   // gw1,gb1,gw2,gb2 are contiguous
    __syncthreads();

    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int b = blk / blocks_per_img;
        const int r0 = (blk - b * blocks_per_img) * TR;
        const int nrays = min(TR, R - r0);
        const int64_t ray_base = (int64_t)b * R + r0;
        const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;
        float* gimg = p.g_planes + (int64_t)b * p.psB;

        for (int t = tid; t < nrays * 3; t += kThreads) { s.ro[t] = p.ray_o[ray_base * 3 + t]; s.rd[t] = p.ray_d[ray_base * 3 + t]; }
        for (int t = tid; t < TR * 6; t += kThreads) g.gro[t] = 0.f;   // gro, grd contiguous
        __syncthreads();

        forward_passes<PT, TR>(s, p, img, ray_base, nrays);

        // ---- D': per-ray compositing forward (transmittances, merge order) + reverse scan
        if (tid < nrays) {
            const int rl = tid;
            float* Tb = g.Tbuf + rl * (2 * N + 1);
            unsigned char* od = g.ord + rl * 2 * N;
            const float big = p.o.use_inf_depth ? 1e10f : 1e-3f;
            const float t0 = p.o.ray_start, t1 = p.o.ray_end;
            const float gr = p.g_rgb[(ray_base + rl) * 3 + 0], gg = p.g_rgb[(ray_base + rl) * 3 + 1], gb = p.g_rgb[(ray_base + rl) * 3 + 2];
            const float gd = p.g_depth[ray_base + rl];
            {
                Merge<TR> mg(s, p, rl);
                float T = 1.f, tcur, tnext = 0.f; float4 cur, nxt;
                int code = mg.pop(tcur, cur), ncode = code;
                nxt = cur;
                for (int m = 0; m < 2 * N; m++) {
                    const bool last = (m == 2 * N - 1);
                    if (!last) ncode = mg.pop(tnext, nxt);
                    const float delta = last ? big : (tnext - tcur);
                    const float alpha = 1.f - expf(-delta * density_act(cur.w, p.o.clamp_mode));
                    Tb[m] = T; od[m] = (unsigned char)code;
                    T *= (1.f - alpha + 1e-10f);
                    if (!last) { cur = nxt; tcur = tnext; code = ncode; }
                }
            }
            // value of the last sample and of the background terms
            auto fetch = [&](int code, float& t, float4& v) {
                if (code < N) { t = s_to_t(s.s_co[rl * NP + code], t0, t1); v = s.out_co[rl * (N + 1) + code]; }
                else          { t = s_to_t(s.s_fi[rl * NP + code - N], t0, t1); v = s.out_fi[rl * (N + 1) + code - N]; }
            };
            float tl; float4 vl;
            fetch(od[2 * N - 1], tl, vl);
            float K = 0.f;
            if (p.o.last_back) K += gr * vl.x + gg * vl.y + gb * vl.z + gd * tl;
            if (p.o.white_back_end_idx > 0) K += gr + (p.o.white_back_end_idx > 1 ? gg : 0.f) + (p.o.white_back_end_idx > 2 ? gb : 0.f);
            // reverse scan
            float S = 0.f, wagg = 0.f, tnext = 0.f;
            float4 glast = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int m = 2 * N - 1; m >= 0; m--) {
                const int code = od[m];
                float tm; float4 vm;
                fetch(code, tm, vm);
                const float delta = (m == 2 * N - 1) ? big : (tnext - tm);
                const float sp = density_act(vm.w, p.o.clamp_mode);
                const float e = expf(-delta * sp);
                const float alpha = 1.f - e;
                const float Tm = Tb[m];
                const float w = alpha * Tm;
                const float Gm = gr * vm.x + gg * vm.y + gb * vm.z + gd * tm - K;
                const float dalpha = Tm * Gm - S / (1.f - alpha + 1e-10f);
                const float dsig = dalpha * (delta * e) * density_act_grad(vm.w, p.o.clamp_mode);
                float4 gv = make_float4(w * gr, w * gg, w * gb, dsig);
                S += w * Gm;
                wagg += w;
                if (m == 2 * N - 1) glast = gv;
                else { if (code < N) g.g_co[rl * (N + 1) + code] = gv; else g.g_fi[rl * (N + 1) + code - N] = gv; }
                tnext = tm;
            }
            if (p.o.last_back) { const float extra = 1.f - wagg; glast.x += extra * gr; glast.y += extra * gg; glast.z += extra * gb; }
            { const int code = od[2 * N - 1]; if (code < N) g.g_co[rl * (N + 1) + code] = glast; else g.g_fi[rl * (N + 1) + code - N] = glast; }
        }
        __syncthreads();

        // ---- E: per-sample backward through MLP, plane interpolation and sample position
        const int total = TR * N;
        const float pixscale = 0.5f * (float)(p.o.P - 1) / p.o.box_half;   // d(pixel)/d(world coordinate)
        for (int pass = 0; pass < 2; pass++) {
            const float4* gsrc = pass ? g.g_fi : g.g_co;
            for (int s0 = 0; s0 < total; s0 += kThreads) {
                const int si = s0 + tid;
                const int rl = si / N, i = si - rl * N;
                const bool valid = (si < total) && (rl < nrays);
                float sd = 0.f;
                if (valid) sd = pass ? s.s_fi[rl * NP + i] : s.s_co[rl * NP + i];
                const float tval = s_to_t(sd, p.o.ray_start, p.o.ray_end);
                Footprint fp;
                footprint_of<TR>(fp, s, p, rl, sd, valid);
                gather_features<PT>(img, fp, featw, p.psX, p.psY, lane);
                float h[kH];
                mlp_hidden(h, featw, s.w1s, s.b1s, lane);
                float4 go = valid ? gsrc[rl * (N + 1) + i] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float sqrt2 = 1.4142135623730951f;

                // (a) stage v = act(h) sample-major; dW2[j][k] += sum_s v[s][j] go[s][k]; db2[k] += sum_s go[s][k]
#pragma unroll
                for (int j4 = 0; j4 < kH / 4; j4++) {
                    float4 v;
                    v.x = (h[4 * j4 + 0] > 0.f ? h[4 * j4 + 0] : 0.2f * h[4 * j4 + 0]) * sqrt2;
                    v.y = (h[4 * j4 + 1] > 0.f ? h[4 * j4 + 1] : 0.2f * h[4 * j4 + 1]) * sqrt2;
                    v.z = (h[4 * j4 + 2] > 0.f ? h[4 * j4 + 2] : 0.2f * h[4 * j4 + 2]) * sqrt2;
                    v.w = (h[4 * j4 + 3] > 0.f ? h[4 * j4 + 3] : 0.2f * h[4 * j4 + 3]) * sqrt2;
                    reinterpret_cast<float4*>(stg + lane * kStageStride)[j4] = v;
                }
                __syncwarp();
                {
                    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int ss = 0; ss < 32; ss++) {
                        const float gx = __shfl_sync(0xffffffffu, go.x, ss), gy = __shfl_sync(0xffffffffu, go.y, ss);
                        const float gz = __shfl_sync(0xffffffffu, go.z, ss), gw = __shfl_sync(0xffffffffu, go.w, ss);
                        const float v0 = stg[ss * kStageStride + lane], v1 = stg[ss * kStageStride + lane + 32];
                        a0[0] = fmaf(v0, gx, a0[0]); a0[1] = fmaf(v0, gy, a0[1]); a0[2] = fmaf(v0, gz, a0[2]); a0[3] = fmaf(v0, gw, a0[3]);
                        a1[0] = fmaf(v1, gx, a1[0]); a1[1] = fmaf(v1, gy, a1[1]); a1[2] = fmaf(v1, gz, a1[2]); a1[3] = fmaf(v1, gw, a1[3]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) { atomicAdd(&g.gw2[lane * 4 + k], a0[k]); atomicAdd(&g.gw2[(lane + 32) * 4 + k], a1[k]); }
                    float sx = go.x, sy = go.y, sz = go.z, sw = go.w;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
                        sz += __shfl_xor_sync(0xffffffffu, sz, o); sw += __shfl_xor_sync(0xffffffffu, sw, o);
                    }
                    if (lane == 0) { atomicAdd(&g.gb2[0], sx); atomicAdd(&g.gb2[1], sy); atomicAdd(&g.gb2[2], sz); atomicAdd(&g.gb2[3], sw); }
                }
                __syncwarp();

                // (b) g_pre (in place of h) and g_f
#pragma unroll
                for (int j = 0; j < kH; j++) {
                    const float4 w = reinterpret_cast<const float4*>(s.w2s)[j];
                    const float gh = go.x * w.x + go.y * w.y + go.z * w.z + go.w * w.w;
                    h[j] = gh * (h[j] > 0.f ? 1.f : 0.2f) * sqrt2;
                }
                float gf[kC];
#pragma unroll 2
                for (int c = 0; c < kC; c++) {
                    const float4* wr = reinterpret_cast<const float4*>(s.w1s + c * kH);
                    float a = 0.f;
#pragma unroll
                    for (int j4 = 0; j4 < kH / 4; j4++) {
                        const float4 w = wr[j4];
                        a = fmaf(w.x, h[4 * j4 + 0], a); a = fmaf(w.y, h[4 * j4 + 1], a);
                        a = fmaf(w.z, h[4 * j4 + 2], a); a = fmaf(w.w, h[4 * j4 + 3], a);
                    }
                    gf[c] = a;
                }
                // (c) stage g_pre sample-major; dW1[c][j] += sum_s f[c][s] g_pre[s][j]; db1[j] += sum_s g_pre[s][j]
#pragma unroll
                for (int j4 = 0; j4 < kH / 4; j4++)
                    reinterpret_cast<float4*>(stg + lane * kStageStride)[j4] = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
                __syncwarp();
                {
                    float fc[32];
#pragma unroll
                    for (int ss = 0; ss < 32; ss++) fc[ss] = featw[lane * 33 + ss];
                    for (int j4 = 0; j4 < kH / 4; j4++) {
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int ss = 0; ss < 32; ss++) {
                            const float4 gp = reinterpret_cast<const float4*>(stg + ss * kStageStride)[j4];
                            a.x = fmaf(fc[ss], gp.x, a.x); a.y = fmaf(fc[ss], gp.y, a.y);
                            a.z = fmaf(fc[ss], gp.z, a.z); a.w = fmaf(fc[ss], gp.w, a.w);
                        }
                        float* dst = g.gw1 + lane * kH + 4 * j4;
                        atomicAdd(dst + 0, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
                    }
                    float b0 = 0.f, b1 = 0.f;
                    for (int ss = 0; ss < 32; ss++) { b0 += stg[ss * kStageStride + lane]; b1 += stg[ss * kStageStride + lane + 32]; }
                    atomicAdd(&g.gb1[lane], b0); atomicAdd(&g.gb1[lane + 32], b1);
                }
                __syncwarp();

                // (d) d(feature) -> featw (overwrites f), scatter-add into the plane gradients, tap dot-products -> stg
#pragma unroll
                for (int c = 0; c < kC; c++) featw[c * 33 + lane] = gf[c] * (1.0f / 3.0f);
                __syncwarp();
                {
                    const int u4 = (lane & 7) * 4, q = lane >> 3;
                    for (int r = 0; r < 8; r++) {
                        const int src = 4 * r + q;
                        const float g0 = featw[(u4 + 0) * 33 + src], g1 = featw[(u4 + 1) * 33 + src];
                        const float g2 = featw[(u4 + 2) * 33 + src], g3 = featw[(u4 + 3) * 33 + src];
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            const int base = __shfl_sync(0xffffffffu, fp.base[k], src);
                            const float wxa = __shfl_sync(0xffffffffu, fp.wxa[k], src);
                            const float wxb = __shfl_sync(0xffffffffu, fp.wxb[k], src);
                            const float wya = __shfl_sync(0xffffffffu, fp.wya[k], src);
                            const float wyb = __shfl_sync(0xffffffffu, fp.wyb[k], src);
                            const PT* t = img + base + u4;
                            float* gt = gimg + base + u4;
                            const int64_t offs[4] = {0, p.psX, p.psY, p.psY + p.psX};
                            const float ws[4] = {wya * wxa, wya * wxb, wyb * wxa, wyb * wxb};
                            float dots[4];
#pragma unroll
                            for (int tp = 0; tp < 4; tp++) {
                                const float4 v = ld_tex4<PT>(t + offs[tp]);
                                dots[tp] = v.x * g0 + v.y * g1 + v.z * g2 + v.w * g3;
                                if (ws[tp] != 0.f) red_add_v4(gt + offs[tp], ws[tp] * g0, ws[tp] * g1, ws[tp] * g2, ws[tp] * g3);
                            }
#pragma unroll
                            for (int tp = 0; tp < 4; tp++) {
                                dots[tp] += __shfl_xor_sync(0xffffffffu, dots[tp], 1);
                                dots[tp] += __shfl_xor_sync(0xffffffffu, dots[tp], 2);
                                dots[tp] += __shfl_xor_sync(0xffffffffu, dots[tp], 4);
                            }
                            if ((lane & 7) == 0) {
#pragma unroll
                                for (int tp = 0; tp < 4; tp++) stg[src * 12 + k * 4 + tp] = dots[tp];
                            }
                        }
                    }
                }
                __syncwarp();

                // (e) d(sample position) -> d(ray_o), d(ray_d)
                if (valid && (p.g_ray_o || p.g_ray_d)) {
                    const int P = p.o.P;
                    const float px = (s.ro[rl * 3 + 0] + tval * s.rd[rl * 3 + 0]) / p.o.box_half;
                    const float py = (s.ro[rl * 3 + 1] + tval * s.rd[rl * 3 + 1]) / p.o.box_half;
                    const float pz = (s.ro[rl * 3 + 2] + tval * s.rd[rl * 3 + 2]) / p.o.box_half;
                    const Axis ax = axis_footprint(px, P), ay = axis_footprint(py, P), az = axis_footprint(pz, P);
                    const Axis* U[3] = {&ax, &ax, &ay};
                    const Axis* V[3] = {&ay, &az, &az};
                    float du[3], dv[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const float d00 = stg[lane * 12 + k * 4 + 0], d01 = stg[lane * 12 + k * 4 + 1];
                        const float d10 = stg[lane * 12 + k * 4 + 2], d11 = stg[lane * 12 + k * 4 + 3];
                        du[k] = U[k]->da * (V[k]->wa * d00 + V[k]->wb * d10) + U[k]->db * (V[k]->wa * d01 + V[k]->wb * d11);
                        dv[k] = V[k]->da * (U[k]->wa * d00 + U[k]->wb * d01) + V[k]->db * (U[k]->wa * d10 + U[k]->wb * d11);
                    }
                    const float gpx = (du[0] + du[1]) * pixscale;   // x is the width coordinate of planes 0 and 1
                    const float gpy = (dv[0] + du[2]) * pixscale;   // y: height of plane 0, width of plane 2
                    const float gpz = (dv[1] + dv[2]) * pixscale;   // z: height of planes 1 and 2
                    atomicAdd(&g.gro[rl * 3 + 0], gpx); atomicAdd(&g.gro[rl * 3 + 1], gpy); atomicAdd(&g.gro[rl * 3 + 2], gpz);
                    atomicAdd(&g.grd[rl * 3 + 0], gpx * tval); atomicAdd(&g.grd[rl * 3 + 1], gpy * tval); atomicAdd(&g.grd[rl * 3 + 2], gpz * tval);
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int t = tid; t < nrays * 3; t += kThreads) {
            if (p.g_ray_o) p.g_ray_o[ray_base * 3 + t] = g.gro[t];
            if (p.g_ray_d) p.g_ray_d[ray_base * 3 + t] = g.grd[t];
        }
        __syncthreads();
    }

    // ---- flush the MLP parameter gradients of this CTA (chain rule through the runtime gains)
    const float g1 = rsqrtf((float)kC), g2 = rsqrtf((float)kH);
    for (int t = tid; t < kC * kH; t += kThreads) { const int c = t / kH, j = t - c * kH; atomicAdd(&p.g_w1[j * kC + c], g.gw1[t] * g1); }
    for (int t = tid; t < kH; t += kThreads) atomicAdd(&p.g_b1[t], g.gb1[t]);
    for (int t = tid; t < kH * 4; t += kThreads) { const int j = t >> 2, k = t & 3; atomicAdd(&p.g_w2[k * kH + j], g.gw2[t] * g2); }
    if (tid < 4) atomicAdd(&p.g_b2[tid], g.gb2[tid]);
}

template <class PT>
int launch_bwd(const Params& p, cudaStream_t st) {
    const size_t smem = ((Block<kTRB>::bytes(p.o.N) + 15) & ~(size_t)15) + BwdSmem::bytes(p.o.N) + 16;
    auto kern = raymarch_bwd_kernel<PT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_backward: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int nblocks = p.o.B * ((p.o.R + kTRB - 1) / kTRB);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int sms = GP3D_NUM_SMS;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms * per_sm;
    if (grid > nblocks) grid = nblocks;
    kern<<<grid, kThreads, smem, st>>>(p, nblocks);
    return 0;
}

}  // namespace rm

int gp3d_raymarch_check(const void* planes, int planes_dtype, int64_t psB, int64_t psP, int64_t psC, int64_t psY,
                        int64_t psX, const gp3d_raymarch_opts* o, const char* who);
int gp3d_raymarch_backward_v2(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st);   // raymarch_bwd2.cu

extern "C" int gp3d_raymarch_backward(const void* planes, int planes_dtype,
                                      int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                                      const float* ray_o, const float* ray_d,
                                      const float* w1, const float* b1, const float* w2, const float* b2,
                                      const float* u_coarse, const float* u_fine,
                                      const float* sn_coarse, const float* sn_fine,
                                      const float* g_rgb, const float* g_depth,
                                      float* g_planes, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                      float* g_ray_o, float* g_ray_d,
                                      const gp3d_raymarch_opts* opts, void* stream) {
    int rc = gp3d_raymarch_check(planes, planes_dtype, psB, psP, psC, psY, psX, opts, "raymarch_backward");
    if (rc != GP3D_OK) return rc;
    GP3D_CHECK_ARG(ray_o && ray_d && w1 && b1 && w2 && b2 && g_rgb && g_depth && g_planes && g_w1 && g_b1 && g_w2 && g_b2,
                   "raymarch_backward: null pointer");
    GP3D_CHECK_ARG((reinterpret_cast<uintptr_t>(g_planes) & 15u) == 0, "raymarch_backward: g_planes must be 16-byte aligned");
    rm::Params p{};
    p.planes = planes; p.psB = psB; p.psP = psP; p.psY = psY; p.psX = psX;
    p.ray_o = ray_o; p.ray_d = ray_d; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.u_coarse = u_coarse; p.u_fine = u_fine; p.sn_coarse = sn_coarse; p.sn_fine = sn_fine;
    p.g_rgb = g_rgb; p.g_depth = g_depth; p.g_planes = g_planes; p.g_w1 = g_w1; p.g_b1 = g_b1; p.g_w2 = g_w2; p.g_b2 = g_b2;
    p.g_ray_o = g_ray_o; p.g_ray_d = g_ray_d;
    p.o = *opts;
    cudaStream_t s = (cudaStream_t)stream;
    GP3D_CHECK_ARG(opts->mlp_mode >= 0 && opts->mlp_mode <= 2, "raymarch_backward: mlp_mode must be 0 (fp32 SIMT), 1 (TF32) or 2 (3xTF32)");
    int r;
    if (opts->mlp_mode == 0) r = (planes_dtype == GP3D_F32) ? rm::launch_bwd<float>(p, s) : rm::launch_bwd<__half>(p, s);
    else r = gp3d_raymarch_backward_v2(p, planes_dtype, opts->mlp_mode, s);
    if (r != 0) return r;
    GP3D_RETURN_LAUNCH();
}
