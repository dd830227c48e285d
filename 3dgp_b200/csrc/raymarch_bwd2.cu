// Fused tri-plane ray-march, backward, second generation (sm_100a).  Same contract as raymarch_bwd.cu (nothing saved by the
// forward; gradients w.r.t. planes, MLP parameters and, optionally, rays), reorganised like the second-generation forward:
//   * the forward is re-run with raymarch2.cuh's pipeline (tensor-core MLP, parallel per-ray phases), keeping T_m;
//   * D' (compositing backward) is parallel over (ray, sample); only the transmittance / suffix scans are serial per ray;
//   * E (per-sample backward) runs every contraction on mma.sync.m16n8k8 TF32 / 3xTF32 with the sample axis as M or K:
//       layer-1 recompute  c   = f  . W1          [16 x 32] x [32 x 64]
//       dW2 += act(c)^T . go                      [64 x 16] x [16 x 4]
//       g_pre = (go . W2^T) * act'(c)             SIMT, in the layer-1 C-fragment registers
//       dW1 += f^T . g_pre                        [32 x 16] x [16 x 64]
//       df   = g_pre . W1^T                       [16 x 64] x [64 x 32]   (A fragments = the C fragments, as in the forward)
//     v and g_pre pass through a 16 x 64 per-warp staging tile to be read back transposed;
//   * d(feature) is scattered into the plane gradients with red.global.add.v4.f32 from the staged footprints (4 taps x 3 planes).
// Parameter gradients are reduced in shared memory and flushed once per CTA (persistent grid).
#include "raymarch2.cuh"

namespace rm2 {

constexpr int STG = 72;      // staging row stride (floats): bank = 8t + g for the transposed fragment reads
constexpr int GW1S = 68;     // dW1 accumulator row stride

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct BSmem {
    float4* w2p;                  // [64]  (w2g[0..3][j]) with the runtime gain
    float* Tb; float* X;          // [TR][2N+1]
    float* gray;                  // [TR][4]  upstream gradients (rgb, depth)
    float* Kb; float* Wagg;       // [TR]
    float* stg;                   // [kWarps][16][STG]
    float4* gow;                  // [kWarps][32]
    float* gw1; float* gb1; float* gw2; float* gb2;   // [32][GW1S], [64], [64][4], [4]   (contiguous)
    float* gro; float* grd;       // [TR][3]
    static __host__ __device__ size_t bytes(int N) {
        size_t b = 64 * 16 + 2 * (size_t)TR * (2 * N + 1) * 4 + TR * 16 + 2 * TR * 4;
        b = (b + 15) & ~(size_t)15;
        b += (size_t)kWarps * 16 * STG * 4 + (size_t)kWarps * 32 * 16;
        b += (size_t)(32 * GW1S + 64 + 256 + 4) * 4 + TR * 6 * 4;
        return b + 32;
    }
    __device__ void carve(unsigned char* raw, int N) {
        w2p = reinterpret_cast<float4*>(raw);
        Tb = reinterpret_cast<float*>(w2p + 64); X = Tb + TR * (2 * N + 1);
        gray = X + TR * (2 * N + 1); Kb = gray + TR * 4; Wagg = Kb + TR;
        size_t off = (size_t)(reinterpret_cast<unsigned char*>(Wagg + TR) - raw);
        off = (off + 15) & ~(size_t)15;
        stg = reinterpret_cast<float*>(raw + off);
        gow = reinterpret_cast<float4*>(stg + kWarps * 16 * STG);
        gw1 = reinterpret_cast<float*>(gow + kWarps * 32);
        gb1 = gw1 + 32 * GW1S; gw2 = gb1 + 64; gb2 = gw2 + 256;
        gro = gb2 + 4; grd = gro + TR * 3;
    }
};

// acc += a * b with a = ah + al, b = bh + bl (3xTF32: the al*bl term is below fp32 resolution)
template <int MODE>
__device__ __forceinline__ void mma3(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma_tf32(acc, ah, bh0, bh1);
    if (MODE == 2) { mma_tf32(acc, ah, bl0, bl1); mma_tf32(acc, al, bh0, bh1); }
}

template <class PT, int MODE>
__global__ void __launch_bounds__(kThreads, 2) raymarch_bwd2_kernel(Params p, int nblocks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s;
    s.carve(smem_raw, p.o.N);
    BSmem g;
    g.carve(smem_raw + ((Smem::bytes(p.o.N) + 15) & ~(size_t)15), p.o.N);
    const int N = p.o.N, NP = s.NP, R = p.o.R, M2 = 2 * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int blocks_per_img = (R + TR - 1) / TR;
    float* featw = s.feat + warp * 32 * FSTR;
    uint32_t* fprw = s.fpr + warp * 32 * FPSTR;
    float* stg = g.stg + warp * 16 * STG;
    float4* gow = g.gow + warp * 32;
    const float t0 = p.o.ray_start, t1 = p.o.ray_end;
    const float big = p.o.use_inf_depth ? 1e10f : 1e-3f;
    const bool want_rays = (p.g_ray_o != nullptr) || (p.g_ray_d != nullptr);
    const float sqrt2 = 1.4142135623730951f;

    // parameter-gradient accumulators live in registers for the whole (persistent) CTA: shared-memory fp32 atomics are CAS loops
    float aW1[2][8][4], aW2[4][4], aB1[8][2], aB2[4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) { aW1[i][j][0] = aW1[i][j][1] = aW1[i][j][2] = aW1[i][j][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < 4; i++) { aW2[i][0] = aW2[i][1] = aW2[i][2] = aW2[i][3] = 0.f; aB2[i] = 0.f; }
#pragma unroll
    for (int j = 0; j < 8; j++) { aB1[j][0] = aB1[j][1] = 0.f; }

    stage_weights(s, p);
    {
        const float g2 = rsqrtf((float)kH);
        for (int j = tid; j < kH; j += kThreads) g.w2p[j] = make_float4(p.w2[0 * kH + j] * g2, p.w2[1 * kH + j] * g2, p.w2[2 * kH + j] * g2, p.w2[3 * kH + j] * g2);
        for (int t = tid; t < 32 * GW1S + 64 + 256 + 4; t += kThreads) g.gw1[t] = 0.f;
    }
    __syncthreads();

    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int b = blk / blocks_per_img;
        const int r0 = (blk - b * blocks_per_img) * TR;
        const int nrays = min(TR, R - r0);
        const int64_t ray_base = (int64_t)b * R + r0;
        const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;
        float* gimg = p.g_planes + (int64_t)b * p.psB;

        for (int t = tid; t < nrays * 3; t += kThreads) { s.ro[t] = p.ray_o[ray_base * 3 + t]; s.rd[t] = p.ray_d[ray_base * 3 + t]; }
        for (int t = tid; t < TR * 6; t += kThreads) g.gro[t] = 0.f;                 // gro, grd contiguous
        if (tid < nrays) {
            g.gray[tid * 4 + 0] = p.g_rgb[(ray_base + tid) * 3 + 0]; g.gray[tid * 4 + 1] = p.g_rgb[(ray_base + tid) * 3 + 1];
            g.gray[tid * 4 + 2] = p.g_rgb[(ray_base + tid) * 3 + 2]; g.gray[tid * 4 + 3] = p.g_depth[ray_base + tid];
        }
        __syncthreads();

        forward_phases2<PT, MODE, true>(s, p, img, ray_base, nrays, g.Tb);

        float* s_fi = s.bufA;
        auto depth_of = [&](int rl, int code) { return s_to_t(code < N ? s.s_co[rl * NP + code] : s_fi[rl * NP + code - N], t0, t1); };
        auto value_ptr = [&](int rl, int code) { return code < N ? &s.out_co[rl * (N + 1) + code] : &s.out_fi[rl * (N + 1) + code - N]; };

        // ---- D': compositing backward.  K: terms that multiply (1 - sum w)
        if (tid < nrays) {
            const int rl = tid, code = s.ord[rl * M2 + M2 - 1];
            const float4 vl = *value_ptr(rl, code);
            const float gr = g.gray[rl * 4], gg = g.gray[rl * 4 + 1], gb = g.gray[rl * 4 + 2], gd = g.gray[rl * 4 + 3];
            float K = 0.f;
            if (p.o.last_back) K += gr * vl.x + gg * vl.y + gb * vl.z + gd * depth_of(rl, code);
            if (p.o.white_back_end_idx > 0) K += gr + (p.o.white_back_end_idx > 1 ? gg : 0.f) + (p.o.white_back_end_idx > 2 ? gb : 0.f);
            g.Kb[rl] = K;
        }
        __syncthreads();
        for (int e = tid; e < nrays * M2; e += kThreads) {                  // X_m = w_m G_m
            const int rl = e / M2, m = e - rl * M2;
            const int code = s.ord[rl * M2 + m];
            const float4 v = *value_ptr(rl, code);
            const float Gm = g.gray[rl * 4] * v.x + g.gray[rl * 4 + 1] * v.y + g.gray[rl * 4 + 2] * v.z + g.gray[rl * 4 + 3] * depth_of(rl, code) - g.Kb[rl];
            g.X[rl * (M2 + 1) + m] = s.wm[rl * (M2 + 1) + m] * Gm;
        }
        __syncthreads();
        if (tid < nrays) {                                                  // exclusive suffix sums S_m = sum_{j>m} w_j G_j, and sum w
            float* x = g.X + tid * (M2 + 1);
            const float* w = s.wm + tid * (M2 + 1);
            float S = 0.f, wagg = 0.f;
            for (int m = M2 - 1; m >= 0; m--) { const float v = x[m]; x[m] = S; S += v; wagg += w[m]; }
            g.Wagg[tid] = wagg;
        }
        __syncthreads();
        for (int e = tid; e < nrays * M2; e += kThreads) {                  // d(rgb_m, sigma_m), written over the sample's value
            const int rl = e / M2, m = e - rl * M2;
            const int code = s.ord[rl * M2 + m];
            float4* vp = value_ptr(rl, code);
            const float4 v = *vp;
            const float tm = depth_of(rl, code);
            const float delta = (m == M2 - 1) ? big : depth_of(rl, s.ord[rl * M2 + m + 1]) - tm;
            const float ex = expf(-delta * density_act(v.w, p.o.clamp_mode));
            const float alpha = 1.f - ex;
            const float Tm = g.Tb[rl * (M2 + 1) + m];
            const float w = alpha * Tm;
            const float gr = g.gray[rl * 4], gg = g.gray[rl * 4 + 1], gb = g.gray[rl * 4 + 2], gd = g.gray[rl * 4 + 3];
            const float Gm = gr * v.x + gg * v.y + gb * v.z + gd * tm - g.Kb[rl];
            const float dalpha = Tm * Gm - g.X[rl * (M2 + 1) + m] / (1.f - alpha + 1e-10f);
            const float dsig = dalpha * (delta * ex) * density_act_grad(v.w, p.o.clamp_mode);
            float4 gv = make_float4(w * gr, w * gg, w * gb, dsig);
            if (m == M2 - 1 && p.o.last_back) { const float extra = 1.f - g.Wagg[rl]; gv.x += extra * gr; gv.y += extra * gg; gv.z += extra * gb; }
            *vp = gv;
        }
        __syncthreads();

        // ---- E: per-sample backward through the MLP and the plane interpolation
        const int total = TR * N;
        const float pixscale = 0.5f * (float)(p.o.P - 1) / p.o.box_half;
        for (int pass = 0; pass < 2; pass++) {
            const float4* gsrc = pass ? s.out_fi : s.out_co;
            for (int s0 = 0; s0 < total; s0 += kThreads) {
                const int si = s0 + tid;
                const int rl = si / N, i = si - rl * N;
                const bool valid = (si < total) && (rl < nrays);
                float sd = 0.f;
                if (valid) sd = pass ? s_fi[rl * NP + i] : s.s_co[rl * NP + i];
                stage_footprint(fprw + lane * FPSTR, p, s.ro + (valid ? rl : 0) * 3, s.rd + (valid ? rl : 0) * 3, sd, valid);
                const float4 go_own = valid ? gsrc[rl * (N + 1) + i] : make_float4(0.f, 0.f, 0.f, 0.f);
                gow[lane] = go_own;
                __syncwarp();
                gather_tile<PT>(img, fprw, featw, p.psX, p.psY, lane);
                aB2[0] += go_own.x; aB2[1] += go_own.y; aB2[2] += go_own.z; aB2[3] += go_own.w;      // db2
                const int lm = lane >> 3, lr = lane & 7;
#pragma unroll 1
                for (int mt = 0; mt < 2; mt++) {
                    // layer-1 recompute: c[j] = pre-activations of rows gq / gq+8 (of this 16-sample half), hidden 8j+2tq / +1
                    float c[8][4];
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float bb0 = s.b1s[8 * j + 2 * tq], bb1 = s.b1s[8 * j + 2 * tq + 1];
                        c[j][0] = bb0; c[j][1] = bb1; c[j][2] = bb0; c[j][3] = bb1;
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                        uint32_t a[4], ah[4], al[4];
                        ldmatrix_x4(a, featw + (mt * 16 + (lm & 1) * 8 + lr) * FSTR + ks * 8 + (lm >> 1) * 4);
#pragma unroll
                        for (int q = 0; q < 4; q++) split_tf32<MODE>(__uint_as_float(a[q]), ah[q], al[q]);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float2 bh = s.w1h[(ks * 8 + j) * 32 + lane];
                            float2 bl = make_float2(0.f, 0.f);
                            if (MODE == 2) bl = s.w1l[(ks * 8 + j) * 32 + lane];
                            mma3<MODE>(c[j], ah, al, __float_as_uint(bh.x), __float_as_uint(bh.y), __float_as_uint(bl.x), __float_as_uint(bl.y));
                        }
                    }
                    // (a) v = act(c) -> staging tile [sample][hidden]
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; q++) v[q] = (c[j][q] > 0.f ? c[j][q] : 0.2f * c[j][q]) * sqrt2;
                        *reinterpret_cast<float2*>(stg + gq * STG + 8 * j + 2 * tq) = make_float2(v[0], v[1]);
                        *reinterpret_cast<float2*>(stg + (gq + 8) * STG + 8 * j + 2 * tq) = make_float2(v[2], v[3]);
                    }
                    __syncwarp();
                    // dW2[j][k] += sum_s v[s][j] go[s][k]:  A = v^T (hidden x sample), B = go (sample x 8, columns >= 4 are zero)
                    {
                        uint32_t bh[2][2], bl[2][2];
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
                            const float* ga = reinterpret_cast<const float*>(gow + mt * 16 + 8 * kk + tq);
                            const float* gb_ = reinterpret_cast<const float*>(gow + mt * 16 + 8 * kk + tq + 4);
                            const float x0 = (gq < 4) ? ga[gq & 3] : 0.f, x1 = (gq < 4) ? gb_[gq & 3] : 0.f;
                            split_tf32<MODE>(x0, bh[kk][0], bl[kk][0]); split_tf32<MODE>(x1, bh[kk][1], bl[kk][1]);
                        }
#pragma unroll
                        for (int mj = 0; mj < 4; mj++) {
#pragma unroll
                            for (int kk = 0; kk < 2; kk++) {
                                uint32_t ah[4], al[4];
                                split_tf32<MODE>(stg[(8 * kk + tq) * STG + 16 * mj + gq], ah[0], al[0]);
                                split_tf32<MODE>(stg[(8 * kk + tq) * STG + 16 * mj + gq + 8], ah[1], al[1]);
                                split_tf32<MODE>(stg[(8 * kk + tq + 4) * STG + 16 * mj + gq], ah[2], al[2]);
                                split_tf32<MODE>(stg[(8 * kk + tq + 4) * STG + 16 * mj + gq + 8], ah[3], al[3]);
                                mma3<MODE>(aW2[mj], ah, al, bh[kk][0], bh[kk][1], bl[kk][0], bl[kk][1]);       // C: hidden 16mj+gq(+8) x k = 2tq(+1), valid for tq < 2
                            }
                        }
                    }
                    __syncwarp();
                    // (b) g_pre = (go . W2^T) * act'(c) * sqrt2, in place of c; db1
                    {
                        const float4 g0 = gow[mt * 16 + gq], g1 = gow[mt * 16 + gq + 8];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const float4 wa = g.w2p[8 * j + 2 * tq], wb = g.w2p[8 * j + 2 * tq + 1];
                            const float h0 = g0.x * wa.x + g0.y * wa.y + g0.z * wa.z + g0.w * wa.w;
                            const float h1 = g0.x * wb.x + g0.y * wb.y + g0.z * wb.z + g0.w * wb.w;
                            const float h2 = g1.x * wa.x + g1.y * wa.y + g1.z * wa.z + g1.w * wa.w;
                            const float h3 = g1.x * wb.x + g1.y * wb.y + g1.z * wb.z + g1.w * wb.w;
                            c[j][0] = h0 * (c[j][0] > 0.f ? 1.f : 0.2f) * sqrt2; c[j][1] = h1 * (c[j][1] > 0.f ? 1.f : 0.2f) * sqrt2;
                            c[j][2] = h2 * (c[j][2] > 0.f ? 1.f : 0.2f) * sqrt2; c[j][3] = h3 * (c[j][3] > 0.f ? 1.f : 0.2f) * sqrt2;
                            aB1[j][0] += c[j][0] + c[j][2]; aB1[j][1] += c[j][1] + c[j][3];      // db1, reduced over the 8 row-lanes at the end
                            *reinterpret_cast<float2*>(stg + gq * STG + 8 * j + 2 * tq) = make_float2(c[j][0], c[j][1]);
                            *reinterpret_cast<float2*>(stg + (gq + 8) * STG + 8 * j + 2 * tq) = make_float2(c[j][2], c[j][3]);
                        }
                    }
                    __syncwarp();
                    // (c) dW1[ch][j] += sum_s f[s][ch] g_pre[s][j]:  A = f^T (channel x sample) from the feature tile, B = g_pre from staging
#pragma unroll
                    for (int mi = 0; mi < 2; mi++) {
                        uint32_t ah[2][4], al[2][4];
#pragma unroll
                        for (int kk = 0; kk < 2; kk++) {
                            const float* fr0 = featw + (mt * 16 + 8 * kk + tq) * FSTR + 16 * mi + gq;
                            const float* fr1 = featw + (mt * 16 + 8 * kk + tq + 4) * FSTR + 16 * mi + gq;
                            split_tf32<MODE>(fr0[0], ah[kk][0], al[kk][0]); split_tf32<MODE>(fr0[8], ah[kk][1], al[kk][1]);
                            split_tf32<MODE>(fr1[0], ah[kk][2], al[kk][2]); split_tf32<MODE>(fr1[8], ah[kk][3], al[kk][3]);
                        }
#pragma unroll
                        for (int nj = 0; nj < 8; nj++) {
#pragma unroll
                            for (int kk = 0; kk < 2; kk++) {
                                uint32_t bh0, bl0, bh1, bl1;
                                split_tf32<MODE>(stg[(8 * kk + tq) * STG + 8 * nj + gq], bh0, bl0);
                                split_tf32<MODE>(stg[(8 * kk + tq + 4) * STG + 8 * nj + gq], bh1, bl1);
                                mma3<MODE>(aW1[mi][nj], ah[kk], al[kk], bh0, bh1, bl0, bl1);                      // C: channel 16mi+gq(+8) x hidden 8nj+2tq(+1)
                            }
                        }
                    }
                    // (d) df = g_pre . W1^T  (A fragments = C fragments with the forward's k permutation; B gathered from the forward's W1 fragments)
                    float d[4][4];
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f; }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        uint32_t ah[4], al[4];
                        split_tf32<MODE>(c[j][0], ah[0], al[0]); split_tf32<MODE>(c[j][2], ah[1], al[1]);
                        split_tf32<MODE>(c[j][1], ah[2], al[2]); split_tf32<MODE>(c[j][3], ah[3], al[3]);
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) {
                            // W1g[hidden h][channel ch] sits in the forward fragment (ks = ch/8, j = h/8) at lane (h%8)*4 + (ch%8)%4, component (ch%8)/4
                            const int e0 = ((nt * 8 + j) * 32 + 8 * tq + (gq & 3)) * 2 + (gq >> 2);     // h = 8j + 2tq,     ch = 8nt + gq
                            const int e1 = e0 + 8;                                                       // h = 8j + 2tq + 1
                            const float* wh = reinterpret_cast<const float*>(s.w1h);
                            const float* wl = reinterpret_cast<const float*>(s.w1l);
                            const uint32_t bh0 = __float_as_uint(wh[e0]), bh1 = __float_as_uint(wh[e1]);
                            uint32_t bl0 = 0u, bl1 = 0u;
                            if (MODE == 2) { bl0 = __float_as_uint(wl[e0]); bl1 = __float_as_uint(wl[e1]); }
                            mma3<MODE>(d[nt], ah, al, bh0, bh1, bl0, bl1);
                        }
                    }
                    __syncwarp();      // every lane is done reading this half of the feature tile
                    const float third = 1.0f / 3.0f;
#pragma unroll
                    for (int nt = 0; nt < 4; nt++) {
                        *reinterpret_cast<float2*>(featw + (mt * 16 + gq) * FSTR + 8 * nt + 2 * tq) = make_float2(d[nt][0] * third, d[nt][1] * third);
                        *reinterpret_cast<float2*>(featw + (mt * 16 + gq + 8) * FSTR + 8 * nt + 2 * tq) = make_float2(d[nt][2] * third, d[nt][3] * third);
                    }
                }
                __syncwarp();

                // (e) scatter-add d(feature) into the plane gradients; optional tap dot-products for d(sample position)
                {
                    const int u4 = (lane & 7) * 4, q = lane >> 3;
                    for (int r = 0; r < 8; r++) {
                        const int src = 4 * r + q;
                        const uint4* rec = reinterpret_cast<const uint4*>(fprw + src * FPSTR);
                        const uint4 bs = rec[0];
                        const uint4 wa = rec[1], wb = rec[2], wc = rec[3];
                        const float w[12] = {__uint_as_float(wa.x), __uint_as_float(wa.y), __uint_as_float(wa.z), __uint_as_float(wa.w),
                                             __uint_as_float(wb.x), __uint_as_float(wb.y), __uint_as_float(wb.z), __uint_as_float(wb.w),
                                             __uint_as_float(wc.x), __uint_as_float(wc.y), __uint_as_float(wc.z), __uint_as_float(wc.w)};
                        const uint32_t bases[3] = {bs.x, bs.y, bs.z};
                        const float4 gf = *reinterpret_cast<const float4*>(featw + src * FSTR + u4);
                        const int64_t offs[4] = {0, p.psX, p.psY, p.psY + p.psX};
#pragma unroll
                        for (int k = 0; k < 3; k++) {
                            float* gt = gimg + (int)bases[k] + u4;
#pragma unroll
                            for (int tp = 0; tp < 4; tp++) {
                                const float ww = w[4 * k + tp];
                                if (ww != 0.f) red_add_v4(gt + offs[tp], ww * gf.x, ww * gf.y, ww * gf.z, ww * gf.w);
                            }
                            if (want_rays) {
                                const PT* tx = img + (int)bases[k] + u4;
                                float dots[4];
#pragma unroll
                                for (int tp = 0; tp < 4; tp++) {
                                    const float4 v = ld_tex4<PT>(tx + offs[tp]);
                                    float dd = v.x * gf.x + v.y * gf.y + v.z * gf.z + v.w * gf.w;
                                    dd += __shfl_xor_sync(0xffffffffu, dd, 1); dd += __shfl_xor_sync(0xffffffffu, dd, 2); dd += __shfl_xor_sync(0xffffffffu, dd, 4);
                                    dots[tp] = dd;
                                }
                                if ((lane & 7) == 0) {
#pragma unroll
                                    for (int tp = 0; tp < 4; tp++) stg[src * 12 + k * 4 + tp] = dots[tp];
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                if (want_rays) {
                    if (valid) {
                        const int P = p.o.P;
                        const float tval = s_to_t(sd, t0, t1);
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
                        const float gpx = (du[0] + du[1]) * pixscale, gpy = (dv[0] + du[2]) * pixscale, gpz = (dv[1] + dv[2]) * pixscale;
                        atomicAdd(&g.gro[rl * 3 + 0], gpx); atomicAdd(&g.gro[rl * 3 + 1], gpy); atomicAdd(&g.gro[rl * 3 + 2], gpz);
                        atomicAdd(&g.grd[rl * 3 + 0], gpx * tval); atomicAdd(&g.grd[rl * 3 + 1], gpy * tval); atomicAdd(&g.grd[rl * 3 + 2], gpz * tval);
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        if (want_rays) {
            for (int t = tid; t < nrays * 3; t += kThreads) {
                if (p.g_ray_o) p.g_ray_o[ray_base * 3 + t] = g.gro[t];
                if (p.g_ray_d) p.g_ray_d[ray_base * 3 + t] = g.grd[t];
            }
        }
        __syncthreads();
    }

    // ---- reduce the four warps' register accumulators in shared memory (one warp at a time, plain read-modify-write)
#pragma unroll
    for (int j = 0; j < 8; j++) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { aB1[j][0] += __shfl_xor_sync(0xffffffffu, aB1[j][0], o); aB1[j][1] += __shfl_xor_sync(0xffffffffu, aB1[j][1], o); }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) aB2[k] += __shfl_xor_sync(0xffffffffu, aB2[k], o);
    }
    for (int w = 0; w < kWarps; w++) {
        if (warp == w) {
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int nj = 0; nj < 8; nj++) {
                    float* d0 = g.gw1 + (16 * mi + gq) * GW1S + 8 * nj + 2 * tq;
                    d0[0] += aW1[mi][nj][0]; d0[1] += aW1[mi][nj][1]; d0[8 * GW1S] += aW1[mi][nj][2]; d0[8 * GW1S + 1] += aW1[mi][nj][3];
                }
            if (tq < 2) {
#pragma unroll
                for (int mj = 0; mj < 4; mj++) {
                    g.gw2[(16 * mj + gq) * 4 + 2 * tq] += aW2[mj][0]; g.gw2[(16 * mj + gq) * 4 + 2 * tq + 1] += aW2[mj][1];
                    g.gw2[(16 * mj + gq + 8) * 4 + 2 * tq] += aW2[mj][2]; g.gw2[(16 * mj + gq + 8) * 4 + 2 * tq + 1] += aW2[mj][3];
                }
            }
            if (gq == 0) {
#pragma unroll
                for (int j = 0; j < 8; j++) { g.gb1[8 * j + 2 * tq] += aB1[j][0]; g.gb1[8 * j + 2 * tq + 1] += aB1[j][1]; }
            }
            if (lane < 4) g.gb2[lane] += (lane == 0 ? aB2[0] : lane == 1 ? aB2[1] : lane == 2 ? aB2[2] : aB2[3]);
        }
        __syncthreads();
    }
    // ---- flush this CTA's parameter gradients (chain rule through the runtime gains, layers.py:39,47)
    const float g1 = rsqrtf((float)kC), g2 = rsqrtf((float)kH);
    for (int t = tid; t < kC * kH; t += kThreads) { const int ch = t / kH, j = t - ch * kH; atomicAdd(&p.g_w1[j * kC + ch], g.gw1[ch * GW1S + j] * g1); }
    for (int t = tid; t < kH; t += kThreads) atomicAdd(&p.g_b1[t], g.gb1[t]);
    for (int t = tid; t < kH * 4; t += kThreads) { const int j = t >> 2, k = t & 3; atomicAdd(&p.g_w2[k * kH + j], g.gw2[t] * g2); }
    if (tid < 4) atomicAdd(&p.g_b2[tid], g.gb2[tid]);
}

template <class PT, int MODE>
int launch_bwd2(const Params& p, cudaStream_t st) {
    const size_t smem = ((Smem::bytes(p.o.N) + 15) & ~(size_t)15) + BSmem::bytes(p.o.N);
    auto kern = raymarch_bwd2_kernel<PT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_backward(v2): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int nblocks = p.o.B * ((p.o.R + TR - 1) / TR);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int sms = GP3D_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms * per_sm;
    if (grid > nblocks) grid = nblocks;
    kern<<<grid, kThreads, smem, st>>>(p, nblocks);
    return 0;
}

}  // namespace rm2

int gp3d_raymarch_backward_v2(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st) {
    if (planes_dtype == GP3D_F32) return mode == 2 ? rm2::launch_bwd2<float, 2>(p, st) : rm2::launch_bwd2<float, 1>(p, st);
    return mode == 2 ? rm2::launch_bwd2<__half, 2>(p, st) : rm2::launch_bwd2<__half, 1>(p, st);
}
