// Fused tri-plane ray-march, forward (sm_100a).
//
// One CTA (128 threads) owns TR consecutive rays of one image and runs, without touching HBM in between:
//   A. coarse pass   : stratified depths -> tri-plane gather (12 x 128-byte taps / sample) -> 2-layer MLP
//   B. per-ray       : coarse alpha-compositing weights (s-space) -> pdf/cdf -> inverse-CDF fine depths -> sort
//   C. fine pass     : same as A at the importance-sampled depths
//   D. per-ray       : merge coarse+fine by depth, alpha-composite in t-space -> rgb, depth, sum(w), T_final
// Algorithmic HBM traffic: every touched plane texel once + 24 B/ray in + 2*N*4 B/ray of injected variates
// (parity mode only) + 24 B/ray out.
//
// Gather layout: planes are channel-minor ([.., y, x, c], 32 channels = one 128-byte line per tap).  Eight lanes
// share one sample and fetch its tap as 8 x float4 (one fully-used L1 wavefront per tap); a warp therefore
// retires 4 samples per load instruction.  Footprints (texel base + per-axis weights) are computed once per sample
// (lane == sample) and broadcast with shuffles.  The interpolated 32-vector is transposed through a padded
// shared-memory tile so that the MLP runs with lane == sample.
#include "raymarch_common.cuh"

namespace rm {

template <int TR>
struct Smem {
    static __host__ __device__ int np(int N) { return N | 1; }
    static __host__ __device__ size_t bytes(int N) {
        size_t floats = kC * kH + kH + kH * 4 + 4 + kWarps * kC * 33 + TR * 3 * 2 + 3 * TR * np(N);
        floats = (floats + 3) & ~(size_t)3;
        return floats * 4 + 2 * (size_t)TR * (N + 1) * 16;
    }
};

// Interpolates the 32-channel feature of the warp's 32 samples into featw[c*33 + sample] (mean over 3 planes).
template <class PT>
__device__ __forceinline__ void gather_features(const PT* __restrict__ img, const Footprint& fp, float* featw,
                                                int64_t psX, int64_t psY, int lane) {
    const int u4 = (lane & 7) * 4, q = lane >> 3;
#pragma unroll 2
    for (int r = 0; r < 8; r++) {
        const int src = 4 * r + q;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int base = __shfl_sync(0xffffffffu, fp.base[k], src);
            const float wxa = __shfl_sync(0xffffffffu, fp.wxa[k], src);
            const float wxb = __shfl_sync(0xffffffffu, fp.wxb[k], src);
            const float wya = __shfl_sync(0xffffffffu, fp.wya[k], src);
            const float wyb = __shfl_sync(0xffffffffu, fp.wyb[k], src);
            const PT* t = img + base + u4;
            const float4 v00 = ld_tex4<PT>(t), v01 = ld_tex4<PT>(t + psX);
            const float4 v10 = ld_tex4<PT>(t + psY), v11 = ld_tex4<PT>(t + psY + psX);
            const float w00 = wya * wxa, w01 = wya * wxb, w10 = wyb * wxa, w11 = wyb * wxb;
            acc.x += w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
            acc.y += w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
            acc.z += w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
            acc.w += w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
        }
        const float third = 1.0f / 3.0f;   // x.mean(dim=1) over the three planes (networks_epigraf.py:54)
        featw[(u4 + 0) * 33 + src] = acc.x * third;
        featw[(u4 + 1) * 33 + src] = acc.y * third;
        featw[(u4 + 2) * 33 + src] = acc.z * third;
        featw[(u4 + 3) * 33 + src] = acc.w * third;
    }
    __syncwarp();
}

// TriPlaneMLP with lane == sample, fp32 SIMT (networks_epigraf.py:55, layers.py:42-58):
//   h = lrelu_0.2(f @ (W1/sqrt(32))^T + b1) * sqrt(2);  out = h @ (W2/sqrt(64))^T + b2.
__device__ __forceinline__ float4 mlp_simt(const float* featw, const float* w1s, const float* b1s,
                                           const float* w2s, const float* b2s, int lane) {
    float h[kH];
#pragma unroll
    for (int j = 0; j < kH; j++) h[j] = 0.f;
#pragma unroll 2
    for (int c = 0; c < kC; c++) {
        const float f = featw[c * 33 + lane];
        const float4* wr = reinterpret_cast<const float4*>(w1s + c * kH);
#pragma unroll
        for (int j4 = 0; j4 < kH / 4; j4++) {
            const float4 w = wr[j4];
            h[4 * j4 + 0] = fmaf(w.x, f, h[4 * j4 + 0]);
            h[4 * j4 + 1] = fmaf(w.y, f, h[4 * j4 + 1]);
            h[4 * j4 + 2] = fmaf(w.z, f, h[4 * j4 + 2]);
            h[4 * j4 + 3] = fmaf(w.w, f, h[4 * j4 + 3]);
        }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const float sqrt2 = 1.4142135623730951f;
#pragma unroll
    for (int j = 0; j < kH; j++) {
        float v = h[j] + b1s[j];
        v = (v > 0.f ? v : v * 0.2f) * sqrt2;
        const float4 w = reinterpret_cast<const float4*>(w2s)[j];
        o.x = fmaf(w.x, v, o.x); o.y = fmaf(w.y, v, o.y); o.z = fmaf(w.z, v, o.z); o.w = fmaf(w.w, v, o.w);
    }
    o.x += b2s[0]; o.y += b2s[1]; o.z += b2s[2]; o.w += b2s[3];
    return o;
}

template <class PT, int TR>
__global__ void __launch_bounds__(kThreads) raymarch_fwd_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = p.o.N, NP = Smem<TR>::np(N), R = p.o.R;
    float* w1s = reinterpret_cast<float*>(smem_raw);
    float* b1s = w1s + kC * kH;
    float* w2s = b1s + kH;
    float* b2s = w2s + kH * 4;
    float* feat = b2s + 4;
    float* ro = feat + kWarps * kC * 33;
    float* rd = ro + TR * 3;
    float* s_co = rd + TR * 3;
    float* cdf = s_co + TR * NP;
    float* s_fi = cdf + TR * NP;
    size_t foff = (size_t)(s_fi + TR * NP - w1s);
    foff = (foff + 3) & ~(size_t)3;
    float4* out_co = reinterpret_cast<float4*>(w1s + foff);
    float4* out_fi = out_co + TR * (N + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blocks_per_img = (R + TR - 1) / TR;
    const int b = blockIdx.x / blocks_per_img;
    const int r0 = (blockIdx.x - b * blocks_per_img) * TR;
    const int nrays = min(TR, R - r0);
    const int64_t ray_base = (int64_t)b * R + r0;
    const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;
    float* featw = feat + warp * kC * 33;

    // ---- stage MLP parameters (with the reference's runtime gains) and rays
    const float g1 = rsqrtf((float)kC), g2 = rsqrtf((float)kH);
    for (int t = tid; t < kC * kH; t += kThreads) { int c = t / kH, j = t - c * kH; w1s[t] = p.w1[j * kC + c] * g1; }
    for (int t = tid; t < kH; t += kThreads) b1s[t] = p.b1[t];
    for (int t = tid; t < kH * 4; t += kThreads) { int j = t >> 2, k = t & 3; w2s[t] = p.w2[k * kH + j] * g2; }
    if (tid < 4) b2s[tid] = p.b2[tid];
    for (int t = tid; t < nrays * 3; t += kThreads) { ro[t] = p.ray_o[ray_base * 3 + t]; rd[t] = p.ray_d[ray_base * 3 + t]; }
    __syncthreads();

    const float t0 = p.o.ray_start, t1 = p.o.ray_end, box = p.o.box_half;
    const int total = TR * N;

    // ---- passes A (coarse) and C (fine)
    for (int pass = 0; pass < 2; pass++) {
        float4* outp = pass ? out_fi : out_co;
        for (int s0 = 0; s0 < total; s0 += kThreads) {
            const int s = s0 + tid;
            const int rl = s / N, i = s - rl * N;
            const bool valid = (s < total) && (rl < nrays);
            Footprint fp;
#pragma unroll
            for (int k = 0; k < 3; k++) { fp.base[k] = 0; fp.wxa[k] = fp.wxb[k] = fp.wya[k] = fp.wyb[k] = 0.f; }
            if (valid) {
                float sd;
                if (pass == 0) {
                    const float u = p.u_coarse ? p.u_coarse[(ray_base + rl) * N + i]
                                               : rng_uniform(p.o, (uint64_t)(ray_base + rl), i, 0);
                    sd = coarse_s(i, N, u);
                    s_co[rl * NP + i] = sd;
                } else {
                    sd = s_fi[rl * NP + i];
                }
                const float t = s_to_t(sd, t0, t1);
                const float px = (ro[rl * 3 + 0] + t * rd[rl * 3 + 0]) / box;
                const float py = (ro[rl * 3 + 1] + t * rd[rl * 3 + 1]) / box;
                const float pz = (ro[rl * 3 + 2] + t * rd[rl * 3 + 2]) / box;
                sample_footprint(fp, px, py, pz, p);
            }
            gather_features<PT>(img, fp, featw, p.psX, p.psY, lane);
            float4 o = mlp_simt(featw, w1s, b1s, w2s, b2s, lane);
            __syncwarp();
            if (valid) {
                if (p.o.noise_std > 0.f) {
                    const float* sn = pass ? p.sn_fine : p.sn_coarse;
                    const float z = sn ? sn[(ray_base + rl) * N + i] : rng_normal(p.o, (uint64_t)(ray_base + rl), i, 2 + pass);
                    o.w += z * p.o.noise_std;
                }
                outp[rl * (N + 1) + i] = o;
            }
        }
        __syncthreads();

        if (pass == 0) {
            // ---- B: per-ray importance sampling (tri_plane_renderer.py:152-153, 237-295, 353-383)
            if (tid < nrays) {
                const int rl = tid;
                const float* sc = s_co + rl * NP;
                float* cd = cdf + rl * NP;
                float* sf = s_fi + rl * NP;
                float T = 1.f;
                for (int i = 0; i < N; i++) {
                    const float sig = density_act(out_co[rl * (N + 1) + i].w, p.o.clamp_mode);
                    const float delta = (i < N - 1) ? sc[i + 1] - sc[i] : (p.o.use_inf_depth ? 1e10f : 1e-3f);
                    const float alpha = 1.f - expf(-delta * sig);
                    cd[i] = alpha * T;
                    T *= (1.f - alpha + 1e-10f);
                }
                float sum = 0.f;
                for (int k = 1; k <= N - 2; k++) { const float w = (cd[k] + 1e-5f) + 1e-5f; cd[k] = w; sum += w; }
                float run = 0.f;
                cd[0] = 0.f;
                for (int k = 1; k <= N - 2; k++) { run += cd[k] / sum; cd[k] = run; }
                // cd[0..N-2] is the cdf (N-1 entries); bins[k] = 0.5 (s[k] + s[k+1]), k = 0..N-2
                for (int k = 0; k < N; k++) {
                    const float u = p.u_fine ? p.u_fine[(ray_base + rl) * N + k]
                                             : rng_uniform(p.o, (uint64_t)(ray_base + rl), k, 1);
                    int lo = 0, hi = N - 1;            // searchsorted(cdf, u, right=True) over N-1 entries
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cd[mid] <= u) lo = mid + 1; else hi = mid; }
                    const int below = max(lo - 1, 0), above = min(lo, N - 2);
                    const float c0 = cd[below], c1 = cd[above];
                    float den = c1 - c0;
                    if (den < 1e-5f) den = 1.f;
                    const float b0 = 0.5f * (sc[below] + sc[below + 1]);
                    const float b1v = 0.5f * (sc[above] + sc[above + 1]);
                    const float v = b0 + (u - c0) / den * (b1v - b0);
                    // insertion into the sorted prefix sf[0..k)
                    int j = k;
                    while (j > 0 && sf[j - 1] > v) { sf[j] = sf[j - 1]; j--; }
                    sf[j] = v;
                }
            }
            __syncthreads();
        }
    }

    // ---- D: merge + final compositing in t-space (tri_plane_renderer.py:163-166, 196-206, 353-405)
    if (tid < nrays) {
        const int rl = tid;
        const float* sc = s_co + rl * NP;
        const float* sf = s_fi + rl * NP;
        int ic = 0, jf = 0;
        // pops the next sample of the depth-merged sequence (coarse wins ties; both lists are ascending)
        auto pop = [&](float& t, float4& v) {
            const float tc = (ic < N) ? s_to_t(sc[ic], t0, t1) : 0.f;
            const float tf = (jf < N) ? s_to_t(sf[jf], t0, t1) : 0.f;
            const bool take_c = (jf >= N) || (ic < N && tc <= tf);
            if (take_c) { t = tc; v = out_co[rl * (N + 1) + ic]; ic++; }
            else        { t = tf; v = out_fi[rl * (N + 1) + jf]; jf++; }
        };
        float T = 1.f, wsum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f;
        float tcur, tnext = 0.f; float4 cur, nxt;
        pop(tcur, cur);
        nxt = cur;
        for (int m = 0; m < 2 * N; m++) {
            const bool last = (m == 2 * N - 1);
            if (!last) pop(tnext, nxt);
            const float delta = last ? (p.o.use_inf_depth ? 1e10f : 1e-3f) : (tnext - tcur);
            const float sig = density_act(cur.w, p.o.clamp_mode);
            const float alpha = 1.f - expf(-delta * sig);
            const float w = alpha * T;
            T *= (1.f - alpha + 1e-10f);
            wsum += w;
            cr += w * cur.x; cg += w * cur.y; cb += w * cur.z; dep += w * tcur;
            if (!last) { cur = nxt; tcur = tnext; }
        }
        const float wagg = wsum;
        if (p.o.last_back) {                       // weights[:, :, -1] += 1 - weights_agg  (:386-387)
            const float extra = 1.f - wagg;
            cr += extra * cur.x; cg += extra * cur.y; cb += extra * cur.z; dep += extra * tcur;
            wsum += extra;
        }
        if (p.o.white_back_end_idx > 0) {          // :392-395
            const float add = 1.f - wagg;
            cr += add;
            if (p.o.white_back_end_idx > 1) cg += add;
            if (p.o.white_back_end_idx > 2) cb += add;
        }
        const int64_t ro_i = ray_base + rl;
        p.rgb[ro_i * 3 + 0] = cr; p.rgb[ro_i * 3 + 1] = cg; p.rgb[ro_i * 3 + 2] = cb;
        p.depth[ro_i] = dep; p.wsum[ro_i] = wsum; p.tfinal[ro_i] = T;
    }
}

constexpr int kTR = 16;

template <class PT>
int launch_fwd(const Params& p, cudaStream_t s) {
    const size_t smem = Smem<kTR>::bytes(p.o.N);
    auto kern = raymarch_fwd_kernel<PT, kTR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_forward: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int blocks = p.o.B * ((p.o.R + kTR - 1) / kTR);
    kern<<<blocks, kThreads, smem, s>>>(p);
    return 0;
}

}  // namespace rm

int gp3d_raymarch_check(const void* planes, int planes_dtype, int64_t psB, int64_t psP, int64_t psC, int64_t psY,
                        int64_t psX, const gp3d_raymarch_opts* o, const char* who) {
    GP3D_CHECK_ARG(o != nullptr, "%s: opts is null", who);
    GP3D_CHECK_ARG(planes != nullptr, "%s: planes is null", who);
    GP3D_CHECK_ARG(o->B >= 1 && o->R >= 1, "%s: empty batch / ray set", who);
    GP3D_CHECK_ARG(o->N >= 3 && o->N <= rm::kMaxN, "%s: samples per pass N=%d outside [3, %d]", who, o->N, rm::kMaxN);
    GP3D_CHECK_ARG(o->P >= 2, "%s: plane resolution must be >= 2", who);
    GP3D_CHECK_ARG(planes_dtype == GP3D_F32 || planes_dtype == GP3D_F16, "%s: planes must be float32 or float16", who);
    if (o->C != rm::kC || o->H != rm::kH) {
        gp3d_set_error("%s: only feat_dim=32 / hid_dim=64 / n_layers=2 tri-plane MLPs are built (got C=%d H=%d)", who, o->C, o->H);
        return GP3D_E_UNSUPPORTED;
    }
    const int64_t al = planes_dtype == GP3D_F32 ? 4 : 4;   // 4 channels per lane -> 16 B (f32) / 8 B (f16)
    if (psC != 1 || psX % al || psY % al || psP % al || psB % al || (reinterpret_cast<uintptr_t>(planes) & 15u)) {
        gp3d_set_error("%s: planes must be channel-minor (stride_c == 1) with 16-byte aligned texels; got strides "
                       "B=%lld P=%lld C=%lld Y=%lld X=%lld", who, (long long)psB, (long long)psP, (long long)psC,
                       (long long)psY, (long long)psX);
        return GP3D_E_UNSUPPORTED;
    }
    GP3D_CHECK_ARG((int64_t)3 * o->C * o->P * o->P * 4 < 2147483647LL || psB < 2147483647LL, "%s: image too large", who);
    return GP3D_OK;
}

extern "C" int gp3d_raymarch_forward(const void* planes, int planes_dtype,
                                     int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                                     const float* ray_o, const float* ray_d,
                                     const float* w1, const float* b1, const float* w2, const float* b2,
                                     const float* u_coarse, const float* u_fine,
                                     const float* sn_coarse, const float* sn_fine,
                                     float* rgb, float* depth, float* wsum, float* tfinal,
                                     const gp3d_raymarch_opts* opts, void* stream) {
    int rc = gp3d_raymarch_check(planes, planes_dtype, psB, psP, psC, psY, psX, opts, "raymarch_forward");
    if (rc != GP3D_OK) return rc;
    GP3D_CHECK_ARG(ray_o && ray_d && w1 && b1 && w2 && b2 && rgb && depth && wsum && tfinal, "raymarch_forward: null pointer");
    rm::Params p{};
    p.planes = planes; p.psB = psB; p.psP = psP; p.psY = psY; p.psX = psX;
    p.ray_o = ray_o; p.ray_d = ray_d; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.u_coarse = u_coarse; p.u_fine = u_fine; p.sn_coarse = sn_coarse; p.sn_fine = sn_fine;
    p.rgb = rgb; p.depth = depth; p.wsum = wsum; p.tfinal = tfinal;
    p.o = *opts;
    cudaStream_t s = (cudaStream_t)stream;
    int r = (planes_dtype == GP3D_F32) ? rm::launch_fwd<float>(p, s) : rm::launch_fwd<__half>(p, s);
    if (r != 0) return r;
    GP3D_RETURN_LAUNCH();
}
