// Per-CTA building blocks of the fused ray-march (shared by forward and backward kernels).
//
// One CTA (128 threads) owns TR consecutive rays of one image:
//   A. coarse pass   : stratified depths -> tri-plane gather (12 x 128-byte taps / sample) -> 2-layer MLP
//   B. per-ray       : coarse alpha-compositing weights (s-space) -> pdf/cdf -> inverse-CDF fine depths -> sort
//   C. fine pass     : same as A at the importance-sampled depths
// leaving (rgb, sigma) of all 2N samples of every ray plus both depth lists in shared memory.
//
// Gather layout: planes are channel-minor ([.., y, x, c], 32 channels = one 128-byte line per tap).  Eight lanes
// share one sample and fetch its tap as 8 x float4 (one fully-used L1 wavefront per tap); a warp retires 4 samples per
// load instruction.  Footprints (texel base + per-axis weights) are computed once per sample (lane == sample) and
// broadcast with shuffles.  The interpolated 32-vector is transposed through a padded shared-memory tile so that the
// MLP runs with lane == sample.
#pragma once
#include "raymarch_common.cuh"

namespace rm {

template <int TR>
struct Block {
    float* w1s;    // [kC][kH]   w1s[c*kH + j] = W1[j][c] / sqrt(kC)
    float* b1s;    // [kH]
    float* w2s;    // [kH][4]    w2s[j*4 + k] = W2[k][j] / sqrt(kH)
    float* b2s;    // [4]
    float* feat;   // [kWarps][kC][33]
    float* ro;     // [TR][3]
    float* rd;     // [TR][3]
    float* s_co;   // [TR][NP]   coarse depths, s-space
    float* cdf;    // [TR][NP]   scratch: coarse weights -> cdf
    float* s_fi;   // [TR][NP]   fine depths, s-space, ascending
    float4* out_co;  // [TR][N+1] (r,g,b,sigma) of coarse samples
    float4* out_fi;  // [TR][N+1]
    unsigned char* fperm;   // [TR][N]  fperm[k] = index (in u_fine order) of the k-th smallest fine sample
    int N, NP;

    static __host__ __device__ int np(int N) { return N | 1; }
    static __host__ __device__ size_t floats_before_out(int N) {
        size_t f = kC * kH + kH + kH * 4 + 4 + kWarps * kC * 33 + TR * 3 * 2 + 3 * TR * np(N);
        return (f + 3) & ~(size_t)3;
    }
    static __host__ __device__ size_t bytes(int N) {
        return floats_before_out(N) * 4 + 2 * (size_t)TR * (N + 1) * 16 + (((size_t)TR * N + 15) & ~(size_t)15);
    }

    __device__ void carve(unsigned char* raw, int N_) {
        N = N_; NP = np(N_);
        w1s = reinterpret_cast<float*>(raw);
        b1s = w1s + kC * kH;
        w2s = b1s + kH;
        b2s = w2s + kH * 4;
        feat = b2s + 4;
        ro = feat + kWarps * kC * 33;
        rd = ro + TR * 3;
        s_co = rd + TR * 3;
        cdf = s_co + TR * NP;
        s_fi = cdf + TR * NP;
        out_co = reinterpret_cast<float4*>(w1s + floats_before_out(N_));
        out_fi = out_co + TR * (N + 1);
        fperm = reinterpret_cast<unsigned char*>(out_fi + TR * (N + 1));
    }
};

// Stages the MLP parameters with the reference's runtime gains (layers.py:39,47).  Needs a __syncthreads() afterwards.
template <int TR>
__device__ __forceinline__ void stage_mlp(const Block<TR>& s, const Params& p) {
    const int tid = threadIdx.x;
    const float g1 = rsqrtf((float)kC), g2 = rsqrtf((float)kH);
    for (int t = tid; t < kC * kH; t += kThreads) { int c = t / kH, j = t - c * kH; s.w1s[t] = p.w1[j * kC + c] * g1; }
    for (int t = tid; t < kH; t += kThreads) s.b1s[t] = p.b1[t];
    for (int t = tid; t < kH * 4; t += kThreads) { int j = t >> 2, k = t & 3; s.w2s[t] = p.w2[k * kH + j] * g2; }
    if (tid < 4) s.b2s[tid] = p.b2[tid];
}

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

// Hidden layer pre-activations (bias included) with lane == sample, fp32 SIMT (layers.py:53-57).
__device__ __forceinline__ void mlp_hidden(float (&h)[kH], const float* featw, const float* w1s, const float* b1s, int lane) {
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
#pragma unroll
    for (int j = 0; j < kH; j++) h[j] += b1s[j];
}

// TriPlaneMLP (networks_epigraf.py:55): h = lrelu_0.2(pre) * sqrt(2);  out = h @ (W2/sqrt(64))^T + b2.
__device__ __forceinline__ float4 mlp_output(const float (&h)[kH], const float* w2s, const float* b2s) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const float sqrt2 = 1.4142135623730951f;
#pragma unroll
    for (int j = 0; j < kH; j++) {
        const float v = (h[j] > 0.f ? h[j] : h[j] * 0.2f) * sqrt2;
        const float4 w = reinterpret_cast<const float4*>(w2s)[j];
        o.x = fmaf(w.x, v, o.x); o.y = fmaf(w.y, v, o.y); o.z = fmaf(w.z, v, o.z); o.w = fmaf(w.w, v, o.w);
    }
    o.x += b2s[0]; o.y += b2s[1]; o.z += b2s[2]; o.w += b2s[3];
    return o;
}

// Footprint of sample (ray rl, depth s) -- zero footprint when !valid.
template <int TR>
__device__ __forceinline__ void footprint_of(Footprint& fp, const Block<TR>& s, const Params& p, int rl, float sd, bool valid) {
#pragma unroll
    for (int k = 0; k < 3; k++) { fp.base[k] = 0; fp.wxa[k] = fp.wxb[k] = fp.wya[k] = fp.wyb[k] = 0.f; }
    if (valid) {
        const float t = s_to_t(sd, p.o.ray_start, p.o.ray_end);
        const float px = (s.ro[rl * 3 + 0] + t * s.rd[rl * 3 + 0]) / p.o.box_half;
        const float py = (s.ro[rl * 3 + 1] + t * s.rd[rl * 3 + 1]) / p.o.box_half;
        const float pz = (s.ro[rl * 3 + 2] + t * s.rd[rl * 3 + 2]) / p.o.box_half;
        sample_footprint(fp, px, py, pz, p);
    }
}

// Passes A, B, C.  On return (after the trailing __syncthreads) s.out_co / s.out_fi / s.s_co / s.s_fi are complete.
template <class PT, int TR>
__device__ __forceinline__ void forward_passes(const Block<TR>& s, const Params& p, const PT* img, int64_t ray_base, int nrays) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = s.N, NP = s.NP;
    float* featw = s.feat + warp * kC * 33;
    const int total = TR * N;
    for (int pass = 0; pass < 2; pass++) {
        float4* outp = pass ? s.out_fi : s.out_co;
        for (int s0 = 0; s0 < total; s0 += kThreads) {
            const int si = s0 + tid;
            const int rl = si / N, i = si - rl * N;
            const bool valid = (si < total) && (rl < nrays);
            float sd = 0.f;
            if (valid) {
                if (pass == 0) {
                    const float u = p.u_coarse ? p.u_coarse[(ray_base + rl) * N + i]
                                               : rng_uniform(p.o, (uint64_t)(ray_base + rl), i, 0);
                    sd = coarse_s(i, N, u);
                    s.s_co[rl * NP + i] = sd;
                } else {
                    sd = s.s_fi[rl * NP + i];
                }
            }
            Footprint fp;
            footprint_of<TR>(fp, s, p, rl, sd, valid);
            gather_features<PT>(img, fp, featw, p.psX, p.psY, lane);
            float h[kH];
            mlp_hidden(h, featw, s.w1s, s.b1s, lane);
            float4 o = mlp_output(h, s.w2s, s.b2s);
            __syncwarp();
            if (valid) {
                if (p.o.noise_std > 0.f) {
                    // the reference draws the noise in sample-generation order (tri_plane_renderer.py:186); fine samples were
                    // sorted afterwards here, so map back through fperm
                    const float* sn = pass ? p.sn_fine : p.sn_coarse;
                    const int ni = pass ? (int)s.fperm[rl * N + i] : i;
                    const float z = sn ? sn[(ray_base + rl) * N + ni] : rng_normal(p.o, (uint64_t)(ray_base + rl), ni, 2 + pass);
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
                const float* sc = s.s_co + rl * NP;
                float* cd = s.cdf + rl * NP;
                float* sf = s.s_fi + rl * NP;
                unsigned char* pm = s.fperm + rl * N;
                float T = 1.f;
                for (int i = 0; i < N; i++) {
                    const float sig = density_act(s.out_co[rl * (N + 1) + i].w, p.o.clamp_mode);
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
                    int j = k;                         // insertion into the sorted prefix sf[0..k)
                    while (j > 0 && sf[j - 1] > v) { sf[j] = sf[j - 1]; pm[j] = pm[j - 1]; j--; }
                    sf[j] = v; pm[j] = (unsigned char)k;
                }
            }
            __syncthreads();
        }
    }
}

// Depth-merged iteration over the 2N samples of one ray (coarse wins ties; both lists ascending).
template <int TR>
struct Merge {
    const float* sc; const float* sf; const float4* oc; const float4* of; int N; float t0, t1; int ic, jf;
    __device__ Merge(const Block<TR>& s, const Params& p, int rl)
        : sc(s.s_co + rl * s.NP), sf(s.s_fi + rl * s.NP), oc(s.out_co + rl * (s.N + 1)), of(s.out_fi + rl * (s.N + 1)),
          N(s.N), t0(p.o.ray_start), t1(p.o.ray_end), ic(0), jf(0) {}
    // returns the source code of the popped sample: index i for coarse, N + j for fine
    __device__ int pop(float& t, float4& v) {
        const float tc = (ic < N) ? s_to_t(sc[ic], t0, t1) : 0.f;
        const float tf = (jf < N) ? s_to_t(sf[jf], t0, t1) : 0.f;
        const bool take_c = (jf >= N) || (ic < N && tc <= tf);
        if (take_c) { t = tc; v = oc[ic]; return ic++; }
        t = tf; v = of[jf]; return N + jf++;
    }
};

}  // namespace rm
