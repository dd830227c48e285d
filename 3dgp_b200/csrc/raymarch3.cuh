// Fused tri-plane ray-march, third generation (sm_100a) -- forward pipeline.  Same math and the same Philox stream as raymarch2.cuh (the backward
// kernels regenerate identical variates), re-cut for instruction count, occupancy and L1 reuse:
//   * persistent 256-thread CTAs (2 per SM, 16 warps / SM): the MLP B fragments are staged ONCE per CTA, not once per 8 rays;
//   * 16-ray tiles that are 4 x 4 PIXEL tiles when the ray grid is known (in-kernel ray generation, or img_w / img_h given), walked sample-major
//     (a warp = 2 depth indices x 16 rays): rays whose samples share texels along a plane's collapsed axis sit in the same CTA at the same time,
//     so the second hit is an L1 hit instead of an L2 round trip;
//   * direct-to-fragment gather: thread (g, t) of the mma quad layout loads 8 consecutive channels (one 256-bit load, LDG.E.256) of the 12 taps
//     of ITS two sample rows and blends them in registers -- the blended values ARE the layer-1 A fragments (the contraction index is permuted
//     consistently in the staged W1 fragments).  No feature tile in shared memory, no STS / LDSM round trip, half the load instructions;
//   * the mean over the three planes (1/3) and the lrelu gain (sqrt 2) are folded into the staged weights;
//   * both 16-row m-tiles of a warp share every W1 B-fragment load (half the shared-memory weight traffic);
//   * density noise and stratification jitter are drawn lane-parallel (one Philox call per sample instead of four per quad lane);
//   * optional in-kernel ray generation from the camera (cam2world, fov, patch transform): no ray tensors in HBM on the forward path.
#pragma once
#include "raymarch2.cuh"

namespace rm3 {
using namespace rm;
using rm2::mma_tf32;
using rm2::split_tf32;
using rm2::to_tf32;

constexpr int TRAYS = 16;          // rays per tile
constexpr int kT3 = 256;           // threads per CTA
constexpr int kW3 = 8;             // warps per CTA
constexpr int FPS = 20;            // footprint record stride (words): 8 records x LDS.128 are bank-conflict free

struct Smem3 {
    float2* w1h; float2* w1l;      // [4 ksteps][8 ntiles][32 lanes]   B fragments of W1 * g1 * sqrt2 / 3, channel-permuted (hi / lo TF32 parts)
    float2* w2h; float2* w2l;      // [8 ksteps][32 lanes]             B fragments of W2 * g2 (rows permuted to match layer-1 C fragments)
    float* b1s; float* b2s;        // [64] (x sqrt2), [4]
    uint32_t* fpr;                 // [kW3][32][FPS]   footprint records: words 0..2 texel bases, 4..15 tap weights
    float* nzs;                    // [kW3][32]        density noise of the warp's samples
    float* ro; float* rd;          // [TRAYS][3]
    int* rid;                      // [TRAYS]          ray index inside the image (-1: outside)
    float* s_co;                   // [TRAYS][NP]  coarse depths (s-space)
    float* bufA;                   // [TRAYS][NP]  alpha -> w' -> pdf/cdf -> (after sort) sorted fine depths
    float* bufB;                   // [TRAYS][NP]  unsorted fine depths
    float* wm;                     // [TRAYS][2N+1] merged alpha -> weights
    float4* out_co; float4* out_fi;  // [TRAYS][N+1]
    unsigned char* fperm;          // [TRAYS][N]
    unsigned char* ord;            // [TRAYS][2N]
    int N, NP;
    static __host__ __device__ int np(int N) { return N | 1; }
    static __host__ __device__ size_t bytes(int N) {
        size_t b = (size_t)(2 * 4 * 8 * 32 + 2 * 8 * 32) * 8 + (64 + 4) * 4;
        b += (size_t)kW3 * 32 * FPS * 4 + (size_t)kW3 * 32 * 4;
        b += (size_t)TRAYS * 7 * 4 + 3 * (size_t)TRAYS * np(N) * 4 + (size_t)TRAYS * (2 * N + 1) * 4;
        b = (b + 15) & ~(size_t)15;
        b += 2 * (size_t)TRAYS * (N + 1) * 16;
        b += (((size_t)TRAYS * N + 15) & ~(size_t)15) + (((size_t)TRAYS * 2 * N + 15) & ~(size_t)15);
        return b + 16;
    }
    __device__ void carve(unsigned char* raw, int N_) {
        N = N_; NP = np(N_);
        w1h = reinterpret_cast<float2*>(raw); w1l = w1h + 4 * 8 * 32;
        w2h = w1l + 4 * 8 * 32; w2l = w2h + 8 * 32;
        b1s = reinterpret_cast<float*>(w2l + 8 * 32); b2s = b1s + 64;
        fpr = reinterpret_cast<uint32_t*>(b2s + 4);
        nzs = reinterpret_cast<float*>(fpr + kW3 * 32 * FPS);
        ro = nzs + kW3 * 32; rd = ro + TRAYS * 3;
        rid = reinterpret_cast<int*>(rd + TRAYS * 3);
        s_co = reinterpret_cast<float*>(rid + TRAYS); bufA = s_co + TRAYS * NP; bufB = bufA + TRAYS * NP;
        wm = bufB + TRAYS * NP;
        size_t off = (size_t)(reinterpret_cast<unsigned char*>(wm + TRAYS * (2 * N + 1)) - raw);
        off = (off + 15) & ~(size_t)15;
        out_co = reinterpret_cast<float4*>(raw + off); out_fi = out_co + TRAYS * (N + 1);
        fperm = reinterpret_cast<unsigned char*>(out_fi + TRAYS * (N + 1));
        ord = fperm + (((size_t)TRAYS * N + 15) & ~(size_t)15);
    }
};

// 8 consecutive channels of one texel through the read-only path: one 256-bit load (fp32 planes) / one 128-bit load (fp16 planes).
template <class T> __device__ __forceinline__ void ld_tex8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld_tex8<float>(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
template <> __device__ __forceinline__ void ld_tex8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// Stages W1 / W2 as mma B fragments.  W1: the reference's 1/sqrt(fan_in) gain (layers.py:39,47), the 1/3 of the plane mean (networks_epigraf.py:54) and
// the sqrt2 activation gain (lrelu is positively homogeneous) folded in; contraction index permuted for the direct gather: in k-step ks, logical
// column t <-> channel 8t + 2ks, column t + 4 <-> channel 8t + 2ks + 1 (thread (g, t) of the quad layout owns channels [8t, 8t + 8) of its rows).
__device__ __forceinline__ void stage_weights3(const Smem3& s, const Params& p) {
    const float g1 = rsqrtf((float)kC) * 1.4142135623730951f * (1.0f / 3.0f), g2 = rsqrtf((float)kH);
    for (int i = threadIdx.x; i < 4 * 8 * 32; i += kT3) {
        const int lane = i & 31, j = (i >> 5) & 7, ks = i >> 8;
        const int g = lane >> 2, t = lane & 3;
        const float v0 = p.w1[(8 * j + g) * kC + 8 * t + 2 * ks] * g1, v1 = p.w1[(8 * j + g) * kC + 8 * t + 2 * ks + 1] * g1;
        const float h0 = __uint_as_float(to_tf32(v0)), h1 = __uint_as_float(to_tf32(v1));
        s.w1h[i] = make_float2(h0, h1);
        s.w1l[i] = make_float2(__uint_as_float(to_tf32(v0 - h0)), __uint_as_float(to_tf32(v1 - h1)));
    }
    for (int i = threadIdx.x; i < 8 * 32; i += kT3) {
        const int lane = i & 31, ks = i >> 5;
        const int g = lane >> 2, t = lane & 3;                  // output column n = g (only n < 4 is real)
        const float v0 = (g < 4) ? p.w2[g * kH + 8 * ks + 2 * t] * g2 : 0.f;          // logical k = t     <-> hidden unit 8ks + 2t
        const float v1 = (g < 4) ? p.w2[g * kH + 8 * ks + 2 * t + 1] * g2 : 0.f;      // logical k = t + 4 <-> hidden unit 8ks + 2t + 1
        const float h0 = __uint_as_float(to_tf32(v0)), h1 = __uint_as_float(to_tf32(v1));
        s.w2h[i] = make_float2(h0, h1);
        s.w2l[i] = make_float2(__uint_as_float(to_tf32(v0 - h0)), __uint_as_float(to_tf32(v1 - h1)));
    }
    for (int i = threadIdx.x; i < kH; i += kT3) s.b1s[i] = p.b1[i] * 1.4142135623730951f;
    if (threadIdx.x < 4) s.b2s[threadIdx.x] = p.b2[threadIdx.x];
}

// Blends the 12 taps of the two sample rows (16 mt + g, 16 mt + g + 8) this thread owns: f[h][q] = sum_taps w * texel[8t + q]  (sum over the three
// planes; the 1/3 lives in W1).  DENSE: channel-minor planes with psX == 3 * kC, so the x-neighbour is an immediate offset.
template <class PT, bool DENSE>
__device__ __forceinline__ void gather_rows(const PT* __restrict__ img, const uint32_t* fprw, int mt, int g, int t, int64_t psX, int64_t psY, float (&f)[2][8]) {
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int q = 0; q < 8; q++) f[h][q] = 0.f;
    uint32_t bases[2][3];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint4 b = *reinterpret_cast<const uint4*>(fprw + (mt * 16 + g + 8 * h) * FPS);
        bases[h][0] = b.x; bases[h][1] = b.y; bases[h][2] = b.z;
    }
    const int64_t sx = DENSE ? (int64_t)(3 * kC) : psX;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float v[2][4][8];
        float w[2][4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const PT* tp = img + bases[h][k] + 8 * t;
            ld_tex8<PT>(tp, v[h][0]); ld_tex8<PT>(tp + sx, v[h][1]);
            ld_tex8<PT>(tp + psY, v[h][2]); ld_tex8<PT>(tp + psY + sx, v[h][3]);
            const uint4 wq = *reinterpret_cast<const uint4*>(fprw + (mt * 16 + g + 8 * h) * FPS + 4 + 4 * k);
            w[h][0] = __uint_as_float(wq.x); w[h][1] = __uint_as_float(wq.y); w[h][2] = __uint_as_float(wq.z); w[h][3] = __uint_as_float(wq.w);
        }
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int tp = 0; tp < 4; tp++)
#pragma unroll
                for (int q = 0; q < 8; q++) f[h][q] = fmaf(w[h][tp], v[h][tp][q], f[h][q]);
    }
}

// Two-layer MLP of the warp's 32 samples.  fa[mt][h][q]: blended channel 8t + q of row 16 mt + g + 8 h (layer-1 A fragments after the TF32 split).
// Result o[mt][0..3]: lane (g, t) with t == 0 holds (r, g) and t == 1 holds (b, sigma) of rows 16 mt + g (o[mt][0..1]) and 16 mt + g + 8 (o[mt][2..3]).
template <int MODE>
__device__ __forceinline__ void mlp_mma3(const Smem3& s, const float (&fa)[2][2][8], int lane, float (&o)[2][4]) {
    const int t = lane & 3;
    float c[2][8][4];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float2 bb = *reinterpret_cast<const float2*>(s.b1s + 8 * j + 2 * t);
#pragma unroll
        for (int mt = 0; mt < 2; mt++) { c[mt][j][0] = bb.x; c[mt][j][1] = bb.y; c[mt][j][2] = bb.x; c[mt][j][3] = bb.y; }
    }
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            split_tf32<MODE>(fa[mt][0][2 * ks], ah[mt][0], al[mt][0]); split_tf32<MODE>(fa[mt][1][2 * ks], ah[mt][1], al[mt][1]);
            split_tf32<MODE>(fa[mt][0][2 * ks + 1], ah[mt][2], al[mt][2]); split_tf32<MODE>(fa[mt][1][2 * ks + 1], ah[mt][3], al[mt][3]);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float2 bh = s.w1h[(ks * 8 + j) * 32 + lane];
            float2 bl = make_float2(0.f, 0.f);
            if (MODE == 2) bl = s.w1l[(ks * 8 + j) * 32 + lane];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                mma_tf32(c[mt][j], ah[mt], __float_as_uint(bh.x), __float_as_uint(bh.y));
                if (MODE == 2) {
                    mma_tf32(c[mt][j], ah[mt], __float_as_uint(bl.x), __float_as_uint(bl.y));
                    mma_tf32(c[mt][j], al[mt], __float_as_uint(bh.x), __float_as_uint(bh.y));
                }
            }
        }
    }
    float d[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        d[mt][0] = (t == 0) ? s.b2s[0] : (t == 1) ? s.b2s[2] : 0.f;
        d[mt][1] = (t == 0) ? s.b2s[1] : (t == 1) ? s.b2s[3] : 0.f;
        d[mt][2] = d[mt][0]; d[mt][3] = d[mt][1];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float2 bh = s.w2h[j * 32 + lane];
        float2 bl = make_float2(0.f, 0.f);
        if (MODE == 2) bl = s.w2l[j * 32 + lane];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
            uint32_t ah[4], al[4];
            // lrelu(x) = max(x, 0.2 x); the sqrt2 gain already sits in W1 / b1.  Layer-1 C fragment (rows g / g+8, hidden 8j+2t / +1) -> layer-2 A fragment
            const float h0 = fmaxf(c[mt][j][0], 0.2f * c[mt][j][0]), h1 = fmaxf(c[mt][j][1], 0.2f * c[mt][j][1]);
            const float h2 = fmaxf(c[mt][j][2], 0.2f * c[mt][j][2]), h3 = fmaxf(c[mt][j][3], 0.2f * c[mt][j][3]);
            split_tf32<MODE>(h0, ah[0], al[0]); split_tf32<MODE>(h2, ah[1], al[1]);
            split_tf32<MODE>(h1, ah[2], al[2]); split_tf32<MODE>(h3, ah[3], al[3]);
            mma_tf32(d[mt], ah, __float_as_uint(bh.x), __float_as_uint(bh.y));
            if (MODE == 2) {
                mma_tf32(d[mt], ah, __float_as_uint(bl.x), __float_as_uint(bl.y));
                mma_tf32(d[mt], al, __float_as_uint(bh.x), __float_as_uint(bh.y));
            }
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int i = 0; i < 4; i++) o[mt][i] = d[mt][i];
}

// Footprint record of this lane's sample (same content as rm2::stage_footprint).
__device__ __forceinline__ void stage_footprint3(uint32_t* rec, const Params& p, const float* ro, const float* rd, float sd, bool valid) {
    rm2::stage_footprint(rec, p, ro, rd, sd, valid);
}

// Passes A-C (coarse march, importance sampling, fine march) and D1-D3 (depth merge, alpha, transmittance scan) for the tile's rays.
// Ray rl of the tile is image ray s.rid[rl] (< 0: outside the image).  On return (after the trailing __syncthreads): s.s_co / s.bufA = coarse / sorted
// fine depths (s-space), s.out_co / s.out_fi = (r, g, b, sigma) per sample, s.fperm = sort permutation, s.ord = merged order, s.wm[m] = w_m,
// s.wm[2N] = final transmittance.
template <class PT, int MODE, bool DENSE>
__device__ __forceinline__ void forward_phases3(const Smem3& s, const Params& p, const PT* __restrict__ img, int64_t img_ray_base) {
    const int N = p.o.N, NP = s.NP, M2 = 2 * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    uint32_t* fprw = s.fpr + warp * 32 * FPS;
    float* nzw = s.nzs + warp * 32;
    const float t0 = p.o.ray_start, t1 = p.o.ray_end;
    const float big = p.o.use_inf_depth ? 1e10f : 1e-3f;
    const int total = TRAYS * N;                   // sample-major: element e <-> depth index e >> 4, ray e & 15
    const int rl_own = tid & (TRAYS - 1);          // every strided loop below keeps this thread on one ray
    const int rid_own = s.rid[rl_own];
    float* s_fi = s.bufA;     // sorted fine depths live in bufA after the sort
    const bool noisy = p.o.noise_std > 0.f;
    for (int pass = 0; pass < 2; pass++) {
        float4* outp = pass ? s.out_fi : s.out_co;
        for (int e0 = 0; e0 < total; e0 += kT3) {
            const int e = e0 + tid;
            const int i = e >> 4;
            const bool valid = (e < total) && (rid_own >= 0);
            float sd = 0.f, nz = 0.f;
            if (valid) {
                const int64_t gr = img_ray_base + rid_own;
                if (pass == 0) {
                    const float u = p.u_coarse ? p.u_coarse[gr * N + i] : rng_uniform(p.o, (uint64_t)gr, i, 0);
                    sd = coarse_s(i, N, u);
                    s.s_co[rl_own * NP + i] = sd;
                } else {
                    sd = s_fi[rl_own * NP + i];
                }
                if (noisy) {
                    const float* sn = pass ? p.sn_fine : p.sn_coarse;
                    const int ni = pass ? (int)s.fperm[rl_own * N + i] : i;
                    nz = (sn ? sn[gr * N + ni] : rng_normal(p.o, (uint64_t)gr, ni, 2 + pass)) * p.o.noise_std;
                }
            }
            stage_footprint3(fprw + lane * FPS, p, s.ro + rl_own * 3, s.rd + rl_own * 3, sd, valid);
            nzw[lane] = nz;
            __syncwarp();
            float fa[2][2][8];
            gather_rows<PT, DENSE>(img, fprw, 0, g, t, p.psX, p.psY, fa[0]);
            gather_rows<PT, DENSE>(img, fprw, 1, g, t, p.psX, p.psY, fa[1]);
            float o[2][4];
            mlp_mma3<MODE>(s, fa, lane, o);
            // lane (g,t): t == 0 -> (r,g), t == 1 -> (b,sigma) of warp samples mt*16+g and mt*16+g+8
            if (t < 2) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int ws = mt * 16 + g + hh * 8;                     // sample index within the warp
                        const int ce = e0 + warp * 32 + ws;                      // element index within the pass
                        const int rl2 = ce & (TRAYS - 1), i2 = ce >> 4;
                        if (ce < total && s.rid[rl2] >= 0) {
                            float v0 = o[mt][2 * hh], v1 = o[mt][2 * hh + 1];
                            if (t == 1) v1 += nzw[ws];
                            float2* dst = reinterpret_cast<float2*>(&outp[rl2 * (N + 1) + i2]) + t;
                            *dst = make_float2(v0, v1);
                        }
                    }
                }
            }
            __syncwarp();
        }
        __syncthreads();

        if (pass == 0) {
            // ---- B: importance sampling (tri_plane_renderer.py:152-153, 237-295, 353-383); thread <-> (ray tid & 15, samples tid >> 4 + 16 n)
            if (rid_own >= 0) {
                for (int i = tid >> 4; i < N; i += kT3 / TRAYS) {           // B1: alpha_i (s-space deltas)
                    const float* sc = s.s_co + rl_own * NP;
                    const float delta = (i < N - 1) ? sc[i + 1] - sc[i] : big;
                    s.bufA[rl_own * NP + i] = 1.f - expf(-delta * density_act(s.out_co[rl_own * (N + 1) + i].w, p.o.clamp_mode));
                }
            }
            __syncthreads();
            if (tid < TRAYS && s.rid[tid] >= 0) {                           // B2: transmittance scan -> w' = (alpha T + 1e-5) + 1e-5, sum
                float* a = s.bufA + tid * NP;
                float T = 1.f, sum = 0.f;
                for (int i = 0; i < N; i++) {
                    const float al = a[i];
                    const float w = (al * T + 1e-5f) + 1e-5f;
                    T *= (1.f - al + 1e-10f);
                    a[i] = w;
                    if (i >= 1 && i <= N - 2) sum += w;
                }
                // B3 + B4 fused: cdf = [0, cumsum(w'_k / sum)], k = 1..N-2  (N-1 entries)
                float run = 0.f;
                a[0] = 0.f;
                for (int k = 1; k <= N - 2; k++) { run += a[k] / sum; a[k] = run; }
            }
            __syncthreads();
            if (rid_own >= 0) {
                const int64_t gr = img_ray_base + rid_own;
                const float* cd = s.bufA + rl_own * NP; const float* sc = s.s_co + rl_own * NP;
                for (int k = tid >> 4; k < N; k += kT3 / TRAYS) {           // B5: inverse CDF -> unsorted fine depths
                    const float u = p.u_fine ? p.u_fine[gr * N + k] : rng_uniform(p.o, (uint64_t)gr, k, 1);
                    int lo = 0, hi = N - 1;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (cd[mid] <= u) lo = mid + 1; else hi = mid; }
                    const int below = max(lo - 1, 0), above = min(lo, N - 2);
                    const float c0 = cd[below], c1 = cd[above];
                    float den = c1 - c0;
                    if (den < 1e-5f) den = 1.f;
                    const float b0 = 0.5f * (sc[below] + sc[below + 1]), b1v = 0.5f * (sc[above] + sc[above + 1]);
                    s.bufB[rl_own * NP + k] = b0 + (u - c0) / den * (b1v - b0);
                }
            }
            __syncthreads();
            if (rid_own >= 0) {
                const float* v = s.bufB + rl_own * NP;
                for (int k = tid >> 4; k < N; k += kT3 / TRAYS) {           // B6: rank sort (stable) into bufA
                    const float vk = v[k];
                    int rank = 0;
                    for (int j = 0; j < N; j++) rank += (v[j] < vk || (v[j] == vk && j < k)) ? 1 : 0;
                    s.bufA[rl_own * NP + rank] = vk;
                    s.fperm[rl_own * N + rank] = (unsigned char)k;
                }
            }
            __syncthreads();
        }
    }

    // ---- D: depth merge + compositing in t-space (tri_plane_renderer.py:163-166, 196-206, 353-405)
    if (rid_own >= 0) {
        const float* sc = s.s_co + rl_own * NP; const float* sf = s_fi + rl_own * NP;
        for (int m = tid >> 4; m < M2; m += kT3 / TRAYS) {                  // D1: merged position of every sample (coarse wins ties)
            int pos;
            if (m < N) {
                const float tc = s_to_t(sc[m], t0, t1);
                int lo = 0, hi = N;                                         // #{j : tf_j < tc}
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_to_t(sf[mid], t0, t1) < tc) lo = mid + 1; else hi = mid; }
                pos = m + lo;
            } else {
                const float tf = s_to_t(sf[m - N], t0, t1);
                int lo = 0, hi = N;                                         // #{i : tc_i <= tf}
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_to_t(sc[mid], t0, t1) <= tf) lo = mid + 1; else hi = mid; }
                pos = (m - N) + lo;
            }
            s.ord[rl_own * M2 + pos] = (unsigned char)m;
        }
    }
    __syncthreads();
    if (rid_own >= 0) {
        const int rl = rl_own;
        auto depth_of = [&](int code) { return s_to_t(code < N ? s.s_co[rl * NP + code] : s_fi[rl * NP + code - N], t0, t1); };
        for (int m = tid >> 4; m < M2; m += kT3 / TRAYS) {                  // D2: alpha of every merged sample
            const int code = s.ord[rl * M2 + m];
            const float tm = depth_of(code);
            const float delta = (m == M2 - 1) ? big : depth_of(s.ord[rl * M2 + m + 1]) - tm;
            const float sig = code < N ? s.out_co[rl * (N + 1) + code].w : s.out_fi[rl * (N + 1) + code - N].w;
            s.wm[rl * (M2 + 1) + m] = 1.f - expf(-delta * density_act(sig, p.o.clamp_mode));
        }
    }
    __syncthreads();
    if (tid < TRAYS && s.rid[tid] >= 0) {                                   // D3: transmittance scan -> weights
        float* a = s.wm + tid * (M2 + 1);
        float T = 1.f;
        for (int m = 0; m < M2; m++) { const float al = a[m]; a[m] = al * T; T *= (1.f - al + 1e-10f); }
        a[M2] = T;
    }
    __syncthreads();
}

}  // namespace rm3
