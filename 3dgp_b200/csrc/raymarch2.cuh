// Fused tri-plane ray-march, second generation (sm_100a), shared by raymarch_fwd2.cu and raymarch_bwd2.cu: same math as raymarch_fwd.cu, reorganised for throughput.
//   * tri-plane MLP on the tensor cores: mma.sync.m16n8k8 TF32 (opts.mlp_mode == 1) or error-compensated 3xTF32 (mlp_mode == 2,
//     a = a_hi + a_lo, w = w_hi + w_lo, a_hi*w_hi + a_hi*w_lo + a_lo*w_hi, fp32 accumulate);  A fragments come straight from the
//     gathered feature tile with ldmatrix, the hidden layer never leaves registers (layer-1 C fragments are layer-2 A fragments
//     after a fixed permutation of W2's rows);
//   * footprints (3 texel bases + 12 tap weights) are staged once per sample in shared memory and fetched with 4 x LDS.128 per
//     gather round instead of 15 shuffles + 12 multiplies;
//   * the per-ray phases are parallel: alpha / softplus / exp / inverse-CDF / rank-sort / depth-merge run on all 128 threads,
//     only the two transmittance scans (one multiply per sample) are serial per ray;
//   * 8 rays per CTA: ~60 KB of shared memory, 3 CTAs / SM.
#pragma once
#include "raymarch_common.cuh"

namespace rm2 {
using namespace rm;

constexpr int TR = 8;              // rays per CTA
constexpr int FSTR = 36;           // feature tile row stride (floats): 144 B rows -> conflict-free ldmatrix
constexpr int FPSTR = 20;          // footprint record stride (words)

struct Smem {
    float2* w1h; float2* w1l;      // [4 ksteps][8 ntiles][32 lanes]   B fragments of W1 (hi / lo TF32 parts)
    float2* w2h; float2* w2l;      // [8 ksteps][32 lanes]             B fragments of W2 (rows permuted to match layer-1 C fragments)
    float* b1s; float* b2s;        // [64], [4]
    float* feat;                   // [kWarps][32][FSTR]
    uint32_t* fpr;                 // [kWarps][32][FPSTR]   footprint records
    float* ro; float* rd;          // [TR][3]
    float* s_co;                   // [TR][NP]  coarse depths (s-space)
    float* bufA;                   // [TR][NP]  alpha -> w' -> pdf/cdf -> (after sort) sorted fine depths
    float* bufB;                   // [TR][NP]  unsorted fine depths
    float* wm;                     // [TR][2N+1] merged alpha -> weights
    float4* out_co; float4* out_fi;  // [TR][N+1]
    unsigned char* fperm;          // [TR][N]
    unsigned char* ord;            // [TR][2N]
    int N, NP;
    static __host__ __device__ int np(int N) { return N | 1; }
    static __host__ __device__ size_t bytes(int N) {
        size_t b = (size_t)(2 * 4 * 8 * 32 + 2 * 8 * 32) * 8 + (64 + 4) * 4;
        b += (size_t)kWarps * 32 * FSTR * 4 + (size_t)kWarps * 32 * FPSTR * 4;
        b += (size_t)TR * 6 * 4 + 3 * (size_t)TR * np(N) * 4 + (size_t)TR * (2 * N + 1) * 4;
        b = (b + 15) & ~(size_t)15;
        b += 2 * (size_t)TR * (N + 1) * 16;
        b += (((size_t)TR * N + 15) & ~(size_t)15) + (((size_t)TR * 2 * N + 15) & ~(size_t)15);
        return b + 16;
    }
    __device__ void carve(unsigned char* raw, int N_) {
        N = N_; NP = np(N_);
        w1h = reinterpret_cast<float2*>(raw); w1l = w1h + 4 * 8 * 32;
        w2h = w1l + 4 * 8 * 32; w2l = w2h + 8 * 32;
        b1s = reinterpret_cast<float*>(w2l + 8 * 32); b2s = b1s + 64;
        feat = b2s + 4;
        fpr = reinterpret_cast<uint32_t*>(feat + kWarps * 32 * FSTR);
        ro = reinterpret_cast<float*>(fpr + kWarps * 32 * FPSTR); rd = ro + TR * 3;
        s_co = rd + TR * 3; bufA = s_co + TR * NP; bufB = bufA + TR * NP;
        wm = bufB + TR * NP;
        // offsets, not integer-cast pointers: keeps the shared address space visible to the compiler (LDS/STS instead of generic LD/ST)
        size_t off = (size_t)(reinterpret_cast<unsigned char*>(wm + TR * (2 * N + 1)) - raw);
        off = (off + 15) & ~(size_t)15;
        out_co = reinterpret_cast<float4*>(raw + off); out_fi = out_co + TR * (N + 1);
        fperm = reinterpret_cast<unsigned char*>(out_fi + TR * (N + 1));
        ord = fperm + (((size_t)TR * N + 15) & ~(size_t)15);
    }
};

__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }

// x = hi + lo with hi, lo valid TF32 bit patterns.  MODE 2 (3xTF32): hi = x truncated to 10 mantissa bits, lo = the exact fp32 remainder
// truncated likewise (residual 2^-21 |x|): one LOP3 + FADD + LOP3 instead of two multi-instruction cvt.rna.  MODE 1: round-to-nearest hi only.
template <int MODE>
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    if (MODE == 2) { hi = __float_as_uint(x) & 0xffffe000u; lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u; }
    else { hi = to_tf32(x); lo = 0u; }
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}

// Stages W1 / W2 as mma B fragments (with the reference's 1/sqrt(fan_in) gains, layers.py:39,47), split into TF32 hi / lo parts.
__device__ __forceinline__ void stage_weights(const Smem& s, const Params& p) {
    const float g1 = rsqrtf((float)kC), g2 = rsqrtf((float)kH);
    for (int i = threadIdx.x; i < 4 * 8 * 32; i += kThreads) {
        const int lane = i & 31, j = (i >> 5) & 7, ks = i >> 8;
        const int g = lane >> 2, t = lane & 3;
        const float v0 = p.w1[(8 * j + g) * kC + 8 * ks + t] * g1, v1 = p.w1[(8 * j + g) * kC + 8 * ks + t + 4] * g1;
        const float h0 = __uint_as_float(to_tf32(v0)), h1 = __uint_as_float(to_tf32(v1));
        s.w1h[i] = make_float2(h0, h1);
        s.w1l[i] = make_float2(__uint_as_float(to_tf32(v0 - h0)), __uint_as_float(to_tf32(v1 - h1)));
    }
    for (int i = threadIdx.x; i < 8 * 32; i += kThreads) {
        const int lane = i & 31, ks = i >> 5;
        const int g = lane >> 2, t = lane & 3;                  // output column n = g (only n < 4 is real), k = t / t + 4
        const float v0 = (g < 4) ? p.w2[g * kH + 8 * ks + 2 * t] * g2 : 0.f;          // logical k = t     <-> hidden unit 8ks + 2t
        const float v1 = (g < 4) ? p.w2[g * kH + 8 * ks + 2 * t + 1] * g2 : 0.f;      // logical k = t + 4 <-> hidden unit 8ks + 2t + 1
        const float h0 = __uint_as_float(to_tf32(v0)), h1 = __uint_as_float(to_tf32(v1));
        s.w2h[i] = make_float2(h0, h1);
        s.w2l[i] = make_float2(__uint_as_float(to_tf32(v0 - h0)), __uint_as_float(to_tf32(v1 - h1)));
    }
    for (int i = threadIdx.x; i < kH; i += kThreads) s.b1s[i] = p.b1[i];
    if (threadIdx.x < 4) s.b2s[threadIdx.x] = p.b2[threadIdx.x];
}

// Writes the footprint record of this lane's sample: words 0..2 = texel bases, 4..15 = tap weights (plane-major: 00, 01, 10, 11).
__device__ __forceinline__ void stage_footprint(uint32_t* rec, const Params& p, const float* ro, const float* rd, float sd, bool valid) {
    uint4 q0 = make_uint4(0, 0, 0, 0);
    float w[12];
#pragma unroll
    for (int i = 0; i < 12; i++) w[i] = 0.f;
    if (valid) {
        const float t = s_to_t(sd, p.o.ray_start, p.o.ray_end);
        Footprint fp;
        sample_footprint(fp, (ro[0] + t * rd[0]) / p.o.box_half, (ro[1] + t * rd[1]) / p.o.box_half, (ro[2] + t * rd[2]) / p.o.box_half, p);
        q0 = make_uint4((uint32_t)fp.base[0], (uint32_t)fp.base[1], (uint32_t)fp.base[2], 0);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            w[4 * k + 0] = fp.wya[k] * fp.wxa[k]; w[4 * k + 1] = fp.wya[k] * fp.wxb[k];
            w[4 * k + 2] = fp.wyb[k] * fp.wxa[k]; w[4 * k + 3] = fp.wyb[k] * fp.wxb[k];
        }
    }
    uint4* r4 = reinterpret_cast<uint4*>(rec);
    r4[0] = q0;
    r4[1] = make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3]));
    r4[2] = make_uint4(__float_as_uint(w[4]), __float_as_uint(w[5]), __float_as_uint(w[6]), __float_as_uint(w[7]));
    r4[3] = make_uint4(__float_as_uint(w[8]), __float_as_uint(w[9]), __float_as_uint(w[10]), __float_as_uint(w[11]));
}

// Gathers the features of the warp's 32 samples into featw[sample][channel] (row stride FSTR), mean over the three planes.
template <class PT>
__device__ __forceinline__ void gather_tile(const PT* __restrict__ img, const uint32_t* fprw, float* featw, int64_t psX, int64_t psY, int lane) {
    const int u4 = (lane & 7) * 4, q = lane >> 3;
#pragma unroll 2
    for (int r = 0; r < 8; r++) {
        const int src = 4 * r + q;
        const uint4* rec = reinterpret_cast<const uint4*>(fprw + src * FPSTR);
        const uint4 b = rec[0];
        const uint4 wa = rec[1], wb = rec[2], wc = rec[3];
        const float w[12] = {__uint_as_float(wa.x), __uint_as_float(wa.y), __uint_as_float(wa.z), __uint_as_float(wa.w),
                             __uint_as_float(wb.x), __uint_as_float(wb.y), __uint_as_float(wb.z), __uint_as_float(wb.w),
                             __uint_as_float(wc.x), __uint_as_float(wc.y), __uint_as_float(wc.z), __uint_as_float(wc.w)};
        const uint32_t bases[3] = {b.x, b.y, b.z};
        float4 v[12];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const PT* t = img + (int)bases[k] + u4;
            v[4 * k + 0] = ld_tex4<PT>(t); v[4 * k + 1] = ld_tex4<PT>(t + psX);
            v[4 * k + 2] = ld_tex4<PT>(t + psY); v[4 * k + 3] = ld_tex4<PT>(t + psY + psX);
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 12; i++) {
            acc.x = fmaf(w[i], v[i].x, acc.x); acc.y = fmaf(w[i], v[i].y, acc.y);
            acc.z = fmaf(w[i], v[i].z, acc.z); acc.w = fmaf(w[i], v[i].w, acc.w);
        }
        const float third = 1.0f / 3.0f;
        *reinterpret_cast<float4*>(featw + src * FSTR + u4) = make_float4(acc.x * third, acc.y * third, acc.z * third, acc.w * third);
    }
    __syncwarp();
}

// Two-layer MLP of the warp's 32 samples on mma.sync TF32.  Result: lane (g,t) with t == 0 holds (r,g) and t == 1 holds (b,sigma)
// of samples mt*16 + g (o[mt][0..1]) and mt*16 + g + 8 (o[mt][2..3]).
template <int MODE>
__device__ __forceinline__ void mlp_mma(const Smem& s, const float* featw, int lane, float (&o)[2][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int lm = lane >> 3, lr = lane & 7;     // ldmatrix: matrix index / row of the address this lane provides
#pragma unroll 1
    for (int mt = 0; mt < 2; mt++) {
        float c[8][4];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float bb0 = s.b1s[8 * j + 2 * t], bb1 = s.b1s[8 * j + 2 * t + 1];
            c[j][0] = bb0; c[j][1] = bb1; c[j][2] = bb0; c[j][3] = bb1;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            uint32_t a[4], ah[4], al[4];
            ldmatrix_x4(a, featw + (mt * 16 + (lm & 1) * 8 + lr) * FSTR + ks * 8 + (lm >> 1) * 4);
#pragma unroll
            for (int i = 0; i < 4; i++) split_tf32<MODE>(__uint_as_float(a[i]), ah[i], al[i]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 bh = s.w1h[(ks * 8 + j) * 32 + lane];
                mma_tf32(c[j], ah, __float_as_uint(bh.x), __float_as_uint(bh.y));
                if (MODE == 2) {
                    const float2 bl = s.w1l[(ks * 8 + j) * 32 + lane];
                    mma_tf32(c[j], ah, __float_as_uint(bl.x), __float_as_uint(bl.y));
                    mma_tf32(c[j], al, __float_as_uint(bh.x), __float_as_uint(bh.y));
                }
            }
        }
        float d[4];
        d[0] = (t == 0) ? s.b2s[0] : (t == 1) ? s.b2s[2] : 0.f;
        d[1] = (t == 0) ? s.b2s[1] : (t == 1) ? s.b2s[3] : 0.f;
        d[2] = d[0]; d[3] = d[1];
        const float sqrt2 = 1.4142135623730951f;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t ah[4], al[4];
            // layer-1 C fragment (rows g / g+8, hidden 8j+2t / 8j+2t+1) -> layer-2 A fragment (a0,a1: k = t ; a2,a3: k = t+4)
            const float h0 = (c[j][0] > 0.f ? c[j][0] : 0.2f * c[j][0]) * sqrt2, h1 = (c[j][1] > 0.f ? c[j][1] : 0.2f * c[j][1]) * sqrt2;
            const float h2 = (c[j][2] > 0.f ? c[j][2] : 0.2f * c[j][2]) * sqrt2, h3 = (c[j][3] > 0.f ? c[j][3] : 0.2f * c[j][3]) * sqrt2;
            const float hv[4] = {h0, h2, h1, h3};
#pragma unroll
            for (int i = 0; i < 4; i++) split_tf32<MODE>(hv[i], ah[i], al[i]);
            const float2 bh = s.w2h[j * 32 + lane];
            mma_tf32(d, ah, __float_as_uint(bh.x), __float_as_uint(bh.y));
            if (MODE == 2) {
                const float2 bl = s.w2l[j * 32 + lane];
                mma_tf32(d, ah, __float_as_uint(bl.x), __float_as_uint(bl.y));
                mma_tf32(d, al, __float_as_uint(bh.x), __float_as_uint(bh.y));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) o[mt][i] = d[i];
    }
}


// Passes A-C (coarse march, importance sampling, fine march) and D1-D3 (depth merge, alpha, transmittance scan) for the CTA's rays.
// On return (after the trailing __syncthreads): s.s_co / s.bufA = coarse / sorted fine depths (s-space), s.out_co / s.out_fi =
// (r,g,b,sigma) per sample, s.fperm = sort permutation, s.ord = merged order, s.wm[m] = w_m, s.wm[2N] = final transmittance,
// and with KEEP, Tb[m] = T_m (transmittance in front of merged sample m).  s.ro / s.rd and the weights must be staged by the caller.
template <class PT, int MODE, bool KEEP>
__device__ __forceinline__ void forward_phases2(const Smem& s, const Params& p, const PT* __restrict__ img, int64_t ray_base, int nrays, float* Tb) {
    const int N = p.o.N, NP = s.NP, M2 = 2 * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* featw = s.feat + warp * 32 * FSTR;
    uint32_t* fprw = s.fpr + warp * 32 * FPSTR;
    const float t0 = p.o.ray_start, t1 = p.o.ray_end;
    const float big = p.o.use_inf_depth ? 1e10f : 1e-3f;
    const int total = TR * N;
    float* s_fi = s.bufA;     // sorted fine depths live in bufA after the sort
    for (int pass = 0; pass < 2; pass++) {
        float4* outp = pass ? s.out_fi : s.out_co;
        for (int s0 = 0; s0 < total; s0 += kThreads) {
            const int si = s0 + tid;
            const int rl = si / N, i = si - rl * N;
            const bool valid = (si < total) && (rl < nrays);
            float sd = 0.f;
            if (valid) {
                if (pass == 0) {
                    const float u = p.u_coarse ? p.u_coarse[(ray_base + rl) * N + i] : rng_uniform(p.o, (uint64_t)(ray_base + rl), i, 0);
                    sd = coarse_s(i, N, u);
                    s.s_co[rl * NP + i] = sd;
                } else {
                    sd = s_fi[rl * NP + i];
                }
            }
            stage_footprint(fprw + lane * FPSTR, p, s.ro + (valid ? rl : 0) * 3, s.rd + (valid ? rl : 0) * 3, sd, valid);
            __syncwarp();
            gather_tile<PT>(img, fprw, featw, p.psX, p.psY, lane);
            float o[2][4];
            mlp_mma<MODE>(s, featw, lane, o);
            __syncwarp();
            // lane (g,t): t == 0 -> (r,g), t == 1 -> (b,sigma) of chunk samples mt*16+g and mt*16+g+8
            const int g = lane >> 2, t = lane & 3;
            if (t < 2) {
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int cs = s0 + warp * 32 + mt * 16 + g + hh * 8;      // sample index within the CTA
                        const int rl2 = cs / N, i2 = cs - rl2 * N;
                        if (cs < total && rl2 < nrays) {
                            float v0 = o[mt][2 * hh], v1 = o[mt][2 * hh + 1];
                            if (t == 1 && p.o.noise_std > 0.f) {
                                const float* sn = pass ? p.sn_fine : p.sn_coarse;
                                const int ni = pass ? (int)s.fperm[rl2 * N + i2] : i2;
                                const float z = sn ? sn[(ray_base + rl2) * N + ni] : rng_normal(p.o, (uint64_t)(ray_base + rl2), ni, 2 + pass);
                                v1 += z * p.o.noise_std;
                            }
                            float2* dst = reinterpret_cast<float2*>(&outp[rl2 * (N + 1) + i2]) + t;
                            *dst = make_float2(v0, v1);
                        }
                    }
                }
            }
        }
        __syncthreads();

        if (pass == 0) {
            // ---- B: importance sampling (tri_plane_renderer.py:152-153, 237-295, 353-383), parallel over (ray, sample)
            for (int e = tid; e < nrays * N; e += kThreads) {               // B1: alpha_i (s-space deltas)
                const int rl = e / N, i = e - rl * N;
                const float* sc = s.s_co + rl * NP;
                const float delta = (i < N - 1) ? sc[i + 1] - sc[i] : big;
                s.bufA[rl * NP + i] = 1.f - expf(-delta * density_act(s.out_co[rl * (N + 1) + i].w, p.o.clamp_mode));
            }
            __syncthreads();
            if (tid < nrays) {                                              // B2: transmittance scan -> w' = (alpha T + 1e-5) + 1e-5, sum
                float* a = s.bufA + tid * NP;
                float T = 1.f, sum = 0.f;
                for (int i = 0; i < N; i++) {
                    const float al = a[i];
                    const float w = (al * T + 1e-5f) + 1e-5f;
                    T *= (1.f - al + 1e-10f);
                    a[i] = w;
                    if (i >= 1 && i <= N - 2) sum += w;
                }
                s.wm[tid * (M2 + 1)] = sum;                                 // scratch
            }
            __syncthreads();
            for (int e = tid; e < nrays * N; e += kThreads) {               // B3: pdf_k = w'_k / sum (k = 1..N-2)
                const int rl = e / N, i = e - rl * N;
                if (i >= 1 && i <= N - 2) s.bufA[rl * NP + i] = s.bufA[rl * NP + i] / s.wm[rl * (M2 + 1)];
            }
            __syncthreads();
            if (tid < nrays) {                                              // B4: cdf = [0, cumsum(pdf)]  (N-1 entries)
                float* a = s.bufA + tid * NP;
                float run = 0.f;
                a[0] = 0.f;
                for (int k = 1; k <= N - 2; k++) { run += a[k]; a[k] = run; }
            }
            __syncthreads();
            for (int e = tid; e < nrays * N; e += kThreads) {               // B5: inverse CDF -> unsorted fine depths
                const int rl = e / N, k = e - rl * N;
                const float* cd = s.bufA + rl * NP; const float* sc = s.s_co + rl * NP;
                const float u = p.u_fine ? p.u_fine[(ray_base + rl) * N + k] : rng_uniform(p.o, (uint64_t)(ray_base + rl), k, 1);
                int lo = 0, hi = N - 1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (cd[mid] <= u) lo = mid + 1; else hi = mid; }
                const int below = max(lo - 1, 0), above = min(lo, N - 2);
                const float c0 = cd[below], c1 = cd[above];
                float den = c1 - c0;
                if (den < 1e-5f) den = 1.f;
                const float b0 = 0.5f * (sc[below] + sc[below + 1]), b1v = 0.5f * (sc[above] + sc[above + 1]);
                s.bufB[rl * NP + k] = b0 + (u - c0) / den * (b1v - b0);
            }
            __syncthreads();
            for (int e = tid; e < nrays * N; e += kThreads) {               // B6: rank sort (stable) into bufA
                const int rl = e / N, k = e - rl * N;
                const float* v = s.bufB + rl * NP;
                const float vk = v[k];
                int rank = 0;
                for (int j = 0; j < N; j++) rank += (v[j] < vk || (v[j] == vk && j < k)) ? 1 : 0;
                s.bufA[rl * NP + rank] = vk;
                s.fperm[rl * N + rank] = (unsigned char)k;
            }
            __syncthreads();
        }
    }

    // ---- D: depth merge + compositing in t-space (tri_plane_renderer.py:163-166, 196-206, 353-405)
    for (int e = tid; e < nrays * M2; e += kThreads) {                      // D1: merged position of every sample (coarse wins ties)
        const int rl = e / M2, m = e - rl * M2;
        const float* sc = s.s_co + rl * NP; const float* sf = s_fi + rl * NP;
        int pos;
        if (m < N) {
            const float tc = s_to_t(sc[m], t0, t1);
            int lo = 0, hi = N;                                             // #{j : tf_j < tc}
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_to_t(sf[mid], t0, t1) < tc) lo = mid + 1; else hi = mid; }
            pos = m + lo;
        } else {
            const float tf = s_to_t(sf[m - N], t0, t1);
            int lo = 0, hi = N;                                             // #{i : tc_i <= tf}
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_to_t(sc[mid], t0, t1) <= tf) lo = mid + 1; else hi = mid; }
            pos = (m - N) + lo;
        }
        s.ord[rl * M2 + pos] = (unsigned char)m;
    }
    __syncthreads();
    auto depth_of = [&](int rl, int code) { return s_to_t(code < N ? s.s_co[rl * NP + code] : s_fi[rl * NP + code - N], t0, t1); };
    auto value_of = [&](int rl, int code) { return code < N ? s.out_co[rl * (N + 1) + code] : s.out_fi[rl * (N + 1) + code - N]; };
    for (int e = tid; e < nrays * M2; e += kThreads) {                      // D2: alpha of every merged sample
        const int rl = e / M2, m = e - rl * M2;
        const int code = s.ord[rl * M2 + m];
        const float tm = depth_of(rl, code);
        const float delta = (m == M2 - 1) ? big : depth_of(rl, s.ord[rl * M2 + m + 1]) - tm;
        s.wm[rl * (M2 + 1) + m] = 1.f - expf(-delta * density_act(value_of(rl, code).w, p.o.clamp_mode));
    }
    __syncthreads();
    if (tid < nrays) {                                                      // D3: transmittance scan -> weights
        float* a = s.wm + tid * (M2 + 1);
        float T = 1.f;
        if (KEEP) {                                                         // the backward also needs T_m itself
            float* tb = Tb + tid * (M2 + 1);
            for (int m = 0; m < M2; m++) { const float al = a[m]; tb[m] = T; a[m] = al * T; T *= (1.f - al + 1e-10f); }
            tb[M2] = T;
        } else {
            for (int m = 0; m < M2; m++) { const float al = a[m]; a[m] = al * T; T *= (1.f - al + 1e-10f); }
        }
        a[M2] = T;
    }
    __syncthreads();
}

}  // namespace rm2
