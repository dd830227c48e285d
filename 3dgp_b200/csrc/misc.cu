// Error plumbing + small streaming kernels of lib3dgp_b200:
//   modulate / demod_act  : the elementwise halves of modulated_conv2d (networks_stylegan2.py:67-76,142-144)
//   filtered_lrelu_act    : sign-coded leaky-ReLU used by filtered_lrelu's generic path (filtered_lrelu.cu:1105-1211)
//   grad_epilogue         : /world + nan_to_num of the flattened gradient (training_loop.py:340-341)
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void gp3d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gp3d_last_error(void) { return g_err; }
extern "C" int gp3d_version(void) { return 1; }
extern "C" int gp3d_built_arch(void) {
#ifdef GP3D_ARCH
    return GP3D_ARCH;
#else
    return 100;
#endif
}

namespace {

// ------------------------------------------------------------------------------------------------
// y[n,c,hw] = x[n,c,hw] * s[n,c]   (cl: element index = (n*HW + hw)*C + c)
template <class T, bool CL>
__global__ void __launch_bounds__(256) modulate_kernel(const T* __restrict__ x, const T* __restrict__ s, T* __restrict__ y,
                                                       int N, int C, int HW) {
    constexpr int VEC = vec16<T>::N;
    const int64_t nvec = (int64_t)N * C * HW / VEC;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = v * VEC;
        vec16<T> vx; float f[VEC];
        vx.load(x + i0); vx.unpack(f);
        if (CL) {
            const int c0 = (int)(i0 % C);
            const int n = (int)(i0 / ((int64_t)C * HW));
#pragma unroll
            for (int k = 0; k < VEC; k++) f[k] *= io_traits<T>::ld(s + (int64_t)n * C + c0 + k);
        } else {
            const float sv = io_traits<T>::ld(s + i0 / HW);
#pragma unroll
            for (int k = 0; k < VEC; k++) f[k] *= sv;
        }
        vx.pack(f); vx.store(y + i0);
    }
}

// y = clamp(act(x * d[n,c] + noise[(n),hw] + b[c]) * gain), act in {linear(1), lrelu(3)}
template <class T, bool CL, int ACT>
__global__ void __launch_bounds__(256) demod_act_kernel(const T* __restrict__ x, const T* __restrict__ d,
                                                        const T* __restrict__ noise, int noise_per_sample,
                                                        const T* __restrict__ b, T* __restrict__ y,
                                                        int N, int C, int HW, float alpha, float gain, float clamp) {
    constexpr int VEC = vec16<T>::N;
    const int64_t nvec = (int64_t)N * C * HW / VEC;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = v * VEC;
        vec16<T> vx; float f[VEC];
        vx.load(x + i0); vx.unpack(f);
        if (CL) {
            const int c0 = (int)(i0 % C);
            const int64_t pix = i0 / C;                  // n*HW + hw
            const int n = (int)(pix / HW);
            const int hw = (int)(pix - (int64_t)n * HW);
            const float nz = noise ? io_traits<T>::ld(noise + (noise_per_sample ? (int64_t)n * HW : 0) + hw) : 0.f;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const float dv = d ? io_traits<T>::ld(d + (int64_t)n * C + c0 + k) : 1.f;
                const float bv = b ? io_traits<T>::ld(b + c0 + k) : 0.f;
                f[k] = fmaf(f[k], dv, nz) + bv;
            }
        } else {
            const int64_t nc = i0 / HW;
            const int hw0 = (int)(i0 - nc * HW);
            const int n = (int)(nc / C), c = (int)(nc - (int64_t)n * C);
            const float dv = d ? io_traits<T>::ld(d + nc) : 1.f;
            const float bv = b ? io_traits<T>::ld(b + c) : 0.f;
#pragma unroll
            for (int k = 0; k < VEC; k++) {
                const float nz = noise ? io_traits<T>::ld(noise + (noise_per_sample ? (int64_t)n * HW : 0) + hw0 + k) : 0.f;
                f[k] = fmaf(f[k], dv, nz) + bv;
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            float t = f[k];
            if (ACT == 3) t = (t > 0.f) ? t : t * alpha;
            t *= gain;
            if (clamp >= 0.f) t = fminf(fmaxf(t, -clamp), clamp);
            f[k] = t;
        }
        vx.pack(f); vx.store(y + i0);
    }
}

template <class T>
__global__ void __launch_bounds__(128) flrelu_act_kernel(T* x, uint8_t* si, int N, int C, int H, int W, int sH, int sW4,
                                                         int sx, int sy, float gain, float slope, float clamp, int write_signs) {
    // one thread per group of 4 consecutive x (one sign byte); grid (ceil(W/4/128), H, N*C)
    const int xq = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    for (int nc = blockIdx.z; nc < N * C; nc += gridDim.z) {
        if (xq * 4 >= W) continue;
        T* row = x + ((int64_t)nc * H + yy) * W;
        const int sxq = (xq * 4 + sx);       // sign x coordinate of element 0 (must be a multiple of 4 for byte packing)
        const int syy = yy + sy;
        const bool sign_ok = si && syy >= 0 && syy < sH;
        uint8_t* sp = si ? si + ((int64_t)nc * sH + (sign_ok ? syy : 0)) * sW4 : nullptr;
        uint32_t code = 0;
        if (si && !write_signs) {
            // read 4 codes starting at sign-x = sxq (may straddle two bytes when sx % 4 != 0)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int sxx = sxq + k;
                uint32_t c = 0;
                if (sign_ok && sxx >= 0 && (sxx >> 2) < sW4) c = (sp[sxx >> 2] >> ((sxx & 3) * 2)) & 3u;
                code |= c << (2 * k);
            }
        }
        uint32_t wcode = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int xx = xq * 4 + k;
            if (xx >= W) break;
            float v = io_traits<T>::ld(row + xx);
            if (si && !write_signs) {
                const uint32_t c = (code >> (2 * k)) & 3u;
                v *= (c == 0) ? gain : (c == 1) ? gain * slope : 0.f;
            } else {
                uint32_t c = 0;
                if (v < 0.f) { v *= slope; c = 1; }
                v *= gain;
                if (clamp >= 0.f && fabsf(v) > clamp) { v = copysignf(clamp, v); c = 2; }   // clamped => code 2 (zero gradient)
                wcode |= c << (2 * k);
            }
            io_traits<T>::st(row + xx, v);
        }
        if (si && write_signs && sign_ok && (sxq & 3) == 0 && sxq >= 0 && (sxq >> 2) < sW4) sp[sxq >> 2] = (uint8_t)wcode;
    }
}

__global__ void __launch_bounds__(256) grad_epilogue_kernel(float* g, int64_t numel, float inv_world, float posinf, float neginf) {
    const int64_t nvec = numel / 4;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        float4 t = reinterpret_cast<float4*>(g)[v];
        float f[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float a = f[k] * inv_world;
            a = (a != a) ? 0.f : a;
            a = isinf(a) ? (a > 0.f ? posinf : neginf) : a;   // torch.nan_to_num: only nan / +-inf are replaced
            f[k] = a;
        }
        reinterpret_cast<float4*>(g)[v] = make_float4(f[0], f[1], f[2], f[3]);
    }
    if (blockIdx.x == 0) {
        const int64_t i = nvec * 4 + threadIdx.x;
        if (i < numel) {
            float a = g[i] * inv_world;
            a = (a != a) ? 0.f : a;
            g[i] = isinf(a) ? (a > 0.f ? posinf : neginf) : a;
        }
    }
}

// hi = bf16(x * s), lo = bf16(x * s - hi); channel-minor tensors, 8 channels per thread
template <class T, class OT = __nv_bfloat16>
__global__ void __launch_bounds__(256) split_bf16_kernel(const T* __restrict__ x, const float* __restrict__ s, OT* __restrict__ hi,
                                                         OT* __restrict__ lo, int N, int HW, int C, int Cp) {
    // Cp >= C: channel count of the outputs (zero-padded tail so that 96-channel tensors fill whole 64-channel TMA blocks)
    const int64_t nvec = (int64_t)N * HW * Cp / 8;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o0 = v * 8;                       // output element index
        const int64_t pix = o0 / Cp;
        const int c0 = (int)(o0 - pix * Cp);
        const int64_t i0 = pix * C + c0;                // input element index
        float f[8];
        if (c0 >= C) {
#pragma unroll
            for (int k = 0; k < 8; k++) f[k] = 0.f;
        } else if (sizeof(T) == 4) {
            const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i0), b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + i0 + 4);
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
            vec16<__half> t; t.load(reinterpret_cast<const __half*>(x) + i0); t.unpack(f);
        }
        if (s && c0 < C) {
            const int n = (int)(pix / HW);
            const float4 sa = *reinterpret_cast<const float4*>(s + (int64_t)n * C + c0), sb = *reinterpret_cast<const float4*>(s + (int64_t)n * C + c0 + 4);
            f[0] *= sa.x; f[1] *= sa.y; f[2] *= sa.z; f[3] *= sa.w; f[4] *= sb.x; f[5] *= sb.y; f[6] *= sb.z; f[7] *= sb.w;
        }
        float r[8];
        vec16<OT> h; h.pack(f); h.store(hi + o0);
        if (lo) {
            float hf[8]; h.unpack(hf);
#pragma unroll
            for (int k = 0; k < 8; k++) r[k] = f[k] - hf[k];
            vec16<OT> l; l.pack(r); l.store(lo + o0);
        }
    }
}

// Division-free form for the common channel counts (C_out / 8 divides 256): grid (pixel blocks, N); a thread keeps ONE 8-channel group
// (its styles stay in registers) and walks the pixels of image blockIdx.y, four independent pixels in flight.
template <class T, class OT = __nv_bfloat16>
__global__ void __launch_bounds__(256) split_bf16_cm_kernel(const T* __restrict__ x, const float* __restrict__ s, OT* __restrict__ hi,
                                                            OT* __restrict__ lo, int HW, int C, int Cp) {
    constexpr int U = 4;
    const int CV = Cp >> 3, PL = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int n = blockIdx.y, c0 = cv * 8;
    const bool pad = c0 >= C;
    float sv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) sv[k] = 1.f;
    if (s && !pad) {
        const float4 sa = *reinterpret_cast<const float4*>(s + (int64_t)n * C + c0), sb = *reinterpret_cast<const float4*>(s + (int64_t)n * C + c0 + 4);
        sv[0] = sa.x; sv[1] = sa.y; sv[2] = sa.z; sv[3] = sa.w; sv[4] = sb.x; sv[5] = sb.y; sv[6] = sb.z; sv[7] = sb.w;
    }
    const T* xb = x + (int64_t)n * HW * C + c0;
    OT* hb = hi + (int64_t)n * HW * Cp + c0;
    OT* lb = lo ? lo + (int64_t)n * HW * Cp + c0 : nullptr;
    for (int px0 = blockIdx.x * PL * U + pl; px0 < HW; px0 += gridDim.x * PL * U) {
        float f[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int px = px0 + u * PL;
#pragma unroll
            for (int k = 0; k < 8; k++) f[u][k] = 0.f;
            if (px < HW && !pad) {
                if (sizeof(T) == 4) {
                    const float* xp = reinterpret_cast<const float*>(xb) + (int64_t)px * C;
                    const float4 a = *reinterpret_cast<const float4*>(xp), b = *reinterpret_cast<const float4*>(xp + 4);
                    f[u][0] = a.x; f[u][1] = a.y; f[u][2] = a.z; f[u][3] = a.w; f[u][4] = b.x; f[u][5] = b.y; f[u][6] = b.z; f[u][7] = b.w;
                } else {
                    vec16<__half> t; t.load(reinterpret_cast<const __half*>(xb) + (int64_t)px * C); t.unpack(f[u]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int px = px0 + u * PL;
            if (px < HW) {
#pragma unroll
                for (int k = 0; k < 8; k++) f[u][k] *= sv[k];
                vec16<OT> h; h.pack(f[u]); h.store(hb + (int64_t)px * Cp);
                if (lb) {
                    float hf[8], r[8]; h.unpack(hf);
#pragma unroll
                    for (int k = 0; k < 8; k++) r[k] = f[u][k] - hf[k];
                    vec16<OT> l; l.pack(r); l.store(lb + (int64_t)px * Cp);
                }
            }
        }
    }
}

template <class T>
int launch_modulate(const void* x, const void* s, void* y, int N, int C, int HW, int cl, cudaStream_t st) {
    constexpr int VEC = vec16<T>::N;
    const int64_t total = (int64_t)N * C * HW;
    if (total % VEC || (!cl && HW % VEC) || (cl && C % VEC) || !gp3d_aligned16(x) || !gp3d_aligned16(y)) return GP3D_E_UNSUPPORTED;
    const int grid = gp3d_grid_for(total / VEC, 256, 8);
    if (cl) modulate_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, (const T*)s, (T*)y, N, C, HW);
    else modulate_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, (const T*)s, (T*)y, N, C, HW);
    return 0;
}

// Channel-minor float32 fast path of demod_act (C / 4 divides 256): no index divisions, demodulation / bias vectors in registers,
// four independent pixels in flight per thread.  Same arithmetic (fma order) as demod_act_kernel.
template <int ACT>
__global__ void __launch_bounds__(256) demod_act_cm_kernel(const float* __restrict__ x, const float* __restrict__ d, const float* __restrict__ noise,
                                                           int noise_per_sample, const float* __restrict__ b, float* __restrict__ y,
                                                           int C, int HW, float alpha, float gain, float clamp) {
    constexpr int U = 4;
    const int CV = C >> 2, PL = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int n = blockIdx.y;
    const float4 dv = d ? *reinterpret_cast<const float4*>(d + (int64_t)n * C + 4 * cv) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bv = b ? *reinterpret_cast<const float4*>(b + 4 * cv) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float* xb = x + (int64_t)n * HW * C + 4 * cv;
    float* yb = y + (int64_t)n * HW * C + 4 * cv;
    const float* nb = noise ? noise + (noise_per_sample ? (int64_t)n * HW : 0) : nullptr;
    for (int px0 = blockIdx.x * PL * U + pl; px0 < HW; px0 += gridDim.x * PL * U) {
        float4 v[U]; float nz[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int px = px0 + u * PL;
            nz[u] = 0.f;
            if (px < HW) { v[u] = *reinterpret_cast<const float4*>(xb + (int64_t)px * C); if (nb) nz[u] = nb[px]; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int px = px0 + u * PL;
            if (px < HW) {
                float t[4] = {fmaf(v[u].x, dv.x, nz[u]) + bv.x, fmaf(v[u].y, dv.y, nz[u]) + bv.y, fmaf(v[u].z, dv.z, nz[u]) + bv.z, fmaf(v[u].w, dv.w, nz[u]) + bv.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (ACT == 3) t[k] = (t[k] > 0.f) ? t[k] : t[k] * alpha;
                    t[k] *= gain;
                    if (clamp >= 0.f) t[k] = fminf(fmaxf(t[k], -clamp), clamp);
                }
                *reinterpret_cast<float4*>(yb + (int64_t)px * C) = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    }
}

template <class T>
int launch_demod(const void* x, const void* d, const void* noise, int nps, const void* b, void* y, int N, int C, int HW,
                 int cl, int act, float alpha, float gain, float clamp, cudaStream_t st) {
    constexpr int VEC = vec16<T>::N;
    const int64_t total = (int64_t)N * C * HW;
    if (total % VEC || (!cl && HW % VEC) || (cl && C % VEC) || !gp3d_aligned16(x) || !gp3d_aligned16(y)) return GP3D_E_UNSUPPORTED;
    const int grid = gp3d_grid_for(total / VEC, 256, 8);
    if (sizeof(T) == 4 && cl && C % 4 == 0 && (C / 4) <= 256 && 256 % (C / 4) == 0 && N <= 65535 && (!d || gp3d_aligned16(d)) && (!b || gp3d_aligned16(b))) {
        const int PL = 256 / (C / 4);
        int64_t gx = ((int64_t)HW + PL * 4 - 1) / (PL * 4);
        const int64_t cap = ((int64_t)GP3D_NUM_SMS * 8 + N - 1) / N;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        const dim3 g2((unsigned)gx, (unsigned)N);
        if (act == 3) demod_act_cm_kernel<3><<<g2, 256, 0, st>>>((const float*)x, (const float*)d, (const float*)noise, nps, (const float*)b, (float*)y, C, HW, alpha, gain, clamp);
        else demod_act_cm_kernel<1><<<g2, 256, 0, st>>>((const float*)x, (const float*)d, (const float*)noise, nps, (const float*)b, (float*)y, C, HW, alpha, gain, clamp);
        return 0;
    }
#define GP3D_DEMOD(CLV, ACTV) demod_act_kernel<T, CLV, ACTV><<<grid, 256, 0, st>>>((const T*)x, (const T*)d, (const T*)noise, nps, (const T*)b, (T*)y, N, C, HW, alpha, gain, clamp)
    if (cl) { if (act == 3) GP3D_DEMOD(true, 3); else GP3D_DEMOD(true, 1); }
    else    { if (act == 3) GP3D_DEMOD(false, 3); else GP3D_DEMOD(false, 1); }
#undef GP3D_DEMOD
    return 0;
}

}  // namespace

extern "C" int gp3d_modulate(const void* x, const void* s, void* y, int dtype, int N, int C, int HW, int cl, void* stream) {
    GP3D_CHECK_ARG(x && s && y && N >= 1 && C >= 1 && HW >= 1, "modulate: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int r = dtype == GP3D_F32 ? launch_modulate<float>(x, s, y, N, C, HW, cl, st)
          : dtype == GP3D_F16 ? launch_modulate<__half>(x, s, y, N, C, HW, cl, st)
          : dtype == GP3D_BF16 ? launch_modulate<__nv_bfloat16>(x, s, y, N, C, HW, cl, st) : GP3D_E_BADARG;
    if (r != 0) { gp3d_set_error("modulate: unsupported shape/alignment (N=%d C=%d HW=%d cl=%d)", N, C, HW, cl); return r; }
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_demod_act(const void* x, const void* d, const void* noise, int noise_per_sample, const void* b,
                              void* y, int dtype, int N, int C, int HW, int cl,
                              int act, float alpha, float gain, float clamp, void* stream) {
    GP3D_CHECK_ARG(x && y && N >= 1 && C >= 1 && HW >= 1, "demod_act: bad arguments");
    GP3D_CHECK_ARG(act == 1 || act == 3, "demod_act: only linear (1) and lrelu (3) are fused, got %d", act);
    cudaStream_t st = (cudaStream_t)stream;
    int r = dtype == GP3D_F32 ? launch_demod<float>(x, d, noise, noise_per_sample, b, y, N, C, HW, cl, act, alpha, gain, clamp, st)
          : dtype == GP3D_F16 ? launch_demod<__half>(x, d, noise, noise_per_sample, b, y, N, C, HW, cl, act, alpha, gain, clamp, st)
          : dtype == GP3D_BF16 ? launch_demod<__nv_bfloat16>(x, d, noise, noise_per_sample, b, y, N, C, HW, cl, act, alpha, gain, clamp, st)
                               : GP3D_E_BADARG;
    if (r != 0) { gp3d_set_error("demod_act: unsupported shape/alignment (N=%d C=%d HW=%d cl=%d)", N, C, HW, cl); return r; }
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_filtered_lrelu_act(void* x, uint8_t* si, int dtype, int N, int C, int H, int W,
                                       int sH, int sW4, int sx, int sy, float gain, float slope, float clamp,
                                       int write_signs, void* stream) {
    GP3D_CHECK_ARG(x && N >= 1 && C >= 1 && H >= 1 && W >= 1, "filtered_lrelu_act: bad arguments");
    GP3D_CHECK_ARG(!si || (sH >= 1 && sW4 >= 1), "filtered_lrelu_act: bad sign tensor geometry");
    GP3D_CHECK_ARG(!(si && write_signs) || (sx & 3) == 0, "filtered_lrelu_act: sign x-offset must be a multiple of 4 when writing");
    GP3D_CHECK_ARG(H <= 65535, "filtered_lrelu_act: H too large");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((W + 4 * 128 - 1) / (4 * 128), H, (unsigned)min((int64_t)N * C, (int64_t)65535));
    if (dtype == GP3D_F32) flrelu_act_kernel<float><<<grid, 128, 0, st>>>((float*)x, si, N, C, H, W, sH, sW4, sx, sy, gain, slope, clamp, write_signs);
    else if (dtype == GP3D_F16) flrelu_act_kernel<__half><<<grid, 128, 0, st>>>((__half*)x, si, N, C, H, W, sH, sW4, sx, sy, gain, slope, clamp, write_signs);
    else if (dtype == GP3D_BF16) flrelu_act_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((__nv_bfloat16*)x, si, N, C, H, W, sH, sW4, sx, sy, gain, slope, clamp, write_signs);
    else { gp3d_set_error("filtered_lrelu_act: unsupported dtype %d", dtype); return GP3D_E_BADARG; }
    GP3D_RETURN_LAUNCH();
}

// uint8 image conversion of the metrics / snapshot path (metric_utils.py:313, training_loop.py:23-49): y = uint8(clamp(x * scale + shift, 0, 255)) for the
// first Cy channels of a float32 [N, Cx, H, W] tensor with arbitrary strides -> NCHW-contiguous uint8 (float -> uint8 truncates, as torch's cast does).
namespace {
__global__ void __launch_bounds__(256) to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int N, int Cy, int HW, int W,
                                                       int64_t sN, int64_t sC, int64_t sH, int64_t sW, float scale, float shift) {
    const int64_t total = (int64_t)N * Cy * HW / 4;              // 4 consecutive pixels of one row per thread (W % 4 == 0)
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = v * 4;
        const int p = (int)(e % HW); const int64_t nc = e / HW;
        const int c = (int)(nc % Cy), n = (int)(nc / Cy);
        const int h = p / W, w = p - h * W;
        const float* src = x + n * sN + c * sC + h * sH + w * sW;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float t = __fadd_rn(__fmul_rn(src[k * sW], scale), shift);      // two roundings, as torch's img * 127.5 + 128 (an FMA differs at x.99999 boundaries)
            t = fminf(fmaxf(t, 0.f), 255.f);
            packed |= (uint32_t)(int)t << (8 * k);                // truncation toward zero
        }
        reinterpret_cast<uint32_t*>(y)[v] = packed;
    }
}
}  // namespace

extern "C" int gp3d_to_uint8(const float* x, uint8_t* y, int N, int Cx, int Cy, int H, int W, int64_t sN, int64_t sC, int64_t sH, int64_t sW,
                             float scale, float shift, void* stream) {
    GP3D_CHECK_ARG(x && y && N >= 1 && Cy >= 1 && Cy <= Cx && H >= 1 && W >= 4 && W % 4 == 0, "to_uint8: bad arguments (W must be a multiple of 4)");
    GP3D_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 3u) == 0, "to_uint8: output must be 4-byte aligned");
    const int64_t total = (int64_t)N * Cy * H * W / 4;
    to_uint8_kernel<<<gp3d_grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, y, N, Cy, H * W, W, sN, sC, sH, sW, scale, shift);
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_grad_epilogue(float* g, int64_t numel, float inv_world, float posinf, float neginf, void* stream) {
    GP3D_CHECK_ARG(g && numel >= 0, "grad_epilogue: bad arguments");
    GP3D_CHECK_ARG(gp3d_aligned16(g), "grad_epilogue: buffer must be 16-byte aligned");
    if (numel == 0) return GP3D_OK;
    grad_epilogue_kernel<<<gp3d_grid_for(numel / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(g, numel, inv_world, posinf, neginf);
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_split_pad(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, int C_out, int hi_format, void* stream) {
    GP3D_CHECK_ARG(x && hi && N >= 1 && HW >= 1 && C >= 1, "split_bf16: bad arguments");
    GP3D_CHECK_ARG(C % 8 == 0 && C_out % 8 == 0 && C_out >= C, "split_bf16: channel counts must be multiples of 8 with C_out >= C (got %d -> %d)", C, C_out);
    GP3D_CHECK_ARG(gp3d_aligned16(x) && gp3d_aligned16(hi) && (!lo || gp3d_aligned16(lo)) && (!s || gp3d_aligned16(s)), "split_bf16: pointers must be 16-byte aligned");
    GP3D_CHECK_ARG(hi_format == 0 || hi_format == 1, "split: hi_format is 0 (bf16) or 1 (fp16); either takes an optional low-order half of the same format");
    const int64_t nvec = (int64_t)N * HW * C_out / 8;
    const int grid = gp3d_grid_for(nvec, 256, 8);
    cudaStream_t st = (cudaStream_t)stream;
    GP3D_CHECK_ARG(src_dtype == GP3D_F32 || src_dtype == GP3D_F16, "split_bf16: source must be float32 or float16");
    const int CV = C_out / 8;
    if (CV <= 256 && 256 % CV == 0 && N <= 65535) {
        const int PL = 256 / CV;
        int64_t gx = ((int64_t)HW + PL * 4 - 1) / (PL * 4);
        const int64_t cap = ((int64_t)GP3D_NUM_SMS * 8 + N - 1) / N;
        if (gx > cap) gx = cap;
        if (gx < 1) gx = 1;
        const dim3 g2((unsigned)gx, (unsigned)N);
        if (hi_format == 1) {
            if (src_dtype == GP3D_F32) split_bf16_cm_kernel<float, __half><<<g2, 256, 0, st>>>((const float*)x, s, (__half*)hi, (__half*)lo, HW, C, C_out);
            else split_bf16_cm_kernel<__half, __half><<<g2, 256, 0, st>>>((const __half*)x, s, (__half*)hi, (__half*)lo, HW, C, C_out);
        } else if (src_dtype == GP3D_F32) split_bf16_cm_kernel<float><<<g2, 256, 0, st>>>((const float*)x, s, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, HW, C, C_out);
        else split_bf16_cm_kernel<__half><<<g2, 256, 0, st>>>((const __half*)x, s, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, HW, C, C_out);
        GP3D_RETURN_LAUNCH();
    }
    if (hi_format == 1) {
        if (src_dtype == GP3D_F32) split_bf16_kernel<float, __half><<<grid, 256, 0, st>>>((const float*)x, s, (__half*)hi, (__half*)lo, N, HW, C, C_out);
        else split_bf16_kernel<__half, __half><<<grid, 256, 0, st>>>((const __half*)x, s, (__half*)hi, (__half*)lo, N, HW, C, C_out);
    } else if (src_dtype == GP3D_F32) split_bf16_kernel<float><<<grid, 256, 0, st>>>((const float*)x, s, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, N, HW, C, C_out);
    else split_bf16_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, s, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, N, HW, C, C_out);
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_split_bf16_pad(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, int C_out, void* stream) {
    return gp3d_split_pad(x, src_dtype, s, hi, lo, N, HW, C, C_out, 0, stream);
}

extern "C" int gp3d_split_bf16(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, void* stream) {
    return gp3d_split_bf16_pad(x, src_dtype, s, hi, lo, N, HW, C, C, stream);
}

// ------------------------------------------------------------------------------------------------------------------------
// Backward halves of the fused modulated-conv layer (channel-minor fp32 tensors [N][HW][C], C % 4 == 0).
namespace {

// Accumulates per-thread 4-channel partial sums across the pixel lanes of a block and adds them to global memory.
__device__ __forceinline__ void block_channel_reduce_add(float4 v, float* smem4, int cv, int CV, int PL, int pl, float* dst /* [C] */) {
    // smem4: [PL][CV] float4; threads with pl >= PL (when CV does not divide the block) hold zeros and stay out
    if (pl < PL) reinterpret_cast<float4*>(smem4)[pl * CV + cv] = v;
    __syncthreads();
    if (pl == 0) {
        float4 a = v;
        for (int k = 1; k < PL; k++) {
            const float4 b = reinterpret_cast<float4*>(smem4)[k * CV + cv];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        atomicAdd(dst + 4 * cv + 0, a.x); atomicAdd(dst + 4 * cv + 1, a.y); atomicAdd(dst + 4 * cv + 2, a.z); atomicAdd(dst + 4 * cv + 3, a.w);
    }
    __syncthreads();
}

// y = clamp-free act(c * d[n,co] + noise[n?,hw] + b[co]) * gain  (forward: demod_act_kernel).  Given dy and the saved y:
//   dt = dy * gain * act'(y);  dc = dt * d;  g_b[co] += dt;  g_d[n,co] += dt * c  with c reconstructed from y;  g_ns += dt * noise.
template <int ACT>
__global__ void __launch_bounds__(256) demod_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ d,
                                                            const float* __restrict__ noise, const float* __restrict__ noise_scale,
                                                            int noise_per_sample, const float* __restrict__ b,
                                                            float* __restrict__ dc, float* __restrict__ g_d, float* __restrict__ g_b,
                                                            float* __restrict__ g_ns, int N, int HW, int C, float alpha, float gain, int chunks,
                                                            __nv_bfloat16* __restrict__ dc_hi, __nv_bfloat16* __restrict__ dc_lo, int Cp, float clamp) {
    extern __shared__ float red_smem[];
    const float nscale = (noise && noise_scale) ? *noise_scale : 1.f;
    const int CV_all = C / 4;
    const int CV = CV_all < 256 ? CV_all : 256;          // channel vectors handled per pass
    const int PL = 256 / CV;                              // pixel lanes
    const int cvl = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    float ns_acc = 0.f;
    for (int cbase = 0; cbase < CV_all; cbase += CV) {
        const int cv = cbase + cvl;
        const bool act_thread = (pl < PL) && (cv < CV_all);
        float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act_thread) {
            const float4 dv = d ? *reinterpret_cast<const float4*>(d + (int64_t)n * C + 4 * cv) : make_float4(1.f, 1.f, 1.f, 1.f);
            const float4 bv = b ? *reinterpret_cast<const float4*>(b + 4 * cv) : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int px = p0 + pl; px < p1; px += PL) {
                const int64_t off = ((int64_t)n * HW + px) * C + 4 * cv;
                const float4 g = *reinterpret_cast<const float4*>(dy + off);
                const float4 yy = *reinterpret_cast<const float4*>(y + off);
                const float nraw = noise ? noise[(noise_per_sample ? (int64_t)n * HW : 0) + px] : 0.f;   // unscaled noise image
                const float nz = nraw * nscale;
                float gy[4] = {g.x, g.y, g.z, g.w}, yv[4] = {yy.x, yy.y, yy.z, yy.w};
                const float dd[4] = {dv.x, dv.y, dv.z, dv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
                float o[4], dt[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float slope = (ACT == 3 && yv[k] < 0.f) ? alpha : 1.f;
                    dt[k] = gy[k] * gain * slope;
                    if (clamp > 0.f && !(yv[k] > -clamp && yv[k] < clamp)) dt[k] = 0.f;      // bias_act.cu: clamped outputs pass no gradient
                    const float t = yv[k] / (gain * slope);               // pre-activation: c * d + noise + b
                    const float c = (t - nz - bb[k]) / dd[k];
                    o[k] = dt[k] * dd[k];
                    ns_acc += dt[k] * nraw;
                    if (k == 0) { sb.x += dt[k]; sd.x += dt[k] * c; } else if (k == 1) { sb.y += dt[k]; sd.y += dt[k] * c; }
                    else if (k == 2) { sb.z += dt[k]; sd.z += dt[k] * c; } else { sb.w += dt[k]; sd.w += dt[k] * c; }
                }
                if (dc_hi) {   // hand the conv kernels their bf16 (hi, lo) operand pair directly, channel tail zero-padded to Cp
                    const int64_t offp = ((int64_t)n * HW + px) * Cp + 4 * cv;
                    __nv_bfloat16 h[4], l[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) { h[k] = __float2bfloat16_rn(o[k]); l[k] = __float2bfloat16_rn(o[k] - __bfloat162float(h[k])); }
                    *reinterpret_cast<uint2*>(dc_hi + offp) = *reinterpret_cast<const uint2*>(h);
                    if (dc_lo) *reinterpret_cast<uint2*>(dc_lo + offp) = *reinterpret_cast<const uint2*>(l);
                    if (4 * cv < Cp - C) {     // Cp - C <= C: the first (Cp-C)/4 channel lanes also clear the padding
                        *reinterpret_cast<uint2*>(dc_hi + ((int64_t)n * HW + px) * Cp + C + 4 * cv) = make_uint2(0u, 0u);
                        if (dc_lo) *reinterpret_cast<uint2*>(dc_lo + ((int64_t)n * HW + px) * Cp + C + 4 * cv) = make_uint2(0u, 0u);
                    }
                } else {
                    *reinterpret_cast<float4*>(dc + off) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
        if (g_b) block_channel_reduce_add(sb, red_smem, cvl, CV, PL, pl, g_b + 4 * cbase);
        if (g_d && d) block_channel_reduce_add(sd, red_smem, cvl, CV, PL, pl, g_d + (int64_t)n * C + 4 * cbase);
    }
    if (g_ns && noise) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ns_acc += __shfl_xor_sync(0xffffffffu, ns_acc, off);
        if ((threadIdx.x & 31) == 0) atomicAdd(g_ns, ns_acc);
    }
}

// dx = dxs * s[n,ci];  g_s[n,ci] += sum_hw dxs * x
__global__ void __launch_bounds__(256) modulate_bwd_kernel(const float* __restrict__ dxs, const float* __restrict__ x, const float* __restrict__ s,
                                                           float* __restrict__ dx, float* __restrict__ g_s, int N, int HW, int C, int chunks) {
    extern __shared__ float red_smem[];
    const int CV_all = C / 4;
    const int CV = CV_all < 256 ? CV_all : 256;
    const int PL = 256 / CV;
    const int cvl = threadIdx.x % CV, pl = threadIdx.x / CV;
    const int n = blockIdx.x / chunks, chunk = blockIdx.x - n * chunks;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    for (int cbase = 0; cbase < CV_all; cbase += CV) {
        const int cv = cbase + cvl;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pl < PL && cv < CV_all) {
            const float4 sv = *reinterpret_cast<const float4*>(s + (int64_t)n * C + 4 * cv);
            for (int px0 = p0 + pl; px0 < p1; px0 += 2 * PL) {      // two independent pixels in flight
                const int px1 = px0 + PL;
                const bool two = px1 < p1;
                const int64_t off0 = ((int64_t)n * HW + px0) * C + 4 * cv, off1 = ((int64_t)n * HW + px1) * C + 4 * cv;
                const float4 g0 = *reinterpret_cast<const float4*>(dxs + off0);
                const float4 x0 = *reinterpret_cast<const float4*>(x + off0);
                float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = g1;
                if (two) { g1 = *reinterpret_cast<const float4*>(dxs + off1); x1 = *reinterpret_cast<const float4*>(x + off1); }
                acc.x += g0.x * x0.x; acc.y += g0.y * x0.y; acc.z += g0.z * x0.z; acc.w += g0.w * x0.w;
                *reinterpret_cast<float4*>(dx + off0) = make_float4(g0.x * sv.x, g0.y * sv.y, g0.z * sv.z, g0.w * sv.w);
                if (two) {
                    acc.x += g1.x * x1.x; acc.y += g1.y * x1.y; acc.z += g1.z * x1.z; acc.w += g1.w * x1.w;
                    *reinterpret_cast<float4*>(dx + off1) = make_float4(g1.x * sv.x, g1.y * sv.y, g1.z * sv.z, g1.w * sv.w);
                }
            }
        }
        block_channel_reduce_add(acc, red_smem, cvl, CV, PL, pl, g_s + (int64_t)n * C + 4 * cbase);
    }
}

static int reduce_chunks(int N, int HW) {
    int chunks = (GP3D_NUM_SMS * 8 + N - 1) / N;      // 8 CTAs of 256 threads per SM: enough loads in flight for the HBM roofline
    if (chunks > HW) chunks = HW;
    if (chunks < 1) chunks = 1;
    return chunks;
}

}  // namespace

extern "C" int gp3d_demod_act_bwd_split(const float* dy, const float* y, const float* d, const float* noise, const float* noise_scale, int noise_per_sample,
                                        const float* b, float* dc, void* dc_hi, void* dc_lo, int C_pad, float* g_d, float* g_b, float* g_ns,
                                        int N, int HW, int C, int act, float alpha, float gain, void* stream) {
    GP3D_CHECK_ARG(dy && y && N >= 1 && HW >= 1 && C >= 4 && C % 4 == 0, "demod_act_bwd: bad arguments (C must be a multiple of 4)");
    GP3D_CHECK_ARG((dc != nullptr) != (dc_hi != nullptr) && (dc_hi == nullptr) == (dc_lo == nullptr), "demod_act_bwd: give either dc (float32) or the dc_hi / dc_lo bf16 pair");
    GP3D_CHECK_ARG(!dc_hi || (C_pad >= C && C_pad % 4 == 0 && C_pad - C <= C), "demod_act_bwd: padded channel count %d incompatible with C=%d", C_pad, C);
    GP3D_CHECK_ARG(act == 1 || act == 3, "demod_act_bwd: only linear (1) and lrelu (3) are fused, got %d", act);
    GP3D_CHECK_ARG(gain != 0.f, "demod_act_bwd: gain must be non-zero");
    GP3D_CHECK_ARG(gp3d_aligned16(dy) && gp3d_aligned16(y) && (!dc || gp3d_aligned16(dc)) && (!dc_hi || (gp3d_aligned16(dc_hi) && gp3d_aligned16(dc_lo)))
                   && (!d || gp3d_aligned16(d)) && (!b || gp3d_aligned16(b)), "demod_act_bwd: pointers must be 16-byte aligned");
    const int chunks = reduce_chunks(N, HW);
    const size_t smem = 256 * sizeof(float4);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* hi = (__nv_bfloat16*)dc_hi; __nv_bfloat16* lo = (__nv_bfloat16*)dc_lo;
    if (act == 3) demod_act_bwd_kernel<3><<<N * chunks, 256, smem, st>>>(dy, y, d, noise, noise_scale, noise_per_sample, b, dc, g_d, g_b, g_ns, N, HW, C, alpha, gain, chunks, hi, lo, C_pad, -1.f);
    else demod_act_bwd_kernel<1><<<N * chunks, 256, smem, st>>>(dy, y, d, noise, noise_scale, noise_per_sample, b, dc, g_d, g_b, g_ns, N, HW, C, alpha, gain, chunks, hi, lo, C_pad, -1.f);
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_act_bwd_split(const float* dy, const float* y, float* dc, void* dc_hi, void* dc_lo, int C_pad, float* g_b,
                                  int N, int HW, int C, int act, float alpha, float gain, float clamp, void* stream) {
    GP3D_CHECK_ARG(dy && y && N >= 1 && HW >= 1 && C >= 4 && C % 4 == 0, "act_bwd_split: bad arguments (C must be a multiple of 4)");
    GP3D_CHECK_ARG((dc != nullptr) != (dc_hi != nullptr) && !(dc_lo && !dc_hi), "act_bwd_split: give either dc (float32) or dc_hi (+ optional dc_lo)");
    GP3D_CHECK_ARG(!dc_hi || (C_pad >= C && C_pad % 4 == 0 && C_pad - C <= C), "act_bwd_split: padded channel count %d incompatible with C=%d", C_pad, C);
    GP3D_CHECK_ARG(act == 1 || act == 3, "act_bwd_split: only linear (1) and lrelu (3) are fused, got %d", act);
    GP3D_CHECK_ARG(gain != 0.f, "act_bwd_split: gain must be non-zero");
    GP3D_CHECK_ARG(gp3d_aligned16(dy) && gp3d_aligned16(y) && (!dc || gp3d_aligned16(dc)) && (!dc_hi || gp3d_aligned16(dc_hi)) && (!dc_lo || gp3d_aligned16(dc_lo)),
                   "act_bwd_split: pointers must be 16-byte aligned");
    const int chunks = reduce_chunks(N, HW);
    const size_t smem = 256 * sizeof(float4);
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* hi = (__nv_bfloat16*)dc_hi; __nv_bfloat16* lo = (__nv_bfloat16*)dc_lo;
    if (act == 3) demod_act_bwd_kernel<3><<<N * chunks, 256, smem, st>>>(dy, y, nullptr, nullptr, nullptr, 0, nullptr, dc, nullptr, g_b, nullptr, N, HW, C, alpha, gain, chunks, hi, lo, C_pad, clamp);
    else demod_act_bwd_kernel<1><<<N * chunks, 256, smem, st>>>(dy, y, nullptr, nullptr, nullptr, 0, nullptr, dc, nullptr, g_b, nullptr, N, HW, C, alpha, gain, chunks, hi, lo, C_pad, clamp);
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_demod_act_bwd(const float* dy, const float* y, const float* d, const float* noise, const float* noise_scale, int noise_per_sample, const float* b,
                                  float* dc, float* g_d, float* g_b, float* g_ns, int N, int HW, int C, int act, float alpha, float gain,
                                  void* stream) {
    GP3D_CHECK_ARG(dc != nullptr, "demod_act_bwd: dc is null");
    return gp3d_demod_act_bwd_split(dy, y, d, noise, noise_scale, noise_per_sample, b, dc, nullptr, nullptr, C, g_d, g_b, g_ns, N, HW, C, act, alpha, gain, stream);
}

extern "C" int gp3d_modulate_bwd(const float* dxs, const float* x, const float* s, float* dx, float* g_s, int N, int HW, int C, void* stream) {
    GP3D_CHECK_ARG(dxs && x && s && dx && g_s && N >= 1 && HW >= 1 && C >= 4 && C % 4 == 0, "modulate_bwd: bad arguments (C must be a multiple of 4)");
    GP3D_CHECK_ARG(gp3d_aligned16(dxs) && gp3d_aligned16(x) && gp3d_aligned16(s) && gp3d_aligned16(dx), "modulate_bwd: pointers must be 16-byte aligned");
    const int chunks = reduce_chunks(N, HW);
    modulate_bwd_kernel<<<N * chunks, 256, 256 * sizeof(float4), (cudaStream_t)stream>>>(dxs, x, s, dx, g_s, N, HW, C, chunks);
    GP3D_RETURN_LAUNCH();
}
