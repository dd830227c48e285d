// bias_act for sm_100a: y = clamp(act(x + b) * gain) and its 1st / 2nd derivative forms.
// Semantics follow the reference plugin (src/torch_utils/ops/bias_act.cu:23-147, bias_act.cpp:32-90):
// fp32 internal math for every storage type, bias index (i / stepB) % sizeB, `yref / gain` as the saved
// activation, clamp applied on the forward value (grad 0) or as a pass-through mask (grad 1, 2).
// Layout: HBM-bound streaming op -> 16-byte vector loads/stores, grid = multiple of 148 SMs, grid-stride.
#include "common.cuh"

namespace {

struct BiasActParams {
    const void* x; const void* b; const void* xref; const void* yref; const void* dy; void* y;
    int64_t numel, sizeB, stepB;
    int grad;
    float alpha, gain, clamp;
};

template <int A>
__device__ __forceinline__ float act_eval(int G, float x, float xref, float& yref, float gain, float alpha) {
    const float expRange = 80.f, halfExpRange = 40.f;
    const float seluScale = 1.0507009873554804934193349852946f;
    const float seluAlpha = 1.6732632423543772848170429916717f;
    float yy = (gain != 0.f) ? yref / gain : 0.f;
    float y = 0.f;
    if (A == 1) { if (G <= 1) y = x; }
    if (A == 2) { if (G == 0) y = (x > 0.f) ? x : 0.f; if (G == 1) y = (yy > 0.f) ? x : 0.f; }
    if (A == 3) { if (G == 0) y = (x > 0.f) ? x : x * alpha; if (G == 1) y = (yy > 0.f) ? x : x * alpha; }
    if (A == 4) {
        if (G == 0) { float c = expf(x), d = 1.f / c; y = (x < -expRange) ? -1.f : (x > expRange) ? 1.f : (c - d) / (c + d); }
        if (G == 1) y = x * (1.f - yy * yy);
        if (G == 2) y = x * (1.f - yy * yy) * (-2.f * yy);
    }
    if (A == 5) {
        if (G == 0) y = (x < -expRange) ? 0.f : 1.f / (expf(-x) + 1.f);
        if (G == 1) y = x * yy * (1.f - yy);
        if (G == 2) y = x * yy * (1.f - yy) * (1.f - 2.f * yy);
    }
    if (A == 6) {
        if (G == 0) y = (x >= 0.f) ? x : expf(x) - 1.f;
        if (G == 1) y = (yy >= 0.f) ? x : x * (yy + 1.f);
        if (G == 2) y = (yy >= 0.f) ? 0.f : x * (yy + 1.f);
    }
    if (A == 7) {
        if (G == 0) y = (x >= 0.f) ? seluScale * x : (seluScale * seluAlpha) * (expf(x) - 1.f);
        if (G == 1) y = (yy >= 0.f) ? x * seluScale : x * (yy + seluScale * seluAlpha);
        if (G == 2) y = (yy >= 0.f) ? 0.f : x * (yy + seluScale * seluAlpha);
    }
    if (A == 8) {
        if (G == 0) y = (x > expRange) ? x : logf(expf(x) + 1.f);
        if (G == 1) y = x * (1.f - expf(-yy));
        if (G == 2) { float c = expf(-yy); y = x * c * (1.f - c); }
    }
    if (A == 9) {
        if (G == 0) y = (x < -expRange) ? 0.f : x / (expf(-x) + 1.f);
        else {
            float c = expf(xref), d = c + 1.f;
            if (G == 1) y = (xref > halfExpRange) ? x : x * c * (xref + d) / (d * d);
            else        y = (xref > halfExpRange) ? 0.f : x * c * (xref * (2.f - d) + 2.f * d) / (d * d * d);
            yref = (xref < -expRange) ? 0.f : xref / (expf(-xref) + 1.f) * gain;
        }
    }
    return y;
}

template <int A>
__device__ __forceinline__ float bias_act_one(int G, float x, float b, float xref, float yref, float dy,
                                              float alpha, float gain, float clamp) {
    if (G == 0) x += b; else xref += b;
    float y = act_eval<A>(G, x, xref, yref, gain, alpha);
    y *= gain * dy;
    if (clamp >= 0.f) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0.f) ? clamp : -clamp;
        else        y = (yref > -clamp && yref < clamp) ? y : 0.f;
    }
    return y;
}

// BMODE: 0 = no bias, 1 = one bias value per 16B vector (stepB % VEC == 0), 2 = per element (generic),
//        3 = channel-minor (stepB == 1, sizeB % VEC == 0): the vector's bias values are VEC consecutive entries of b.
// IDX: uint32_t when numel < 2^32 (the per-vector bias index costs one 32-bit division instead of two 64-bit ones).
// Each thread keeps kUnroll independent 16-byte vectors of every input stream in flight: one vector per thread leaves a B200 SM
// (2048 threads, ~1 us HBM latency, 44 B/ns per SM) short of the bytes in flight the roofline needs.
// FWD: only x (and b) are read (xref, yref, dy absent: the grad == 0 call) -- four vectors in flight at 4 CTAs / SM; the general form
// carries up to four input streams and keeps two vectors of each in flight.
template <bool FWD> struct ba_unroll { static constexpr int value = FWD ? 4 : 2; };

template <class T, int A, int BMODE, class IDX, bool FWD>
__global__ void __launch_bounds__(256, FWD ? 4 : 3) bias_act_vec_kernel(BiasActParams p) {
    constexpr int VEC = vec16<T>::N;
    constexpr int kUnroll = ba_unroll<FWD>::value;
    const IDX nvec = (IDX)(p.numel / VEC);
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = FWD ? nullptr : (const T*)p.xref;
    const T* yr = FWD ? nullptr : (const T*)p.yref; const T* dyp = FWD ? nullptr : (const T*)p.dy; T* y = (T*)p.y;
    const int G = p.grad;
    const IDX stride = (IDX)gridDim.x * blockDim.x;
    const IDX stepB = (IDX)p.stepB, sizeB = (IDX)p.sizeB;
    for (IDX v0 = (IDX)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * kUnroll) {
        vec16<T> vx[kUnroll], vxr[kUnroll], vyr[kUnroll], vdy[kUnroll], vb[kUnroll];
        float bs[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {                    // all loads first
            const IDX v = v0 + (IDX)u * stride;
            if (v < nvec) {
                const IDX i0 = v * VEC;
                vx[u].load(x + i0);
                if (xr) vxr[u].load(xr + i0);
                if (yr) vyr[u].load(yr + i0);
                if (dyp) vdy[u].load(dyp + i0);
                if (BMODE == 1) bs[u] = io_traits<T>::ld(b + (i0 / stepB) % sizeB);
                if (BMODE == 3) vb[u].load(b + i0 % sizeB);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u++) {
            const IDX v = v0 + (IDX)u * stride;
            if (v < nvec) {
                const IDX i0 = v * VEC;
                float fx[VEC], fxr[VEC], fyr[VEC], fdy[VEC], fy[VEC], fb[VEC];
                vx[u].unpack(fx);
                if (xr) vxr[u].unpack(fxr);
                if (yr) vyr[u].unpack(fyr);
                if (dyp) vdy[u].unpack(fdy);
                if (BMODE == 3) vb[u].unpack(fb);
#pragma unroll
                for (int k = 0; k < VEC; k++) {
                    float bb = 0.f;
                    if (BMODE == 1) bb = bs[u];
                    if (BMODE == 2) bb = io_traits<T>::ld(b + ((i0 + k) / stepB) % sizeB);
                    if (BMODE == 3) bb = fb[k];
                    fy[k] = bias_act_one<A>(G, fx[k], bb, xr ? fxr[k] : 0.f, yr ? fyr[k] : 0.f, dyp ? fdy[k] : 1.f, p.alpha, p.gain, p.clamp);
                }
                vec16<T> vy; vy.pack(fy); vy.store(y + i0);
            }
        }
    }
    // scalar tail (numel % VEC elements), handled by the first few threads of block 0
    if (blockIdx.x == 0) {
        int64_t i = (int64_t)nvec * VEC + threadIdx.x;
        if (i < p.numel) {
            float bb = (BMODE != 0) ? io_traits<T>::ld(b + (i / p.stepB) % p.sizeB) : 0.f;
            float r = bias_act_one<A>(G, io_traits<T>::ld(x + i), bb, xr ? io_traits<T>::ld(xr + i) : 0.f,
                                      yr ? io_traits<T>::ld(yr + i) : 0.f, dyp ? io_traits<T>::ld(dyp + i) : 1.f,
                                      p.alpha, p.gain, p.clamp);
            io_traits<T>::st(y + i, r);
        }
    }
}

// Fully scalar fallback for unaligned pointers.
template <class T, int A>
__global__ void __launch_bounds__(256) bias_act_scalar_kernel(BiasActParams p) {
    const T* x = (const T*)p.x; const T* b = (const T*)p.b; const T* xr = (const T*)p.xref;
    const T* yr = (const T*)p.yref; const T* dyp = (const T*)p.dy; T* y = (T*)p.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.numel; i += (int64_t)gridDim.x * blockDim.x) {
        float bb = b ? io_traits<T>::ld(b + (i / p.stepB) % p.sizeB) : 0.f;
        float r = bias_act_one<A>(p.grad, io_traits<T>::ld(x + i), bb, xr ? io_traits<T>::ld(xr + i) : 0.f,
                                  yr ? io_traits<T>::ld(yr + i) : 0.f, dyp ? io_traits<T>::ld(dyp + i) : 1.f,
                                  p.alpha, p.gain, p.clamp);
        io_traits<T>::st(y + i, r);
    }
}

template <class T, int A>
int launch_bias_act(const BiasActParams& p, cudaStream_t stream) {
    constexpr int VEC = vec16<T>::N;
    bool aligned = gp3d_aligned16(p.x) && gp3d_aligned16(p.y) && (!p.xref || gp3d_aligned16(p.xref)) &&
                   (!p.yref || gp3d_aligned16(p.yref)) && (!p.dy || gp3d_aligned16(p.dy));
    if (!aligned) {
        int grid = gp3d_grid_for(p.numel, 256, 8);
        bias_act_scalar_kernel<T, A><<<grid, 256, 0, stream>>>(p);
        return 0;
    }
    int64_t nvec = p.numel / VEC;
    const bool fwd = !p.xref && !p.yref && !p.dy;
    const int kUnroll = fwd ? 4 : 2;
    int grid = gp3d_grid_for(nvec > 0 ? (nvec + kUnroll - 1) / kUnroll : 1, 256, fwd ? 4 : 3);
    int bmode = (!p.b) ? 0 : (p.stepB % VEC == 0) ? 1 : (p.stepB == 1 && p.sizeB % VEC == 0 && gp3d_aligned16(p.b)) ? 3 : 2;
    const bool small = p.numel < (int64_t)4294967295LL - (int64_t)VEC * 256 * 8 * GP3D_NUM_SMS * 4;   // v0 + u*stride stays below 2^32
#define GP3D_BA(M) do { if (small && fwd) bias_act_vec_kernel<T, A, M, uint32_t, true><<<grid, 256, 0, stream>>>(p); \
                        else if (small) bias_act_vec_kernel<T, A, M, uint32_t, false><<<grid, 256, 0, stream>>>(p); \
                        else bias_act_vec_kernel<T, A, M, int64_t, false><<<grid, 256, 0, stream>>>(p); } while (0)
    if (bmode == 0) GP3D_BA(0); else if (bmode == 1) GP3D_BA(1); else if (bmode == 3) GP3D_BA(3); else GP3D_BA(2);
#undef GP3D_BA
    return 0;
}

template <class T>
int dispatch_act(const BiasActParams& p, int act, cudaStream_t s) {
    switch (act) {
        case 1: return launch_bias_act<T, 1>(p, s);
        case 2: return launch_bias_act<T, 2>(p, s);
        case 3: return launch_bias_act<T, 3>(p, s);
        case 4: return launch_bias_act<T, 4>(p, s);
        case 5: return launch_bias_act<T, 5>(p, s);
        case 6: return launch_bias_act<T, 6>(p, s);
        case 7: return launch_bias_act<T, 7>(p, s);
        case 8: return launch_bias_act<T, 8>(p, s);
        case 9: return launch_bias_act<T, 9>(p, s);
    }
    return -1;
}

}  // namespace

extern "C" int gp3d_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy,
                             void* y, int dtype, int64_t numel, int64_t sizeB, int64_t stepB,
                             int grad, int act, float alpha, float gain, float clamp, void* stream) {
    GP3D_CHECK_ARG(x && y, "bias_act: x and y must be non-null");
    GP3D_CHECK_ARG(numel >= 0 && numel <= 2147483647LL, "bias_act: x is too large (numel=%lld)", (long long)numel);
    GP3D_CHECK_ARG(act >= 1 && act <= 9, "bias_act: unknown activation index %d", act);
    GP3D_CHECK_ARG(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
    GP3D_CHECK_ARG(dtype >= 0 && dtype <= 2, "bias_act: unsupported dtype %d", dtype);
    GP3D_CHECK_ARG(!b || (sizeB >= 1 && stepB >= 1), "bias_act: bad bias geometry");
    if (numel == 0) return GP3D_OK;
    BiasActParams p{x, b, xref, yref, dy, y, numel, b ? sizeB : 1, b ? stepB : 1, grad, alpha, gain, clamp};
    cudaStream_t s = (cudaStream_t)stream;
    int r = (dtype == GP3D_F32) ? dispatch_act<float>(p, act, s)
          : (dtype == GP3D_F16) ? dispatch_act<__half>(p, act, s)
                                : dispatch_act<__nv_bfloat16>(p, act, s);
    GP3D_CHECK_ARG(r == 0, "bias_act: dispatch failed");
    GP3D_RETURN_LAUNCH();
}
