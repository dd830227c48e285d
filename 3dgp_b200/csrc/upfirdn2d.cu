// upfirdn2d for sm_100a: pad -> zero-insert upsample -> FIR -> decimate, per channel.
//
// Index contract (bit-exact with the reference, upfirdn2d.cpp:35-36 and upfirdn2d.cu:20-24,45-49,176-183):
//   outW = (inW*upx + padx0 + padx1 - fw + downx) / downx               (C integer division)
//   for output o the contributing inputs are i in [ceil((o*down - pad0)/up), floor((o*down - pad0 + fw-1)/up)]
//   clipped to [0, in), visited in ascending order, with filter tap k = i*up + pad0 - o*down and coefficient
//   f[fsize-1-k] (true convolution) or f[k] (flip).  Accumulation is fp32 FMA in (y ascending, x ascending)
//   order, gain applied once at the end -- the same order as the reference kernels, so fp32 results agree
//   to the last bit whenever the reference's own two code paths agree.
//
// Three kernels, all HBM-bound streaming ops (algorithmic bytes = (numel_in + numel_out) * sizeof(T)):
//   * W-minor (NCHW): each thread produces 4 consecutive outputs of one row -> one 16B/8B store; input rows are
//     re-read through L1 (every input pixel is touched by <= ceil(fw/up)*ceil(fh/up) neighbouring threads of the
//     same CTA tile, so DRAM sees each byte once).  Specialised for the 4x4 [1,3,3,1] up-2 / down-2 / pad-only
//     cases that occur in G and D, generic otherwise.
//   * C-minor (channels-last): each thread produces one 16-byte channel vector of one pixel; every tap is a
//     coalesced 16-byte load.
//   * any-stride scalar fallback.
#include "common.cuh"

namespace {

struct UpfirdnParams {
    const void* x; const float* f; void* y;
    int N, C, inH, inW;
    int64_t xsN, xsC, xsH, xsW;
    int fh, fw, upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    int outH, outW;
    int64_t ysN, ysC, ysH, ysW;
    int tilesX, tilesY;
};

constexpr int kMaxTaps = 4096;

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
    int q = a / b;
    return (a - q * b < 0) ? q - 1 : q;
}
__device__ __forceinline__ int ceil_div_s(int a, int b) { return -floor_div(-a, b); }

// Loads the (optionally flipped) filter into shared memory as sf[ky*fw + kx] = coefficient of tap (ky,kx).
__device__ __forceinline__ void stage_filter(float* sf, const UpfirdnParams& p) {
    const int taps = p.fh * p.fw;
    for (int t = threadIdx.x; t < taps; t += blockDim.x) {
        int ky = t / p.fw, kx = t - ky * p.fw;
        int sy = p.flip ? ky : p.fh - 1 - ky;
        int sx = p.flip ? kx : p.fw - 1 - kx;
        sf[t] = p.f[sy * p.fw + sx];
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// W-minor kernel.  UP/DOWN/FS == 0 means "runtime value".
template <class T, int UP, int DOWN, int FS>
__global__ void __launch_bounds__(256) upfirdn2d_wminor_kernel(UpfirdnParams p) {
    __shared__ float sf[(FS > 0) ? FS * FS : kMaxTaps];
    stage_filter(sf, p);
    const int upx = UP ? UP : p.upx, upy = UP ? UP : p.upy;
    const int downx = DOWN ? DOWN : p.downx, downy = DOWN ? DOWN : p.downy;
    const int fw = FS ? FS : p.fw, fh = FS ? FS : p.fh;
    constexpr int VX = 4;

    int64_t tile = blockIdx.x;
    const int tx = (int)(tile % p.tilesX); tile /= p.tilesX;
    const int ty = (int)(tile % p.tilesY); tile /= p.tilesY;
    const int nc = (int)tile;
    const int n = nc / p.C, c = nc - n * p.C;

    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;   // 32 x 8 threads -> 128 x 8 outputs
    const int ox0 = (tx * 32 + lx) * VX;
    const int oy = ty * 8 + ly;
    if (ox0 >= p.outW || oy >= p.outH) return;

    const T* xb = (const T*)p.x + n * p.xsN + c * p.xsC;
    const int by = oy * downy - p.pady0;
    int iy0 = ceil_div_s(by, upy); if (iy0 < 0) iy0 = 0;
    int iy1 = floor_div(by + fh - 1, upy); if (iy1 > p.inH - 1) iy1 = p.inH - 1;

    float acc[VX];
    int ix0[VX], ix1[VX], bx[VX];
#pragma unroll
    for (int v = 0; v < VX; v++) {
        acc[v] = 0.f;
        bx[v] = (ox0 + v) * downx - p.padx0;
        int a = ceil_div_s(bx[v], upx); ix0[v] = a < 0 ? 0 : a;
        int b = floor_div(bx[v] + fw - 1, upx); ix1[v] = b > p.inW - 1 ? p.inW - 1 : b;
    }
    for (int iy = iy0; iy <= iy1; iy++) {
        const int ky = iy * upy - by;
        const T* xr = xb + iy * p.xsH;
        const float* fr = sf + ky * fw;
#pragma unroll
        for (int v = 0; v < VX; v++) {
            for (int ix = ix0[v]; ix <= ix1[v]; ix++) {
                const int kx = ix * upx - bx[v];
                acc[v] = fmaf(io_traits<T>::ld(xr + ix), fr[kx], acc[v]);
            }
        }
    }
    T* yo = (T*)p.y + n * p.ysN + c * p.ysC + oy * p.ysH + ox0;
    if (ox0 + VX <= p.outW && sizeof(T) == 4 && ((reinterpret_cast<uintptr_t>(yo) & 15u) == 0)) {
        *reinterpret_cast<float4*>(yo) = make_float4(acc[0] * p.gain, acc[1] * p.gain, acc[2] * p.gain, acc[3] * p.gain);
    } else {
#pragma unroll
        for (int v = 0; v < VX; v++)
            if (ox0 + v < p.outW) io_traits<T>::st(yo + v, acc[v] * p.gain);
    }
}

// ---------------------------------------------------------------------------------------------
// C-minor (channels-last) kernel: one 16-byte channel vector per thread.
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_cminor_kernel(UpfirdnParams p) {
    __shared__ float sf[kMaxTaps];
    stage_filter(sf, p);
    constexpr int VEC = vec16<T>::N;
    const int cv = p.C / VEC;
    const int64_t total = (int64_t)p.N * p.outH * p.outW * cv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx;
        const int c0 = (int)(r % cv) * VEC; r /= cv;
        const int ox = (int)(r % p.outW); r /= p.outW;
        const int oy = (int)(r % p.outH); r /= p.outH;
        const int n = (int)r;
        const int by = oy * p.downy - p.pady0, bx = ox * p.downx - p.padx0;
        int iy0 = ceil_div_s(by, p.upy); if (iy0 < 0) iy0 = 0;
        int iy1 = floor_div(by + p.fh - 1, p.upy); if (iy1 > p.inH - 1) iy1 = p.inH - 1;
        int ix0 = ceil_div_s(bx, p.upx); if (ix0 < 0) ix0 = 0;
        int ix1 = floor_div(bx + p.fw - 1, p.upx); if (ix1 > p.inW - 1) ix1 = p.inW - 1;
        float acc[VEC];
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] = 0.f;
        const T* xb = (const T*)p.x + n * p.xsN + c0;
        for (int iy = iy0; iy <= iy1; iy++) {
            const float* fr = sf + (iy * p.upy - by) * p.fw;
            for (int ix = ix0; ix <= ix1; ix++) {
                const float w = fr[ix * p.upx - bx];
                vec16<T> v; float fv[VEC];
                v.load(xb + iy * p.xsH + ix * p.xsW); v.unpack(fv);
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[k] = fmaf(fv[k], w, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) acc[k] *= p.gain;
        vec16<T> o; o.pack(acc);
        o.store((T*)p.y + n * p.ysN + oy * p.ysH + ox * p.ysW + c0);
    }
}

// ---------------------------------------------------------------------------------------------
// C-minor register-tiled kernel for the 4x4 [1,3,3,1] cases that dominate G and D (fp32): each thread produces PX consecutive
// output pixels of one row for one 16-byte channel vector, loading every needed input vector ONCE (7 x 4 loads for 4 outputs of the
// up=1 FIR instead of 64).  Grid: (x-groups, 1, N*outH); threads: channel vector fastest -> 512 contiguous bytes per warp and pixel.
template <int UP, int DOWN, int PX>
__global__ void __launch_bounds__(256) upfirdn2d_cminor4_kernel(UpfirdnParams p) {
    constexpr int FS = 4;
    constexpr int NIX = ((PX - 1) * DOWN + FS - 1) / UP + 2;       // input columns that can touch PX outputs
    __shared__ float sf[FS * FS];
    stage_filter(sf, p);
    const int CV = p.C / 4;
    const int cv_per = CV < 256 ? CV : 256;
    const int groups = 256 / cv_per;                                 // x-groups per block
    const int cvl = threadIdx.x % cv_per, xg = threadIdx.x / cv_per;
    const int n = blockIdx.z / p.outH, oy = blockIdx.z - n * p.outH;
    const int ox0 = (blockIdx.x * groups + xg) * PX;
    if (xg >= groups || ox0 >= p.outW) return;
    const int by = oy * DOWN - p.pady0;
    int iy0 = ceil_div_s(by, UP); if (iy0 < 0) iy0 = 0;
    int iy1 = floor_div(by + FS - 1, UP); if (iy1 > p.inH - 1) iy1 = p.inH - 1;
    const int bx0 = ox0 * DOWN - p.padx0;
    const int ixb = ceil_div_s(bx0, UP);                             // first input column that can contribute to output ox0
    for (int cv = cvl; cv < CV; cv += cv_per) {
        const float* xb = (const float*)p.x + (int64_t)n * p.xsN + 4 * cv;
        float4 acc[PX];
#pragma unroll
        for (int j = 0; j < PX; j++) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int iy = iy0; iy <= iy1; iy++) {
            const float* fr = sf + (iy * UP - by) * FS;
            const float* xr = xb + (int64_t)iy * p.xsH;
#pragma unroll
            for (int t = 0; t < NIX; t++) {
                const int ix = ixb + t;
                if (ix < 0 || ix >= p.inW) continue;
                const float4 v = *reinterpret_cast<const float4*>(xr + (int64_t)ix * p.xsW);
#pragma unroll
                for (int j = 0; j < PX; j++) {
                    const int kx = ix * UP - (bx0 + j * DOWN);
                    if (kx >= 0 && kx < FS) {
                        const float w = fr[kx];
                        acc[j].x = fmaf(v.x, w, acc[j].x); acc[j].y = fmaf(v.y, w, acc[j].y);
                        acc[j].z = fmaf(v.z, w, acc[j].z); acc[j].w = fmaf(v.w, w, acc[j].w);
                    }
                }
            }
        }
        float* yo = (float*)p.y + (int64_t)n * p.ysN + (int64_t)oy * p.ysH + 4 * cv;
#pragma unroll
        for (int j = 0; j < PX; j++)
            if (ox0 + j < p.outW)
                *reinterpret_cast<float4*>(yo + (int64_t)(ox0 + j) * p.ysW) =
                    make_float4(acc[j].x * p.gain, acc[j].y * p.gain, acc[j].z * p.gain, acc[j].w * p.gain);
    }
}

// ---------------------------------------------------------------------------------------------
// Any-stride scalar fallback: one output element per thread.
template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_generic_kernel(UpfirdnParams p) {
    __shared__ float sf[kMaxTaps];
    stage_filter(sf, p);
    const int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx;
        const int ox = (int)(r % p.outW); r /= p.outW;
        const int oy = (int)(r % p.outH); r /= p.outH;
        const int c = (int)(r % p.C); r /= p.C;
        const int n = (int)r;
        const int by = oy * p.downy - p.pady0, bx = ox * p.downx - p.padx0;
        int iy0 = ceil_div_s(by, p.upy); if (iy0 < 0) iy0 = 0;
        int iy1 = floor_div(by + p.fh - 1, p.upy); if (iy1 > p.inH - 1) iy1 = p.inH - 1;
        int ix0 = ceil_div_s(bx, p.upx); if (ix0 < 0) ix0 = 0;
        int ix1 = floor_div(bx + p.fw - 1, p.upx); if (ix1 > p.inW - 1) ix1 = p.inW - 1;
        const T* xb = (const T*)p.x + n * p.xsN + c * p.xsC;
        float acc = 0.f;
        for (int iy = iy0; iy <= iy1; iy++) {
            const float* fr = sf + (iy * p.upy - by) * p.fw;
            for (int ix = ix0; ix <= ix1; ix++)
                acc = fmaf(io_traits<T>::ld(xb + iy * p.xsH + ix * p.xsW), fr[ix * p.upx - bx], acc);
        }
        io_traits<T>::st((T*)p.y + n * p.ysN + c * p.ysC + oy * p.ysH + ox * p.ysW, acc * p.gain);
    }
}

template <class T>
int launch_upfirdn(UpfirdnParams& p, cudaStream_t s) {
    constexpr int VEC = vec16<T>::N;
    const bool wminor = (p.xsW == 1 && p.ysW == 1);
    const bool cminor = (p.xsC == 1 && p.ysC == 1 && p.C % VEC == 0 && gp3d_aligned16(p.x) && gp3d_aligned16(p.y) &&
                         p.xsW % VEC == 0 && p.xsH % VEC == 0 && p.xsN % VEC == 0 &&
                         p.ysW % VEC == 0 && p.ysH % VEC == 0 && p.ysN % VEC == 0);
    if (cminor && !(wminor && p.C == 1) && sizeof(T) == 4 && p.fw == 4 && p.fh == 4 && p.upx == p.upy && p.downx == p.downy &&
        ((p.upx == 1 && p.downx == 1) || (p.upx == 2 && p.downx == 1) || (p.upx == 1 && p.downx == 2)) && (int64_t)p.N * p.outH <= 65535) {
        constexpr int PX = 4;
        const int CV = p.C / 4, cv_per = CV < 256 ? CV : 256, groups = 256 / cv_per;
        dim3 grid((p.outW + groups * PX - 1) / (groups * PX), 1, p.N * p.outH);
        if (p.upx == 1 && p.downx == 1) upfirdn2d_cminor4_kernel<1, 1, PX><<<grid, 256, 0, s>>>(p);
        else if (p.upx == 2) upfirdn2d_cminor4_kernel<2, 1, PX><<<grid, 256, 0, s>>>(p);
        else upfirdn2d_cminor4_kernel<1, 2, PX><<<grid, 256, 0, s>>>(p);
        return 0;
    }
    if (cminor && !(wminor && p.C == 1)) {
        int64_t total = (int64_t)p.N * p.outH * p.outW * (p.C / VEC);
        upfirdn2d_cminor_kernel<T><<<gp3d_grid_for(total, 256, 8), 256, 0, s>>>(p);
        return 0;
    }
    if (wminor) {
        p.tilesX = (p.outW + 127) / 128;
        p.tilesY = (p.outH + 7) / 8;
        int64_t blocks = (int64_t)p.N * p.C * p.tilesX * p.tilesY;
        if (blocks > 2147483647LL) return GP3D_E_TOOLARGE;
        const bool sq = (p.upx == p.upy && p.downx == p.downy && p.fw == p.fh);
        if (sq && p.fw == 4 && p.upx == 2 && p.downx == 1) upfirdn2d_wminor_kernel<T, 2, 1, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
        else if (sq && p.fw == 4 && p.upx == 1 && p.downx == 2) upfirdn2d_wminor_kernel<T, 1, 2, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
        else if (sq && p.fw == 4 && p.upx == 1 && p.downx == 1) upfirdn2d_wminor_kernel<T, 1, 1, 4><<<(unsigned)blocks, 256, 0, s>>>(p);
        else upfirdn2d_wminor_kernel<T, 0, 0, 0><<<(unsigned)blocks, 256, 0, s>>>(p);
        return 0;
    }
    int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    upfirdn2d_generic_kernel<T><<<gp3d_grid_for(total, 256, 8), 256, 0, s>>>(p);
    return 0;
}

}  // namespace

extern "C" int gp3d_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int fsize) {
    if (up < 1 || down < 1 || fsize < 1) return GP3D_E_BADARG;
    return (in_size * up + pad0 + pad1 - fsize + down) / down;
}

extern "C" int gp3d_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                              int N, int C, int inH, int inW,
                              int64_t xsN, int64_t xsC, int64_t xsH, int64_t xsW,
                              int fh, int fw, int upx, int upy, int downx, int downy,
                              int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                              int outH, int outW,
                              int64_t ysN, int64_t ysC, int64_t ysH, int64_t ysW, void* stream) {
    GP3D_CHECK_ARG(x && f && y, "upfirdn2d: null pointer");
    GP3D_CHECK_ARG(N >= 1 && C >= 1 && inH >= 1 && inW >= 1, "upfirdn2d: x has zero size");
    GP3D_CHECK_ARG(fh >= 1 && fw >= 1, "upfirdn2d: f must be at least 1x1");
    GP3D_CHECK_ARG(fh * fw <= kMaxTaps, "upfirdn2d: filter with %d taps exceeds the %d-tap limit", fh * fw, kMaxTaps);
    GP3D_CHECK_ARG(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");
    GP3D_CHECK_ARG(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");
    GP3D_CHECK_ARG(dtype >= 0 && dtype <= 2, "upfirdn2d: unsupported dtype %d", dtype);
    const int eW = gp3d_upfirdn2d_out_size(inW, upx, downx, padx0, padx1, fw);
    const int eH = gp3d_upfirdn2d_out_size(inH, upy, downy, pady0, pady1, fh);
    GP3D_CHECK_ARG(eW >= 1 && eH >= 1, "upfirdn2d: output must be at least 1x1");
    GP3D_CHECK_ARG(eW == outW && eH == outH, "upfirdn2d: output extent %dx%d does not match the index formula %dx%d", outH, outW, eH, eW);
    GP3D_CHECK_ARG((int64_t)N * C * inH * inW <= 2147483647LL && (int64_t)N * C * outH * outW <= 2147483647LL,
                   "upfirdn2d: tensor is too large");
    UpfirdnParams p{x, f, y, N, C, inH, inW, xsN, xsC, xsH, xsW, fh, fw, upx, upy, downx, downy, padx0, pady0,
                    flip ? 1 : 0, gain, outH, outW, ysN, ysC, ysH, ysW, 0, 0};
    cudaStream_t s = (cudaStream_t)stream;
    int r = (dtype == GP3D_F32) ? launch_upfirdn<float>(p, s)
          : (dtype == GP3D_F16) ? launch_upfirdn<__half>(p, s)
                                : launch_upfirdn<__nv_bfloat16>(p, s);
    if (r != 0) { gp3d_set_error("upfirdn2d: launch geometry too large"); return r; }
    GP3D_RETURN_LAUNCH();
}
