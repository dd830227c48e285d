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
#include <type_traits>
#include "common.cuh"

int gp3d_fir4_launch(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C, int padx0, int padx1, int pady0, int pady1,
                     float* y, void* hi, void* lo, const gp3d_conv_epilogue* epi, cudaStream_t st);   // fir_tma.cu

int gp3d_fir4_up2_launch(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C, int padx0, int pady0, int outH, int outW,
                         float* y, cudaStream_t st);   // fir_tma.cu

int gp3d_fir4_nchw_launch(const void* x, const float* f, void* y, int is_half, int mode, int planes, int H, int W, int outH, int outW, int flip, float gain,
                          cudaStream_t st);            // fir_nchw_tma.cu

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
// W-minor 4x4 kernel for NCHW planes with up / down in {(1,1), (2,1), (1,2)} (upsample2d / downsample2d / filter2d of the reference with
// the [1,3,3,1] filter; BASELINE configs[3]).  A CTA stages the input window of a 128 x 16 output tile of one (n, c) plane in shared
// memory as float; a thread then produces 8 consecutive outputs of one row from registers (every staged value is read once per thread,
// the valid (column, output) tap pairs are resolved at compile time) and stores them as one 16-byte vector when aligned.
// Shared-memory column index with one pad float every S = 8 * DOWN / UP columns (the distance between the windows of neighbouring threads):
// the 16 threads of a row then start on 16 different banks instead of 2 (down = 2), 4 (no resampling) or 8 (up = 2).
template <int LOGS> __device__ __forceinline__ int wm4_col(int c) { return c + (c >> LOGS); }

template <class T, int UP, int DOWN, int KX0, int LOGS>
__device__ __forceinline__ void wminor4_row(const float* __restrict__ row, int w0, const float* __restrict__ fr, float (&acc)[8]) {
    constexpr int NC = (7 * DOWN + 3) / UP + 2;
    float v[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) v[c] = row[wm4_col<LOGS>(w0 + c)];
#pragma unroll
    for (int c = 0; c < NC; c++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int kx = KX0 + c * UP - j * DOWN;          // compile-time after unrolling
            if (kx >= 0 && kx < 4) acc[j] = fmaf(v[c], fr[kx], acc[j]);
        }
    }
}

template <class T, int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_wminor4_kernel(UpfirdnParams p) {
    constexpr int FS = 4, TOW = 128, TOH = 16;
    constexpr int IW = ((TOW - 1) * DOWN + FS - 1) / UP + 2, IH = ((TOH - 1) * DOWN + FS - 1) / UP + 2;
    constexpr int NC = (7 * DOWN + 3) / UP + 2;
    constexpr int LOGS = (DOWN == 2) ? 4 : (UP == 2) ? 2 : 3;        // log2(8 * DOWN / UP)
    constexpr int IWT = IW + NC;                                       // columns staged per row (tail: zero slack for the last thread's window)
    constexpr int IWP = (IWT + (IWT >> LOGS) + 1) | 1;
    __shared__ float sf[FS * FS];
    __shared__ float st[IH * IWP];
    stage_filter(sf, p);
    int64_t tile = blockIdx.x;
    const int tx = (int)(tile % p.tilesX); tile /= p.tilesX;
    const int ty = (int)(tile % p.tilesY); tile /= p.tilesY;
    const int nc = (int)tile;
    const int n = nc / p.C, c = nc - n * p.C;
    const int ox_t0 = tx * TOW, oy_t0 = ty * TOH;
    const int ix_t0 = ceil_div_s(ox_t0 * DOWN - p.padx0, UP), iy_t0 = ceil_div_s(oy_t0 * DOWN - p.pady0, UP);
    const T* xb = (const T*)p.x + n * p.xsN + c * p.xsC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < IH; r += 8) {
        const int iy = iy_t0 + r;
        const bool rowin = (iy >= 0 && iy < p.inH);
        const T* xr = xb + (int64_t)iy * p.xsH;
        for (int cc = lane; cc < IWT; cc += 32) {
            const int ix = ix_t0 + cc;
            st[r * IWP + wm4_col<LOGS>(cc)] = (rowin && cc < IW && ix >= 0 && ix < p.inW) ? io_traits<T>::ld(xr + ix) : 0.f;
        }
    }
    __syncthreads();
    const int xg = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int oy = oy_t0 + ly, ox = ox_t0 + 8 * xg;
    if (oy >= p.outH || ox >= p.outW) return;
    const int by = oy * DOWN - p.pady0, bx0 = ox * DOWN - p.padx0;
    const int iya = ceil_div_s(by, UP), ixa = ceil_div_s(bx0, UP);
    const int ky0 = iya * UP - by, kx0 = ixa * UP - bx0;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    const float* base = st + (iya - iy_t0) * IWP;
    const int w0 = ixa - ix_t0;
#pragma unroll
    for (int r = 0; r < (FS + UP - 1) / UP; r++) {
        const int ky = ky0 + r * UP;
        if (ky < FS) {
            if (UP == 1 || kx0 == 0) wminor4_row<T, UP, DOWN, 0, LOGS>(base + r * IWP, w0, sf + ky * FS, acc);
            else wminor4_row<T, UP, DOWN, 1, LOGS>(base + r * IWP, w0, sf + ky * FS, acc);
        }
    }
    T* yo = (T*)p.y + n * p.ysN + c * p.ysC + (int64_t)oy * p.ysH + ox;
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] *= p.gain;
    if (ox + 8 <= p.outW && (reinterpret_cast<uintptr_t>(yo) & 15u) == 0) {
        if (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(yo) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(yo) + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
            vec16<T> o; o.pack(acc); o.store(yo);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (ox + j < p.outW) io_traits<T>::st(yo + j, acc[j]);
    }
}

// ---------------------------------------------------------------------------------------------
// W-minor tiled kernel (NCHW planes, any dtype / up / down / filter): a CTA stages the input window of a 64 x 16 output tile of one
// (n, c) plane in shared memory as float (coalesced row loads, zero-filled outside the image = the padding rule), then every thread
// evaluates four outputs of one column from shared memory.  The window is read from HBM once (halo excepted) instead of once per tap.
constexpr int kTOW = 64, kTOH = 16;

template <class T>
__global__ void __launch_bounds__(256) upfirdn2d_wminor_tiled_kernel(UpfirdnParams p, int TIW, int TIH) {
    extern __shared__ float tsm[];
    const int taps = p.fh * p.fw;
    float* sf = tsm;
    float* st = tsm + ((taps + 3) & ~3);
    stage_filter(sf, p);
    int64_t tile = blockIdx.x;
    const int tx = (int)(tile % p.tilesX); tile /= p.tilesX;
    const int ty = (int)(tile % p.tilesY); tile /= p.tilesY;
    const int nc = (int)tile;
    const int n = nc / p.C, c = nc - n * p.C;
    const int ox_t0 = tx * kTOW, oy_t0 = ty * kTOH;
    const int ix_t0 = ceil_div_s(ox_t0 * p.downx - p.padx0, p.upx), iy_t0 = ceil_div_s(oy_t0 * p.downy - p.pady0, p.upy);
    const T* xb = (const T*)p.x + n * p.xsN + c * p.xsC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < TIH; r += 8) {
        const int iy = iy_t0 + r;
        const bool rowin = (iy >= 0 && iy < p.inH);
        const T* xr = xb + (int64_t)iy * p.xsH;
        for (int cc = lane; cc < TIW; cc += 32) {
            const int ix = ix_t0 + cc;
            st[r * TIW + cc] = (rowin && ix >= 0 && ix < p.inW) ? io_traits<T>::ld(xr + ix) : 0.f;
        }
    }
    __syncthreads();
    const int lx = threadIdx.x & (kTOW - 1), ly0 = threadIdx.x >> 6;
    const int ox = ox_t0 + lx;
    if (ox >= p.outW) return;
    const int bx = ox * p.downx - p.padx0;
    const int ixa = ceil_div_s(bx, p.upx);
    const int kx0 = ixa * p.upx - bx;
    const int cx = ixa - ix_t0;
    T* yb = (T*)p.y + n * p.ysN + c * p.ysC + ox;
#pragma unroll
    for (int rr = 0; rr < kTOH / 4; rr++) {
        const int oy = oy_t0 + ly0 + 4 * rr;
        if (oy >= p.outH) continue;
        const int by = oy * p.downy - p.pady0;
        const int iya = ceil_div_s(by, p.upy);
        float acc = 0.f;
        int ry = iya - iy_t0;
        for (int ky = iya * p.upy - by; ky < p.fh; ky += p.upy, ry++) {
            const float* row = st + ry * TIW + cx;
            const float* fr = sf + ky * p.fw;
            int q = 0;
            for (int kx = kx0; kx < p.fw; kx += p.upx, q++) acc = fmaf(row[q], fr[kx], acc);
        }
        io_traits<T>::st(yb + (int64_t)oy * p.ysH, acc * p.gain);
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
// C-minor register-tiled kernel for the 4x4 [1,3,3,1] cases that dominate G and D (fp32): each thread produces a PX x PY patch of output
// pixels for one 16-byte channel vector and walks the input rows the patch touches ONCE (up=1: 7 x 7 loads for 16 outputs, i.e. 3 loads
// per output instead of the 16 of a tap-by-tap kernel); the NIX loads of a row are independent, which is what keeps enough bytes in
// flight for the HBM roofline.  Grid: (x-groups, 1, N * ceil(outH / PY)); threads: channel vector fastest -> 512 contiguous bytes per
// warp and pixel.
template <class T, int UP, int DOWN, int PX, int PY>
__global__ void __launch_bounds__(256) upfirdn2d_cminor4_kernel(UpfirdnParams p) {
    constexpr int FS = 4;
    constexpr int VEC = vec16<T>::N;                                 // 4 float / 8 half channels per thread
    constexpr int NIX = ((PX - 1) * DOWN + FS - 1) / UP + 2;       // input columns that can touch PX outputs
    __shared__ float sf[FS * FS];
    stage_filter(sf, p);
    const int CV = p.C / VEC;
    const int cv_per = CV < 256 ? CV : 256;
    const int groups = 256 / cv_per;                                 // x-groups per block
    const int cvl = threadIdx.x % cv_per, xg = threadIdx.x / cv_per;
    const int rows = (p.outH + PY - 1) / PY;
    const int n = blockIdx.z / rows, oy0 = (blockIdx.z - n * rows) * PY;
    const int ox0 = (blockIdx.x * groups + xg) * PX;
    if (xg >= groups || ox0 >= p.outW) return;
    const int by0 = oy0 * DOWN - p.pady0;
    int iy0 = ceil_div_s(by0, UP); if (iy0 < 0) iy0 = 0;
    int iy1 = floor_div(by0 + (PY - 1) * DOWN + FS - 1, UP); if (iy1 > p.inH - 1) iy1 = p.inH - 1;
    const int bx0 = ox0 * DOWN - p.padx0;
    const int ixb = ceil_div_s(bx0, UP);                             // first input column that can contribute to output ox0
    for (int cv = cvl; cv < CV; cv += cv_per) {
        const T* xb = (const T*)p.x + (int64_t)n * p.xsN + VEC * cv;
        float acc[PY][PX][VEC];
#pragma unroll
        for (int i = 0; i < PY; i++)
#pragma unroll
            for (int j = 0; j < PX; j++)
#pragma unroll
                for (int k = 0; k < VEC; k++) acc[i][j][k] = 0.f;
        for (int iy = iy0; iy <= iy1; iy++) {
            const T* xr = xb + (int64_t)iy * p.xsH;
            vec16<T> v[NIX];
#pragma unroll
            for (int t = 0; t < NIX; t++) {
                const int ix = ixb + t;
                if (ix >= 0 && ix < p.inW) v[t].load(xr + (int64_t)ix * p.xsW);
                else v[t].zero();
            }
#pragma unroll
            for (int i = 0; i < PY; i++) {
                const int ky = iy * UP - (by0 + i * DOWN);
                if (ky < 0 || ky >= FS) continue;
                const float* fr = sf + ky * FS;
#pragma unroll
                for (int t = 0; t < NIX; t++) {
                    float fv[VEC];
                    v[t].unpack(fv);
#pragma unroll
                    for (int j = 0; j < PX; j++) {
                        const int kx = (ixb + t) * UP - (bx0 + j * DOWN);
                        if (kx >= 0 && kx < FS) {
                            const float w = fr[kx];
#pragma unroll
                            for (int k = 0; k < VEC; k++) acc[i][j][k] = fmaf(fv[k], w, acc[i][j][k]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < PY; i++) {
            if (oy0 + i >= p.outH) continue;
            T* yo = (T*)p.y + (int64_t)n * p.ysN + (int64_t)(oy0 + i) * p.ysH + VEC * cv;
#pragma unroll
            for (int j = 0; j < PX; j++)
                if (ox0 + j < p.outW) {
#pragma unroll
                    for (int k = 0; k < VEC; k++) acc[i][j][k] *= p.gain;
                    vec16<T> o; o.pack(acc[i][j]); o.store(yo + (int64_t)(ox0 + j) * p.ysW);
                }
        }
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
    // dense channel-minor float32, 4x4, no resampling, C % 32 == 0: the TMA-staged kernel of fir_tma.cu (bit-identical taps order)
    if (std::is_same<T, float>::value && p.fw == 4 && p.fh == 4 && p.upx == 1 && p.upy == 1 && p.downx == 1 && p.downy == 1 && p.C % 32 == 0 && p.C >= 32 &&
        p.xsC == 1 && p.xsW == p.C && p.xsH == (int64_t)p.inW * p.C && p.xsN == (int64_t)p.inH * p.inW * p.C &&
        p.ysC == 1 && p.ysW == p.C && p.ysH == (int64_t)p.outW * p.C && p.ysN == (int64_t)p.outH * p.outW * p.C &&
        (reinterpret_cast<uintptr_t>(p.x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15u) == 0 &&
        (int64_t)p.N * (p.C / 32) <= 65535 && (p.outH + 7) / 8 <= 65535) {
        const int padx1 = p.outW - p.inW - p.padx0 + 3, pady1 = p.outH - p.inH - p.pady0 + 3;
        return gp3d_fir4_launch((const float*)p.x, p.f, p.flip, p.gain, p.N, p.inH, p.inW, p.C, p.padx0, padx1, p.pady0, pady1, (float*)p.y, nullptr, nullptr, nullptr, s);
    }
    if (std::is_same<T, float>::value && p.fw == 4 && p.fh == 4 && p.upx == 2 && p.upy == 2 && p.downx == 1 && p.downy == 1 && p.C % 32 == 0 && p.C >= 32 &&
        p.xsC == 1 && p.xsW == p.C && p.xsH == (int64_t)p.inW * p.C && p.xsN == (int64_t)p.inH * p.inW * p.C &&
        p.ysC == 1 && p.ysW == p.C && p.ysH == (int64_t)p.outW * p.C && p.ysN == (int64_t)p.outH * p.outW * p.C &&
        (reinterpret_cast<uintptr_t>(p.x) & 15u) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15u) == 0 &&
        (int64_t)p.N * (p.C / 32) <= 65535 && (p.outH + 7) / 8 <= 65535) {
        return gp3d_fir4_up2_launch((const float*)p.x, p.f, p.flip, p.gain, p.N, p.inH, p.inW, p.C, p.padx0, p.pady0, p.outH, p.outW, (float*)p.y, s);
    }
    if (cminor && !(wminor && p.C == 1) && (sizeof(T) == 4 || sizeof(T) == 2) && !std::is_same<T, __nv_bfloat16>::value && p.fw == 4 && p.fh == 4 && p.upx == p.upy && p.downx == p.downy &&
        ((p.upx == 1 && p.downx == 1) || (p.upx == 2 && p.downx == 1) || (p.upx == 1 && p.downx == 2)) && (int64_t)p.N * ((p.outH + 1) / 2) <= 65535) {
        constexpr int PX = (sizeof(T) == 4) ? 4 : 2;       // half: 8 channels per thread, so a 2 x 2 patch for the same register budget
        constexpr int PYA = (sizeof(T) == 4) ? 4 : 2;
        const int CV = p.C / VEC, cv_per = CV < 256 ? CV : 256, groups = 256 / cv_per;
        const unsigned gx = (p.outW + groups * PX - 1) / (groups * PX);
        auto gridz = [&](int PY) { return dim3(gx, 1, (unsigned)(p.N * ((p.outH + PY - 1) / PY))); };
        if (p.upx == 1 && p.downx == 1) upfirdn2d_cminor4_kernel<T, 1, 1, PX, PYA><<<gridz(PYA), 256, 0, s>>>(p);
        else if (p.upx == 2) upfirdn2d_cminor4_kernel<T, 2, 1, PX, PYA><<<gridz(PYA), 256, 0, s>>>(p);
        else upfirdn2d_cminor4_kernel<T, 1, 2, PX, 2><<<gridz(2), 256, 0, s>>>(p);     // decimating: 11 input columns per 4 outputs, keep the patch 4 x 2
        return 0;
    }
    if (cminor && !(wminor && p.C == 1)) {
        int64_t total = (int64_t)p.N * p.outH * p.outW * (p.C / VEC);
        upfirdn2d_cminor_kernel<T><<<gp3d_grid_for(total, 256, 8), 256, 0, s>>>(p);
        return 0;
    }
    if (wminor && !std::is_same<T, __nv_bfloat16>::value && p.fw == 4 && p.fh == 4 && p.upx == p.upy && p.downx == p.downy &&
        ((p.upx == 2 && p.downx == 1 && p.padx0 == 2 && p.pady0 == 2) || (p.upx == 1 && p.downx == 2 && p.padx0 == 1 && p.pady0 == 1))) {
        // dense NCHW planes of upsample2d / downsample2d: TMA-staged window, register-blocked taps (fir_nchw_tma.cu)
        const bool dense = p.xsH == p.inW && p.xsC == (int64_t)p.inH * p.inW && p.xsN == (int64_t)p.C * p.inH * p.inW &&
                           p.ysW == 1 && p.ysH == p.outW && p.ysC == (int64_t)p.outH * p.outW && p.ysN == (int64_t)p.C * p.outH * p.outW;
        if (dense && gp3d_fir4_nchw_launch(p.x, p.f, p.y, sizeof(T) == 2, p.upx == 2 ? 0 : 1, p.N * p.C, p.inH, p.inW, p.outH, p.outW, p.flip, p.gain, s) == 0)
            return 0;
    }
    if (wminor) {
        const int TIW = ((kTOW - 1) * p.downx + p.fw - 1) / p.upx + 2, TIH = ((kTOH - 1) * p.downy + p.fh - 1) / p.upy + 2;
        const size_t tsmem = ((size_t)((p.fh * p.fw + 3) & ~3) + (size_t)TIW * TIH) * sizeof(float);
        const bool small4 = (p.fw == 4 && p.fh == 4 && p.upx == p.upy && p.downx == p.downy && p.upx <= 2 && p.downx <= 2);   // unrolled per-tap kernels below
        if (tsmem <= 96 * 1024 && !small4) {
            p.tilesX = (p.outW + kTOW - 1) / kTOW;
            p.tilesY = (p.outH + kTOH - 1) / kTOH;
            const int64_t tblocks = (int64_t)p.N * p.C * p.tilesX * p.tilesY;
            if (tblocks > 2147483647LL) return GP3D_E_TOOLARGE;
            auto kern = upfirdn2d_wminor_tiled_kernel<T>;
            if (tsmem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
            kern<<<(unsigned)tblocks, 256, tsmem, s>>>(p, TIW, TIH);
            return 0;
        }
        if (small4 && ((p.upx == 1 && p.downx == 1) || (p.upx == 2 && p.downx == 1) || (p.upx == 1 && p.downx == 2))) {
            p.tilesX = (p.outW + 127) / 128;
            p.tilesY = (p.outH + 15) / 16;
            const int64_t b4 = (int64_t)p.N * p.C * p.tilesX * p.tilesY;
            if (b4 > 2147483647LL) return GP3D_E_TOOLARGE;
            if (p.upx == 2) upfirdn2d_wminor4_kernel<T, 2, 1><<<(unsigned)b4, 256, 0, s>>>(p);
            else if (p.downx == 2) upfirdn2d_wminor4_kernel<T, 1, 2><<<(unsigned)b4, 256, 0, s>>>(p);
            else upfirdn2d_wminor4_kernel<T, 1, 1><<<(unsigned)b4, 256, 0, s>>>(p);
            return 0;
        }
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
