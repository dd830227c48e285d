// 4x4 FIR (up = down = 1) over channel-minor float32 tensors with the input window staged by TMA (sm_100a).
// This is the filter that follows every up-sampling convolution of the tri-plane decoder (conv2d_resample.py:119-126: upfirdn2d with
// pad 1 and gain up^2 on the (2H+1)^2 transposed-conv output) and, in the backward, its adjoint (pad 2, flipped filter).
//   * one CTA = 32 x 8 output pixels x 32 channels: ONE 4-D TMA box {32 ch, 35, 11, 1} lands the haloed input window in shared memory
//     (49 KB; out-of-range rows / columns are zero-filled by the TMA unit = the padding rule, no bounds logic in the kernel);
//     3 CTAs per SM keep ~150 KB of loads in flight per SM, which is what the HBM roofline needs (the register-tiled LDG version ran at 57 %);
//   * a thread owns 4 channels and a 4 x 2 pixel patch: 35 conflict-free LDS.128 feed 128 float4 FMAs; taps accumulate in the same
//     order as upfirdn2d_cminor4_kernel, so results are bit-identical to the generic op;
//   * fused epilogues: (1) the modulated-conv layer's demodulation + noise + bias + lrelu (networks_stylegan2.py:71,144) - the filtered
//     tensor is never stored; (2) bf16 (hi, lo) operand pair for the tensor-core input- / weight-gradient kernels (backward).
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int FTW = 32, FTH = 8, FCB = 32, FIW = FTW + 3, FIH = FTH + 3;
constexpr uint32_t kFirTileBytes = FIH * FIW * FCB * 4;

struct Fir4Params {
    float* y; __nv_bfloat16* hi; __nv_bfloat16* lo;
    const float* f; int flip; float gain;
    int N, C, outH, outW, padx0, pady0;
    const float* d; const float* nz; const float* b; int nps, act; float alpha, g2;
};

// EPI: 0 plain, 1 demod + noise + bias + activation, 2 bf16 hi/lo pair
template <int EPI>
__global__ void __launch_bounds__(256, 3) fir4_tma_kernel(const __grid_constant__ CUtensorMap tmX, Fir4Params p) {
    extern __shared__ unsigned char fsm_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fsm_raw) + 127) & ~(uintptr_t)127);
    float* tile = reinterpret_cast<float*>(base);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + kFirTileBytes);
    float* sf = reinterpret_cast<float*>(bar + 1);
    const int tid = threadIdx.x;
    const int cblocks = p.C / FCB;
    const int n = blockIdx.z / cblocks, cb = blockIdx.z - n * cblocks;
    const int ox0 = blockIdx.x * FTW, oy0 = blockIdx.y * FTH;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (tid < 16) {                                   // sf[ky*4+kx] = coefficient of tap (ky, kx), as upfirdn2d.cu::stage_filter
        const int ky = tid >> 2, kx = tid & 3;
        const int sy = p.flip ? ky : 3 - ky, sx = p.flip ? kx : 3 - kx;
        sf[tid] = p.f[sy * 4 + sx];
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, kFirTileBytes);
        tma_load_4d(tile, &tmX, bar, cb * FCB, ox0 - p.padx0, oy0 - p.pady0, n);
    }
    const int cv = tid & 7, xg = (tid >> 3) & 7, yg = tid >> 6;
    float4 acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait(bar, 0);
    const float* tb = tile + ((2 * yg) * FIW + 4 * xg) * FCB + 4 * cv;
#pragma unroll
    for (int r = 0; r < 5; r++) {
        float4 v[7];
#pragma unroll
        for (int t = 0; t < 7; t++) v[t] = *reinterpret_cast<const float4*>(tb + (r * FIW + t) * FCB);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int ky = r - i;
            if (ky < 0 || ky >= 4) continue;
#pragma unroll
            for (int t = 0; t < 7; t++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int kx = t - j;
                    if (kx >= 0 && kx < 4) {
                        const float w = sf[ky * 4 + kx];
                        acc[i][j].x = fmaf(v[t].x, w, acc[i][j].x); acc[i][j].y = fmaf(v[t].y, w, acc[i][j].y);
                        acc[i][j].z = fmaf(v[t].z, w, acc[i][j].z); acc[i][j].w = fmaf(v[t].w, w, acc[i][j].w);
                    }
                }
            }
        }
    }
    const int c0 = cb * FCB + 4 * cv;
    float4 dv = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == 1) {
        if (p.d) dv = *reinterpret_cast<const float4*>(p.d + (size_t)n * p.C + c0);
        if (p.b) bv = *reinterpret_cast<const float4*>(p.b + c0);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int oy = oy0 + 2 * yg + i;
        if (oy >= p.outH) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int ox = ox0 + 4 * xg + j;
            if (ox >= p.outW) continue;
            float o[4] = {acc[i][j].x * p.gain, acc[i][j].y * p.gain, acc[i][j].z * p.gain, acc[i][j].w * p.gain};
            const size_t pix = ((size_t)n * p.outH + oy) * p.outW + ox;
            if (EPI == 1) {
                const float nz = p.nz ? p.nz[(p.nps ? (size_t)n * p.outH * p.outW : 0) + (size_t)oy * p.outW + ox] : 0.f;
                o[0] = fmaf(o[0], dv.x, nz) + bv.x; o[1] = fmaf(o[1], dv.y, nz) + bv.y; o[2] = fmaf(o[2], dv.z, nz) + bv.z; o[3] = fmaf(o[3], dv.w, nz) + bv.w;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (p.act == 3) o[k] = (o[k] > 0.f) ? o[k] : o[k] * p.alpha;
                    o[k] *= p.g2;
                }
            }
            if (EPI == 2) {
                __nv_bfloat16 h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { h[k] = __float2bfloat16_rn(o[k]); l[k] = __float2bfloat16_rn(o[k] - __bfloat162float(h[k])); }
                *reinterpret_cast<uint2*>(p.hi + pix * p.C + c0) = *reinterpret_cast<const uint2*>(h);
                *reinterpret_cast<uint2*>(p.lo + pix * p.C + c0) = *reinterpret_cast<const uint2*>(l);
            } else {
                *reinterpret_cast<float4*>(p.y + pix * p.C + c0) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// ---- up = 2 (zero-stuffing) form: the skip-image up-sampling of the tri-plane decoder (upsample2d, networks_stylegan2.py:268).
// Output tile 32 x 8; the input window is at most 18 x 6 pixels (13.8 KB); a thread's 4 x 2 patch touches 4 x 3 input pixels, 4 taps per output.
constexpr int UIW = 18, UIH = 6;
constexpr uint32_t kUpTileBytes = UIH * UIW * FCB * 4;

__global__ void __launch_bounds__(256, 4) fir4_up2_tma_kernel(const __grid_constant__ CUtensorMap tmX, Fir4Params p) {
    extern __shared__ unsigned char fsm_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(fsm_raw) + 127) & ~(uintptr_t)127);
    float* tile = reinterpret_cast<float*>(base);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + kUpTileBytes);
    float* sf = reinterpret_cast<float*>(bar + 1);
    const int tid = threadIdx.x;
    const int cblocks = p.C / FCB;
    const int n = blockIdx.z / cblocks, cb = blockIdx.z - n * cblocks;
    const int ox0 = blockIdx.x * FTW, oy0 = blockIdx.y * FTH;
    auto cdiv2 = [](int a) { return (a >= 0) ? (a + 1) / 2 : -((-a) / 2); };          // ceil(a / 2)
    const int ix_t0 = cdiv2(ox0 - p.padx0), iy_t0 = cdiv2(oy0 - p.pady0);
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (tid < 16) {
        const int ky = tid >> 2, kx = tid & 3;
        const int sy = p.flip ? ky : 3 - ky, sx = p.flip ? kx : 3 - kx;
        sf[tid] = p.f[sy * 4 + sx];
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, kUpTileBytes);
        tma_load_4d(tile, &tmX, bar, cb * FCB, ix_t0, iy_t0, n);
    }
    const int cv = tid & 7, xg = (tid >> 3) & 7, yg = tid >> 6;
    const int oyb = oy0 + 2 * yg, oxb = ox0 + 4 * xg;
    const int iyb = cdiv2(oyb - p.pady0), ixb = cdiv2(oxb - p.padx0);
    float4 acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    mbar_wait(bar, 0);
    const float* tb = tile + ((iyb - iy_t0) * UIW + (ixb - ix_t0)) * FCB + 4 * cv;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float4 v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) v[t] = *reinterpret_cast<const float4*>(tb + (r * UIW + t) * FCB);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int ky = 2 * (iyb + r) - (oyb + i - p.pady0);
            if (ky < 0 || ky >= 4) continue;
#pragma unroll
            for (int t = 0; t < 4; t++) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int kx = 2 * (ixb + t) - (oxb + j - p.padx0);
                    if (kx >= 0 && kx < 4) {
                        const float w = sf[ky * 4 + kx];
                        acc[i][j].x = fmaf(v[t].x, w, acc[i][j].x); acc[i][j].y = fmaf(v[t].y, w, acc[i][j].y);
                        acc[i][j].z = fmaf(v[t].z, w, acc[i][j].z); acc[i][j].w = fmaf(v[t].w, w, acc[i][j].w);
                    }
                }
            }
        }
    }
    const int c0 = cb * FCB + 4 * cv;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int oy = oyb + i;
        if (oy >= p.outH) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int ox = oxb + j;
            if (ox >= p.outW) continue;
            const size_t pix = ((size_t)n * p.outH + oy) * p.outW + ox;
            *reinterpret_cast<float4*>(p.y + pix * p.C + c0) = make_float4(acc[i][j].x * p.gain, acc[i][j].y * p.gain, acc[i][j].z * p.gain, acc[i][j].w * p.gain);
        }
    }
}

}  // namespace

// up = 2, down = 1, 4x4 filter, dense [N][H][W][C] float32 -> [N][outH][outW][C] (called by upfirdn2d.cu's dispatcher)
int gp3d_fir4_up2_launch(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C, int padx0, int pady0, int outH, int outW,
                         float* y, cudaStream_t st) {
    GP3D_CHECK_ARG(x && f && y && N >= 1 && H >= 1 && W >= 1 && C >= FCB && C % FCB == 0 && outH >= 1 && outW >= 1, "fir4_up2: bad shape (C must be a multiple of 32)");
    GP3D_CHECK_ARG((int64_t)N * (C / FCB) <= 65535 && (outH + FTH - 1) / FTH <= 65535, "fir4_up2: grid too large");
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {FCB, UIW, UIH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gp3d_set_error("fir4_up2: tensor map encode failed (CUresult %d)", (int)r); return GP3D_E_BADARG; }
    Fir4Params p{};
    p.y = y; p.f = f; p.flip = flip; p.gain = gain; p.N = N; p.C = C; p.outH = outH; p.outW = outW; p.padx0 = padx0; p.pady0 = pady0;
    const size_t smem = 128 + kUpTileBytes + 8 + 64;
    const dim3 grid((outW + FTW - 1) / FTW, (outH + FTH - 1) / FTH, N * (C / FCB));
    fir4_up2_tma_kernel<<<grid, 256, smem, st>>>(tm, p);
    return 0;
}

// Internal launcher shared with upfirdn2d.cu's dispatcher.  x: dense [N][H][W][C] float32.
int gp3d_fir4_launch(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C, int padx0, int padx1, int pady0, int pady1,
                     float* y, void* hi, void* lo, const gp3d_conv_epilogue* epi, cudaStream_t st) {
    const int outH = H + pady0 + pady1 - 3, outW = W + padx0 + padx1 - 3;
    GP3D_CHECK_ARG(x && f && N >= 1 && H >= 1 && W >= 1 && C >= FCB && C % FCB == 0 && outH >= 1 && outW >= 1, "fir4_nhwc: bad shape (C must be a multiple of 32)");
    GP3D_CHECK_ARG((y != nullptr) != (hi != nullptr) && (hi == nullptr) == (lo == nullptr), "fir4_nhwc: give either y or the (hi, lo) pair");
    GP3D_CHECK_ARG(!(epi && hi), "fir4_nhwc: the demodulation epilogue writes float32");
    GP3D_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15u) == 0 && (!y || (reinterpret_cast<uintptr_t>(y) & 15u) == 0), "fir4_nhwc: pointers must be 16-byte aligned");
    GP3D_CHECK_ARG((int64_t)N * (C / FCB) <= 65535 && (outH + FTH - 1) / FTH <= 65535, "fir4_nhwc: grid too large");
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {FCB, FIW, FIH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gp3d_set_error("fir4_nhwc: tensor map encode failed (CUresult %d)", (int)r); return GP3D_E_BADARG; }
    Fir4Params p{};
    p.y = y; p.hi = (__nv_bfloat16*)hi; p.lo = (__nv_bfloat16*)lo; p.f = f; p.flip = flip; p.gain = gain;
    p.N = N; p.C = C; p.outH = outH; p.outW = outW; p.padx0 = padx0; p.pady0 = pady0;
    if (epi) {
        GP3D_CHECK_ARG(epi->act == 1 || epi->act == 3, "fir4_nhwc: fused activation must be linear (1) or lrelu (3)");
        p.d = epi->dcoef; p.nz = epi->noise; p.b = epi->bias; p.nps = epi->noise_per_sample; p.act = epi->act; p.alpha = epi->alpha; p.g2 = epi->gain;
    }
    const size_t smem = 128 + kFirTileBytes + 8 + 64;
    const dim3 grid((outW + FTW - 1) / FTW, (outH + FTH - 1) / FTH, N * (C / FCB));
#define GP3D_FIR(E) do { auto k = fir4_tma_kernel<E>; \
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) { gp3d_set_error("fir4_nhwc: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; } \
        k<<<grid, 256, smem, st>>>(tm, p); } while (0)
    if (epi) GP3D_FIR(1); else if (hi) GP3D_FIR(2); else GP3D_FIR(0);
#undef GP3D_FIR
    return 0;
}

extern "C" int gp3d_fir4_nhwc(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C,
                              int padx0, int padx1, int pady0, int pady1, float* y, void* hi, void* lo,
                              const gp3d_conv_epilogue* epi, void* stream) {
    int rc = gp3d_fir4_launch(x, f, flip, gain, N, H, W, C, padx0, padx1, pady0, pady1, y, hi, lo, epi, (cudaStream_t)stream);
    if (rc) return rc;
    GP3D_RETURN_LAUNCH();
}
