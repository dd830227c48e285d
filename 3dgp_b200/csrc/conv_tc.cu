// Stride-1 "same" convolution (correlation, 1x1 or 3x3) as an implicit GEMM on tcgen05 / TMEM (sm_100a).
//
//   y[n][oy][ox][co] (+)= sum_{ky,kx,ci} x[n][oy+ky-p][ox+kx-p][ci] * w[co][ky][kx][ci]       (zero padding, p = k/2)
//
// GEMM view: M = output pixels (tile of 128 = TN images x TH rows x TW cols), N = Cout (tile BN), K = k*k*Cin in blocks
// of 64 channels of one filter tap.  The A operand of a k-block is the [128 pixels][64 channels] window of the NHWC
// activation tensor shifted by the tap offset -- fetched by ONE 4-D TMA box {C:64, W:TW, H:TH, N:TN} whose out-of-bounds
// coordinates (the padding halo, including negative ones) are zero-filled by the TMA unit, landing in shared memory
// directly in the 128B-swizzled K-major layout tcgen05.mma consumes.  No im2col buffer exists anywhere.
// The B operand is the matching [BN couts][64 channels] slab of the [Cout][k*k*Cin] weight matrix (2-D TMA).
// Pipeline, warp roles and epilogue are those of gemm_tc.cu.
#include "tc_common.cuh"

namespace tc {

constexpr int CBM = 128, CBK = 64;
constexpr int kConvThreads = 256;

struct ConvGeom {
    int N, H, W, Cin, Cout, ks;
    int TW, TH, TN;            // pixel tile: TN images x TH rows x TW cols = 128
    int tiles_x, tiles_y, tiles_n, tiles_co;
};

// TERMS == 1: y += xh * wh.   TERMS == 3 (error-compensated "bf16x3", ~2^-16 relative): y += xh*wh + xh*wl + xl*wh with
// x = xh + xl, w = wh + wl split into bf16 pairs; all three products accumulate into the same TMEM tile.
template <int BN, int TERMS>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_nhwc_bf16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmXl, const __grid_constant__ CUtensorMap tmWl,
                      float* __restrict__ Y, ConvGeom g, int accumulate) {
    constexpr int CSTAGES = (TERMS == 3) ? 3 : 4;
    constexpr uint32_t kOperand = (CBM + BN) * CBK * 2;
    constexpr uint32_t kStage = kOperand * (TERMS == 3 ? 2 : 1);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // keep every stage 1024-byte aligned: A tile is 16 KB, B tile BN*128 B (multiple of 1024 for BN % 8 == 0)
    unsigned char* tiles = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)CSTAGES * kStage);
    uint64_t* empty_bar = full_bar + CSTAGES;
    uint64_t* accum_bar = empty_bar + CSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // tile decomposition: blockIdx.x = ((tn * tiles_y + ty) * tiles_x + tx) * tiles_co + tco   (cout fastest: activations reused from L2)
    int t = blockIdx.x;
    const int tco = t % g.tiles_co; t /= g.tiles_co;
    const int tx = t % g.tiles_x; t /= g.tiles_x;
    const int ty = t % g.tiles_y; t /= g.tiles_y;
    const int tn = t;
    const int x0 = tx * g.TW, y0 = ty * g.TH, n0 = tn * g.TN, co0 = tco * BN;
    const int pad = g.ks / 2;
    const int cblocks = g.Cin / CBK;
    const int num_kb = g.ks * g.ks * cblocks;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < CSTAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; kb++) {
                const int st = kb % CSTAGES; const uint32_t ph = (kb / CSTAGES) & 1;
                const int tap = kb / cblocks, cb = kb - tap * cblocks;
                const int ky = tap / g.ks, kx = tap - ky * g.ks;
                mbar_wait(&empty_bar[st], ph ^ 1);
                unsigned char* sa = tiles + (size_t)st * kStage;
                unsigned char* sb = sa + CBM * CBK * 2;
                mbar_expect_tx(&full_bar[st], kStage);
                tma_load_4d(sa, &tmX, &full_bar[st], cb * CBK, x0 + kx - pad, y0 + ky - pad, n0);
                tma_load_2d(sb, &tmW, &full_bar[st], tap * g.Cin + cb * CBK, co0);
                if (TERMS == 3) {
                    tma_load_4d(sa + kOperand, &tmXl, &full_bar[st], cb * CBK, x0 + kx - pad, y0 + ky - pad, n0);
                    tma_load_2d(sb + kOperand, &tmWl, &full_bar[st], tap * g.Cin + cb * CBK, co0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16_f32(CBM, BN);
            for (int kb = 0; kb < num_kb; kb++) {
                const int st = kb % CSTAGES; const uint32_t ph = (kb / CSTAGES) & 1;
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + (size_t)st * kStage);
                const uint32_t sb = sa + CBM * CBK * 2;
#pragma unroll
                for (int k = 0; k < CBK / 16; k++) {
                    const uint64_t dah = make_desc_k_sw128(sa + k * 32), dbh = make_desc_k_sw128(sb + k * 32);
                    umma_bf16(tmem_base, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    if (TERMS == 3) {
                        umma_bf16(tmem_base, dah, make_desc_k_sw128(sb + kOperand + k * 32), idesc, 1u);
                        umma_bf16(tmem_base, make_desc_k_sw128(sa + kOperand + k * 32), dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[st]);
            }
            umma_commit(accum_bar);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int r = q * 32 + lane;                       // pixel index inside the tile: ((n*TH + h)*TW + w)
        const int wl = r % g.TW, hl = (r / g.TW) % g.TH, nl = r / (g.TW * g.TH);
        float* yrow = Y + (((size_t)(n0 + nl) * g.H + (y0 + hl)) * g.W + (x0 + wl)) * g.Cout + co0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                float4* dst = reinterpret_cast<float4*>(yrow + c + j);
                if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                *dst = o;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

template <int BN, int TERMS>
int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmXl, const CUtensorMap& tmWl, float* y, const ConvGeom& g,
                int accumulate, cudaStream_t s) {
    constexpr int CSTAGES = (TERMS == 3) ? 3 : 4;
    constexpr uint32_t kStage = (CBM + BN) * CBK * 2 * (TERMS == 3 ? 2 : 1);
    const size_t smem = 1024 + (size_t)CSTAGES * kStage + 256;
    auto kern = conv_nhwc_bf16_kernel<BN, TERMS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("conv2d_nhwc_bf16: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int64_t grid = (int64_t)g.tiles_n * g.tiles_y * g.tiles_x * g.tiles_co;
    kern<<<(unsigned)grid, kConvThreads, smem, s>>>(tmX, tmW, tmXl, tmWl, y, g, accumulate);
    return 0;
}

}  // namespace tc

static int conv_impl(const void* x, const void* xl, const void* w, const void* wl, float* y, int N, int H, int W, int Cin, int Cout,
                     int ksize, int accumulate, void* stream) {
    GP3D_CHECK_ARG(x && w && y, "conv2d_nhwc_bf16: null pointer");
    GP3D_CHECK_ARG((xl == nullptr) == (wl == nullptr), "conv2d_nhwc_bf16x3: both low-order operands are required");
    GP3D_CHECK_ARG(ksize == 1 || ksize == 3, "conv2d_nhwc_bf16: kernel size must be 1 or 3 (got %d)", ksize);
    GP3D_CHECK_ARG(N > 0 && H > 0 && W > 0, "conv2d_nhwc_bf16: empty tensor");
    if (Cin % 64 != 0 || !(Cout % 128 == 0 || Cout == 96 || Cout == 64)) {
        gp3d_set_error("conv2d_nhwc_bf16: need Cin %% 64 == 0 and Cout %% 128 == 0 (or Cout in {64, 96}); got Cin=%d Cout=%d", Cin, Cout);
        return GP3D_E_UNSUPPORTED;
    }
    tc::ConvGeom g{};
    g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.ks = ksize;
    g.TW = W < 16 ? W : 16;
    g.TH = (128 / g.TW) < H ? (128 / g.TW) : H;
    g.TN = 128 / (g.TW * g.TH);
    if (g.TW * g.TH * g.TN != 128 || W % g.TW || H % g.TH || N % g.TN) {
        gp3d_set_error("conv2d_nhwc_bf16: cannot tile N=%d H=%d W=%d into 128-pixel blocks (W, H powers of two; N %% %d == 0)", N, H, W, g.TN);
        return GP3D_E_UNSUPPORTED;
    }
    const int BN = (Cout % 128 == 0) ? 128 : Cout;
    g.tiles_x = W / g.TW; g.tiles_y = H / g.TH; g.tiles_n = N / g.TN; g.tiles_co = Cout / BN;
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    CUtensorMap tmX, tmW, tmXl, tmWl;
    for (int part = 0; part < (xl ? 2 : 1); part++) {
        const void* xp = part ? xl : x;
        CUtensorMap* tm = part ? &tmXl : &tmX;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TN};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { gp3d_set_error("conv2d_nhwc_bf16: activation tensor map encode failed (CUresult %d)", (int)r); return GP3D_E_BADARG; }
    }
    for (int part = 0; part < (wl ? 2 : 1); part++) {
        const void* wp = part ? wl : w;
        CUtensorMap* tm = part ? &tmWl : &tmW;
        const cuuint64_t Kt = (cuuint64_t)ksize * ksize * Cin;
        cuuint64_t dims[2] = {Kt, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {Kt * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { gp3d_set_error("conv2d_nhwc_bf16: weight tensor map encode failed (CUresult %d)", (int)r); return GP3D_E_BADARG; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (!xl) {
        tmXl = tmX; tmWl = tmW;
        rc = (BN == 128) ? tc::launch_conv<128, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s)
           : (BN == 96)  ? tc::launch_conv<96, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s)
                         : tc::launch_conv<64, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s);
    } else {
        rc = (BN == 128) ? tc::launch_conv<128, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s)
           : (BN == 96)  ? tc::launch_conv<96, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s)
                         : tc::launch_conv<64, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s);
    }
    if (rc) return rc;
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_conv2d_nhwc_bf16(const void* x, const void* w, float* y, int N, int H, int W, int Cin, int Cout,
                                     int ksize, int accumulate, void* stream) {
    return conv_impl(x, nullptr, w, nullptr, y, N, H, W, Cin, Cout, ksize, accumulate, stream);
}

extern "C" int gp3d_conv2d_nhwc_bf16x3(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                                       int Cin, int Cout, int ksize, int accumulate, void* stream) {
    GP3D_CHECK_ARG(xl && wl, "conv2d_nhwc_bf16x3: null low-order operand");
    return conv_impl(xh, xl, wh, wl, y, N, H, W, Cin, Cout, ksize, accumulate, stream);
}
