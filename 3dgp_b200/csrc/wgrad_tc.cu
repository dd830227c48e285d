// Weight gradient of the tap convolutions (conv_tc.cu) as a tcgen05 / TMEM GEMM over pixels (sm_100a).
//
//   dW[co][slab_t][ci] += sum_{n, iy, ix}  dy[n][iy*sa + ay_t][ix*sa + ax_t][co] * x[n][iy*sb + by_t][ix*sb + bx_t][ci]
//
// GEMM view per tap: M = Cout (tile 128), N = Cin (tile 128), K = pixels (blocks of 64).  Both operands are "MN-major":
// the NHWC tensors have the M / N index (channels) contiguous and the K index (pixels) strided, so a k-block is fetched by
// 4-D TMA boxes {64 channels, TW, TH, TN} straight into the canonical MN-major 128B-swizzled layout (rows = pixels, 128 bytes
// of channels per row); out-of-range pixels (padding halo, tile overshoot) arrive as zeros and contribute nothing.
// Split-K: every CTA owns one (cout tile, cin tile, tap) and a strided share of the pixel blocks, accumulates them in TMEM and
// adds its 128 x 128 partial into dW with red.global.add.v4.f32 (dW is zero-filled by the caller).
// TERMS == 3: dy = dh + dl, x = xh + xl (bf16 pairs): dW += dh*xh + dh*xl + dl*xh in the same accumulator.
#include "tc_common.cuh"

namespace tc {

constexpr int WBM = 128, WBK = 64;          // the Cin tile (WBN) is a template parameter: 128, or 256 when Cin % 256 == 0
constexpr int kWgradThreads = 256;

struct WgradGeom {
    int N, Cin, Cout, num_slabs;
    int ntaps;
    int ay[25], ax[25], by[25], bx[25], slab[25];
    int sa, sb;                 // traversal strides of the dy / x tensor maps
    int HoP, WoP;               // pixel domain
    int TW, TH, TN, tiles_x, tiles_y, tiles_n;
    int tiles_co, tiles_ci, splitk;
    int dy_fp16, x_fp16;         // operand element formats (0 bf16, 1 fp16)
};

// Shared-memory descriptor of an MN-major operand tile made of 64-channel (128-byte) column blocks of `rows_k` pixel rows:
//   LBO = byte distance between consecutive 64-channel blocks, SBO = 1024 B (8 pixel rows), 128B swizzle, version 1.
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int TERMS, int WBN>
__global__ void __launch_bounds__(kWgradThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmXh,
             const __grid_constant__ CUtensorMap tmDl, const __grid_constant__ CUtensorMap tmXl,
             float* __restrict__ dW, WgradGeom g) {
    // 256-wide Cin tiles halve the dy operand traffic per MMA (the kernel is bound by the L2 -> SM operand stream, like conv_tc.cu): the three-term
    // form then runs a two-stage ring of 96 KB stages.
    constexpr int STAGES = (TERMS == 3) ? (WBN == 256 ? 2 : 3) : 4;
    constexpr uint32_t kBlock = WBK * 128;                 // one 64-channel column block of 64 pixel rows: 8 KB
    constexpr uint32_t kOperand = 2 * kBlock;              // dy tile: 128 channels
    constexpr uint32_t kOperandB = (WBN / 64) * kBlock;    // x tile: WBN channels
    constexpr uint32_t kPart = kOperand + kOperandB;       // dy tile + x tile
    constexpr uint32_t kStage = kPart * (TERMS == 3 ? 2 : 1);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * kStage);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x;
    const int sk = t % g.splitk; t /= g.splitk;
    const int tap = t % g.ntaps; t /= g.ntaps;
    const int tci = t % g.tiles_ci; t /= g.tiles_ci;
    const int tco = t;
    const int ptiles = g.tiles_n * g.tiles_y * g.tiles_x;
    const int my_tiles = (ptiles - sk + g.splitk - 1) / g.splitk;      // pixel blocks sk, sk + splitk, ...

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmDh); prefetch_tmap(&tmXh); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, WBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < my_tiles; kb++) {
                const int st = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                int pt = sk + kb * g.splitk;
                const int tx = pt % g.tiles_x; pt /= g.tiles_x;
                const int ty = pt % g.tiles_y; pt /= g.tiles_y;
                const int n0 = pt * g.TN, x0 = tx * g.TW, y0 = ty * g.TH;
                const int acx = x0 * g.sa + g.ax[tap], acy = y0 * g.sa + g.ay[tap];
                const int bcx = x0 * g.sb + g.bx[tap], bcy = y0 * g.sb + g.by[tap];
                mbar_wait(&empty_bar[st], ph ^ 1);
                unsigned char* sA = smem + (size_t)st * kStage;
                unsigned char* sB = sA + kOperand;
                mbar_expect_tx(&full_bar[st], kStage);
                tma_load_4d(sA, &tmDh, &full_bar[st], tco * WBM, acx, acy, n0);
                tma_load_4d(sA + kBlock, &tmDh, &full_bar[st], tco * WBM + 64, acx, acy, n0);
#pragma unroll
                for (int cb = 0; cb < WBN / 64; cb++) tma_load_4d(sB + cb * kBlock, &tmXh, &full_bar[st], tci * WBN + cb * 64, bcx, bcy, n0);
                if (TERMS == 3) {
                    tma_load_4d(sA + kPart, &tmDl, &full_bar[st], tco * WBM, acx, acy, n0);
                    tma_load_4d(sA + kPart + kBlock, &tmDl, &full_bar[st], tco * WBM + 64, acx, acy, n0);
#pragma unroll
                    for (int cb = 0; cb < WBN / 64; cb++) tma_load_4d(sB + kPart + cb * kBlock, &tmXl, &full_bar[st], tci * WBN + cb * 64, bcx, bcy, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16kind(WBM, WBN, g.dy_fp16, g.x_fp16, 1, 1);
            for (int kb = 0; kb < my_tiles; kb++) {
                const int st = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t sA = smem_u32(smem + (size_t)st * kStage);
                const uint32_t sB = sA + kOperand;
#pragma unroll
                for (int k = 0; k < WBK / 16; k++) {
                    const uint32_t ko = k * 16 * 128;                // 16 pixel rows per MMA
                    const uint64_t dah = make_desc_mn_sw128(sA + ko, kBlock), dbh = make_desc_mn_sw128(sB + ko, kBlock);
                    umma_bf16(tmem_base, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    if (TERMS == 3) {
                        umma_bf16(tmem_base, dah, make_desc_mn_sw128(sB + kPart + ko, kBlock), idesc, 1u);
                        umma_bf16(tmem_base, make_desc_mn_sw128(sA + kPart + ko, kBlock), dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty_bar[st]);
            }
            umma_commit(accum_bar);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        if (my_tiles > 0) {
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int co = tco * WBM + q * 32 + lane;
            float* row = dW + ((size_t)co * g.num_slabs + g.slab[tap]) * g.Cin + (size_t)tci * WBN;
#pragma unroll 1
            for (int c = 0; c < WBN; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // warp-collective
                tmem_ld_wait();
                if (co >= g.Cout || tci * WBN + c >= g.Cin) continue;                       // half-empty tiles of 64-channel tensors
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    red_add_v4f(row + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, WBN);
}

template <int TERMS, int WBN>
int launch_wgrad(const CUtensorMap& tmDh, const CUtensorMap& tmXh, const CUtensorMap& tmDl, const CUtensorMap& tmXl, float* dW, const WgradGeom& g,
                 int64_t grid, cudaStream_t s, const char* who) {
    constexpr int STAGES = (TERMS == 3) ? (WBN == 256 ? 2 : 3) : 4;
    constexpr size_t stage = (size_t)(TERMS == 3 ? 2 : 1) * (2 + WBN / 64) * 64 * 128;
    const size_t smem = 1024 + STAGES * stage + 256;
    auto kern = wgrad_kernel<TERMS, WBN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("%s: cannot reserve %zu B of shared memory: %s", who, smem, cudaGetErrorString(e)); return (int)e; }
    kern<<<(unsigned)grid, kWgradThreads, smem, s>>>(tmDh, tmXh, tmDl, tmXl, dW, g);
    return 0;
}

}  // namespace tc

extern "C" int gp3d_conv_set_wide3(int on);   // conv_tc.cu: the same switch selects the 256-wide tiles here (query = set + restore)

static int wg_pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int encode_act_map(CUtensorMap* tm, const void* p, int N, int H, int W, int C, int TW, int TH, int TN, int stride, const char* who, int fp16 = 0) {
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(TW * stride), (cuuint32_t)(TH * stride), (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gp3d_set_error("%s: tensor map encode failed (CUresult %d)", who, (int)r); return GP3D_E_BADARG; }
    return 0;
}

extern "C" int gp3d_wgrad_taps_nhwc(const void* dyh, const void* dyl, const void* xh, const void* xl, float* dW,
                                    int N, int Hd, int Wd, int Cout, int Hx, int Wx, int Cin, int num_slabs,
                                    int ntaps, const int* h_taps, int sa, int sb, int HoP, int WoP, void* stream) {
    return gp3d_wgrad_taps_nhwc_fmt(dyh, dyl, xh, xl, 0, 0, dW, N, Hd, Wd, Cout, Hx, Wx, Cin, num_slabs, ntaps, h_taps, sa, sb, HoP, WoP, stream);
}

extern "C" int gp3d_wgrad_taps_nhwc_fmt(const void* dyh, const void* dyl, const void* xh, const void* xl, int dy_format, int x_format, float* dW,
                                        int N, int Hd, int Wd, int Cout, int Hx, int Wx, int Cin, int num_slabs,
                                        int ntaps, const int* h_taps, int sa, int sb, int HoP, int WoP, void* stream) {
    const char* who = "wgrad_taps_nhwc";
    GP3D_CHECK_ARG(dyh && xh && dW && h_taps, "%s: null pointer", who);
    GP3D_CHECK_ARG((dy_format == 0 || dy_format == 1) && (x_format == 0 || x_format == 1), "%s: operand formats are 0 (bf16) or 1 (fp16)", who);
    GP3D_CHECK_ARG(dy_format == x_format, "%s: both operands of a tcgen05 kind::f16 product must have the same element format (got dy %d, x %d)", who, dy_format, x_format);
    GP3D_CHECK_ARG(dyl == nullptr || (dy_format == 0 && x_format == 0), "%s: the three-term form takes bf16 pairs", who);
    GP3D_CHECK_ARG((dyl == nullptr) == (xl == nullptr), "%s: both low-order operands are required", who);
    GP3D_CHECK_ARG(ntaps >= 1 && ntaps <= 25 && (sa == 1 || sa == 2) && (sb == 1 || sb == 2), "%s: bad tap list / strides", who);
    GP3D_CHECK_ARG(N > 0 && HoP > 0 && WoP > 0, "%s: empty domain", who);
    if (Cin % 64 != 0 || Cout % 64 != 0) {      // 64-channel tensors use half of a 128-wide tile: the missing block is TMA zero fill
        gp3d_set_error("%s: need Cin %% 64 == 0 and Cout %% 64 == 0 (got Cin=%d Cout=%d)", who, Cin, Cout);
        return GP3D_E_UNSUPPORTED;
    }
    tc::WgradGeom g{};
    g.N = N; g.Cin = Cin; g.Cout = Cout; g.num_slabs = num_slabs; g.ntaps = ntaps; g.sa = sa; g.sb = sb; g.HoP = HoP; g.WoP = WoP;
    g.dy_fp16 = dy_format; g.x_fp16 = x_format;
    for (int t = 0; t < ntaps; t++) {
        g.ay[t] = h_taps[5 * t]; g.ax[t] = h_taps[5 * t + 1]; g.by[t] = h_taps[5 * t + 2]; g.bx[t] = h_taps[5 * t + 3]; g.slab[t] = h_taps[5 * t + 4];
        GP3D_CHECK_ARG(g.slab[t] >= 0 && g.slab[t] < num_slabs, "%s: weight slab out of range", who);
    }
    g.TW = wg_pow2ceil(WoP) < 16 ? wg_pow2ceil(WoP) : 16;
    g.TH = wg_pow2ceil(HoP) < (64 / g.TW) ? wg_pow2ceil(HoP) : (64 / g.TW);
    g.TN = 64 / (g.TW * g.TH);
    g.tiles_x = (WoP + g.TW - 1) / g.TW; g.tiles_y = (HoP + g.TH - 1) / g.TH; g.tiles_n = (N + g.TN - 1) / g.TN;
    int wide_on = gp3d_conv_set_wide3(1); gp3d_conv_set_wide3(wide_on);
    const int WBN = (Cin % 256 == 0 && wide_on) ? 256 : 128;
    g.tiles_co = (Cout + 127) / 128; g.tiles_ci = (Cin + WBN - 1) / WBN;
    const int64_t ptiles = (int64_t)g.tiles_x * g.tiles_y * g.tiles_n;
    const int out_tiles = g.tiles_co * g.tiles_ci * ntaps;
    int sms = GP3D_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // One CTA per SM (197 KB of shared memory): the grid runs in waves of `sms` CTAs and a partial last wave costs a full CTA time.  Choose the split
    // that minimises waves x (pixel blocks per CTA) -- e.g. 18 output tiles: split 16 (288 CTAs, two full waves) instead of 17 (306 CTAs: a third wave
    // of 10 CTAs, +50 % time; ncu: tensor pipe 82 % while active but 55 % of the elapsed cycles, profiles/r2_conv_wide_tiles_ncu.txt).
    // At most two waves: more CTAs only add red.add traffic (a 7-wave split of the 128-channel 512^2 layer measured 4 % slower than the old 3-wave one).
    int64_t splitk = 1, best = -1;
    const int64_t smax = ptiles < 2 * (int64_t)sms ? ptiles : 2 * (int64_t)sms;
    for (int64_t sk = 1; sk <= smax; sk++) {
        const int64_t waves = (out_tiles * sk + sms - 1) / sms;
        if (waves > 2 && sk > 1) break;
        const int64_t per_cta = (ptiles + sk - 1) / sk;
        const int64_t cost = waves * (per_cta + 6);      // pixel blocks per CTA + pipeline fill / 128 x WBN red.add epilogue (~6 block times) per wave
        if (best < 0 || cost < best) { best = cost; splitk = sk; }
    }
    g.splitk = (int)splitk;
    CUtensorMap tmDh, tmXh, tmDl, tmXl;
    int rc = encode_act_map(&tmDh, dyh, N, Hd, Wd, Cout, g.TW, g.TH, g.TN, sa, who, dy_format); if (rc) return rc;
    rc = encode_act_map(&tmXh, xh, N, Hx, Wx, Cin, g.TW, g.TH, g.TN, sb, who, x_format); if (rc) return rc;
    if (dyl) {
        rc = encode_act_map(&tmDl, dyl, N, Hd, Wd, Cout, g.TW, g.TH, g.TN, sa, who); if (rc) return rc;
        rc = encode_act_map(&tmXl, xl, N, Hx, Wx, Cin, g.TW, g.TH, g.TN, sb, who); if (rc) return rc;
    } else { tmDl = tmDh; tmXl = tmXh; }
    const int terms = dyl ? 3 : 1;
    const int64_t grid = (int64_t)out_tiles * g.splitk;
    cudaStream_t s = (cudaStream_t)stream;
    rc = (terms == 3) ? (WBN == 256 ? tc::launch_wgrad<3, 256>(tmDh, tmXh, tmDl, tmXl, dW, g, grid, s, who) : tc::launch_wgrad<3, 128>(tmDh, tmXh, tmDl, tmXl, dW, g, grid, s, who))
                      : (WBN == 256 ? tc::launch_wgrad<1, 256>(tmDh, tmXh, tmDl, tmXl, dW, g, grid, s, who) : tc::launch_wgrad<1, 128>(tmDh, tmXh, tmDl, tmXl, dW, g, grid, s, who));
    if (rc) return rc;
    GP3D_RETURN_LAUNCH();
}
