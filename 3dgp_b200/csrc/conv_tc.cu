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
// 256 threads, warp-specialised: warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one lane: tcgen05.mma into TMEM, tcgen05.commit frees
// the stage), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32 lanes x 32 columns per warp).
#include "tc_common.cuh"

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency); shared by every TMA user of the library.
gp3d_encode_tiled_fn gp3d_get_encode_tiled() {
    static gp3d_encode_tiled_fn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = reinterpret_cast<gp3d_encode_tiled_fn>(p);
    return fn;
}

namespace tc {

constexpr int CBM = 128, CBK = 64;
constexpr int kConvThreads = 256;

struct ConvGeom {
    int N, H, W, Cin, Cout;      // input tensor [N][H][W][Cin]; weights [Cout][num_slabs][Cin]
    int ntaps;                   // filter taps of THIS launch
    int tdy[25], tdx[25], tslab[25];  // input offset (in input pixels) and weight slab of each tap (up to 5x5)
    int in_stride;               // 1, or 2 for strided gathers (the tensor map carries matching elementStrides)
    int HoP, WoP;                // logical output grid of this launch (tile domain); pixel (iy, ix) reads input (iy*in_stride+tdy, ix*in_stride+tdx)
    int Hout, Wout;              // output tensor [N][Hout][Wout][Cout]; pixel (iy, ix) is stored at (iy*osy + oy0, ix*osx + ox0)
    int osy, osx, oy0, ox0;
    int TW, TH, TN;              // pixel tile: TN images x TH rows x TW cols = 128
    int tiles_x, tiles_y, tiles_n, tiles_co;
    int x_fp16, w_fp16;          // operand element formats (0 bf16, 1 fp16)
    // Phases: up to four tap sub-lists with their own output lattice offset and domain, processed by ONE launch (the polyphase form of the stride-2
    // transposed convolution: the four launches of one layer re-read the activation tensor from DRAM four times; as phases of one launch the tiles of
    // one pixel window are scheduled back to back and the re-reads hit L2).  nphases == 1: the plain form (phase 0 = the fields above).
    int nphases;
    int ph_ntaps[4], ph_tap0[4], ph_oy0[4], ph_ox0[4], ph_HoP[4], ph_WoP[4];
};

// Optional fused epilogue of the modulated-conv layer (networks_stylegan2.py:71 fma + :144 bias_act, no clamp):
//   y = act(acc * dcoef[n][co] + noise[n?][oy][ox] + bias[co]) * gain      -- the raw conv output never reaches HBM.
struct ConvEpi {
    const float* dcoef; const float* noise; const float* bias;
    int enabled, noise_per_sample, act;
    float alpha, gain, clamp;      // clamp <= 0: none
};

// TERMS == 1: y += xh * wh.   TERMS == 3 (error-compensated "bf16x3", ~2^-16 relative): y += xh*wh + xh*wl + xl*wh with
// x = xh + xl, w = wh + wl split into bf16 pairs; all three products accumulate into the same TMEM tile.
// TERMS == 2 ("x2w16"): y += xh*w16 + xl*w16 -- the activation is an fp16 PAIR (22 mantissa bits; forward activations are O(1), far inside fp16's
// range), the weight ONE fp16 operand (11 bits, relative rounding 2^-12).  Both operands of one tcgen05.mma.kind::f16 must have the SAME element
// format: mixed bf16 x fp16 instruction descriptors raise an illegal-instruction fault on sm_100a (measured, tools/probe_formats.py).
template <int BN, int TERMS>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_nhwc_bf16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmXl, const __grid_constant__ CUtensorMap tmWl,
                      float* __restrict__ Y, ConvGeom g, int accumulate, int num_tiles, ConvEpi ep) {
    // Persistent: one CTA per SM loops over output tiles.  Two TMEM accumulator buffers (2 x 128 columns) let the epilogue of
    // tile i overlap the TMA / MMA main loop of tile i+1; the smem ring and its phases run continuously across tiles.
    constexpr int CSTAGES = (TERMS == 3) ? (BN == 256 ? 2 : 3) : (TERMS == 2 && BN == 256) ? 3 : 4;
    constexpr uint32_t kOperand = (CBM + BN) * CBK * 2;
    constexpr uint32_t kStage = kOperand + (TERMS == 3 ? kOperand : TERMS == 2 ? CBM * CBK * 2 : 0);     // [A_hi | B_hi] [A_lo] [B_lo]
    constexpr uint32_t kAccStride = (BN <= 128) ? 128 : 256;      // TMEM columns per accumulator buffer (two buffers: 256 or all 512 columns)
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tiles = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)CSTAGES * kStage);
    uint64_t* empty_bar = full_bar + CSTAGES;
    uint64_t* acc_full = empty_bar + CSTAGES;       // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;             // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    // epilogue staging: 4 warps x [32 pixel rows][36 floats] (pitch 36: 16-byte aligned rows, conflict-free STS.128 / LDS.128)
    float* stg_all = reinterpret_cast<float*>(smem + (size_t)CSTAGES * kStage + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cblocks = g.Cin / CBK;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < CSTAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }   // 4 epilogue warps arrive
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 2 * kAccStride);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile decomposition: tile = (((tn * tiles_y + ty) * tiles_x + tx) * nphases + phase') * tiles_co + tco   (cout, then phase fastest: activations
    // reused from L2).  The phase is rotated with the pixel-tile index so that a persistent CTA (stride gridDim.x, a multiple of 4) does not always draw
    // the same phase -- the phases have 1, 2, 2 and 4 taps.
    auto decode = [&](int t, int& x0, int& y0, int& n0, int& co0, int& ph) {
        const int tco = t % g.tiles_co; t /= g.tiles_co;
        ph = 0;
        if (g.nphases > 1) { const int q = t; t /= g.nphases; ph = (q + t / 37) % g.nphases; }
        const int tx = t % g.tiles_x; t /= g.tiles_x;
        const int ty = t % g.tiles_y; t /= g.tiles_y;
        x0 = tx * g.TW; y0 = ty * g.TH; n0 = t * g.TN; co0 = tco * BN;
    };

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                                  // global k-block counter (ring position)
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int x0, y0, n0, co0, phs;
                decode(tile, x0, y0, n0, co0, phs);
                const int num_kb = g.ph_ntaps[phs] * cblocks, tap0 = g.ph_tap0[phs];
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int st = it % CSTAGES; const uint32_t ph = (it / CSTAGES) & 1;
                    const int tl = kb / cblocks, cb = kb - tl * cblocks, tap = tap0 + tl;
                    const int cx = x0 * g.in_stride + g.tdx[tap], cy = y0 * g.in_stride + g.tdy[tap], slab = g.tslab[tap];
                    mbar_wait(&empty_bar[st], ph ^ 1);
                    unsigned char* sa = tiles + (size_t)st * kStage;
                    unsigned char* sb = sa + CBM * CBK * 2;
                    mbar_expect_tx(&full_bar[st], kStage);
                    tma_load_4d(sa, &tmX, &full_bar[st], cb * CBK, cx, cy, n0);
                    tma_load_2d(sb, &tmW, &full_bar[st], slab * g.Cin + cb * CBK, co0);
                    if (TERMS >= 2) tma_load_4d(sa + kOperand, &tmXl, &full_bar[st], cb * CBK, cx, cy, n0);
                    if (TERMS == 3) tma_load_2d(sb + kOperand, &tmWl, &full_bar[st], slab * g.Cin + cb * CBK, co0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16kind(CBM, BN, g.x_fp16, g.w_fp16);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
                const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
                int x0_, y0_, n0_, co0_, phs;
                decode(tile, x0_, y0_, n0_, co0_, phs);
                const int num_kb = g.ph_ntaps[phs] * cblocks;
                mbar_wait(&acc_empty[buf], aph ^ 1);          // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * kAccStride;
                for (int kb = 0; kb < num_kb; kb++, it++) {
                    const int st = it % CSTAGES; const uint32_t ph = (it / CSTAGES) & 1;
                    mbar_wait(&full_bar[st], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + (size_t)st * kStage);
                    const uint32_t sb = sa + CBM * CBK * 2;
#pragma unroll
                    for (int k = 0; k < CBK / 16; k++) {
                        const uint64_t dah = make_desc_k_sw128(sa + k * 32), dbh = make_desc_k_sw128(sb + k * 32);
                        umma_bf16(tmem_d, dah, dbh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        if (TERMS == 3) umma_bf16(tmem_d, dah, make_desc_k_sw128(sb + kOperand + k * 32), idesc, 1u);
                        if (TERMS >= 2) umma_bf16(tmem_d, make_desc_k_sw128(sa + kOperand + k * 32), dbh, idesc, 1u);
                    }
                    umma_commit(&empty_bar[st]);
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, tcount++) {
            const uint32_t buf = tcount & 1, aph = (tcount >> 1) & 1;
            int x0, y0, n0, co0, phs;
            decode(tile, x0, y0, n0, co0, phs);
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const int oy0 = g.ph_oy0[phs], ox0 = g.ph_ox0[phs];
            if (g.TN == 1) {
                // ---- coalesced epilogue (one image per tile).  A TMEM row is a pixel, so lane i holds 32 channels of pixel i: stored directly, one
                // instruction scatters 16 bytes into 32 different 128-byte lines (measured: the stores cost 18 % of the 128-channel 512^2 layers and half of
                // toRGB, profiles/r2_conv_epilogue_stores.txt).  The 32 x 32 chunk is transposed through shared memory instead: 8 lanes write the 128
                // contiguous bytes of one pixel, a warp instruction covers 4 whole lines.
                float* stg = stg_all + (size_t)q * 32 * 36;
                const int cg = lane & 7;                                   // 4-channel group of this lane inside a 32-channel chunk
                int off[8]; uint32_t in_mask = 0; float nzv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int r = q * 32 + 4 * i + (lane >> 3);            // pixel of this lane in store instruction i
                    const int wl = r % g.TW, hl = r / g.TW;
                    const int ix = x0 + wl, iy = y0 + hl;
                    const bool in = (ix < g.ph_WoP[phs]) && (iy < g.ph_HoP[phs]) && (n0 < g.N);
                    in_mask |= (in ? 1u : 0u) << i;
                    off[i] = ((hl * g.osy) * g.Wout + wl * g.osx) * g.Cout;
                    nzv[i] = 0.f;
                    if (ep.enabled && ep.noise && in)
                        nzv[i] = ep.noise[(ep.noise_per_sample ? (size_t)n0 * g.Hout * g.Wout : 0) + (size_t)(iy * g.osy + oy0) * g.Wout + (size_t)(ix * g.osx + ox0)];
                }
                float* ybase = Y + (((size_t)n0 * g.Hout + (size_t)(y0 * g.osy + oy0)) * g.Wout + (size_t)(x0 * g.osx + ox0)) * g.Cout + co0;
                const float* drow = (ep.enabled && ep.dcoef && n0 < g.N) ? ep.dcoef + (size_t)n0 * g.Cout + co0 : nullptr;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + buf * kAccStride + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // warp-collective
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<uint4*>(stg + lane * 36 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                    float4 dv = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ep.enabled) {
                        if (drow) dv = *reinterpret_cast<const float4*>(drow + c + 4 * cg);
                        if (ep.bias) bv = *reinterpret_cast<const float4*>(ep.bias + co0 + c + 4 * cg);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        float4 o = *reinterpret_cast<const float4*>(stg + (4 * i + (lane >> 3)) * 36 + 4 * cg);
                        if (!((in_mask >> i) & 1u)) continue;
                        if (ep.enabled) {
                            const float nz = nzv[i];
                            o.x = fmaf(o.x, dv.x, nz) + bv.x; o.y = fmaf(o.y, dv.y, nz) + bv.y; o.z = fmaf(o.z, dv.z, nz) + bv.z; o.w = fmaf(o.w, dv.w, nz) + bv.w;
                            if (ep.act == 3) {
                                o.x = (o.x > 0.f) ? o.x : o.x * ep.alpha; o.y = (o.y > 0.f) ? o.y : o.y * ep.alpha;
                                o.z = (o.z > 0.f) ? o.z : o.z * ep.alpha; o.w = (o.w > 0.f) ? o.w : o.w * ep.alpha;
                            }
                            o.x *= ep.gain; o.y *= ep.gain; o.z *= ep.gain; o.w *= ep.gain;
                            if (ep.clamp > 0.f) {
                                o.x = fminf(fmaxf(o.x, -ep.clamp), ep.clamp); o.y = fminf(fmaxf(o.y, -ep.clamp), ep.clamp);
                                o.z = fminf(fmaxf(o.z, -ep.clamp), ep.clamp); o.w = fminf(fmaxf(o.w, -ep.clamp), ep.clamp);
                            }
                        }
                        float4* dst = reinterpret_cast<float4*>(ybase + off[i] + c + 4 * cg);
                        if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                        *dst = o;
                    }
                    __syncwarp();                                          // the staging rows are rewritten by the next chunk
                }
            } else {
            // ---- several images per tile (tensors smaller than 128 pixels per image): a lane stores its own pixel row
            const int r = q * 32 + lane;                       // pixel index inside the tile: ((n*TH + h)*TW + w)
            const int wl = r % g.TW, hl = (r / g.TW) % g.TH, nl = r / (g.TW * g.TH);
            const int ix = x0 + wl, iy = y0 + hl, nn = n0 + nl;
            const bool inside = (ix < g.ph_WoP[phs]) && (iy < g.ph_HoP[phs]) && (nn < g.N);
            float* yrow = Y + (((size_t)nn * g.Hout + (size_t)(iy * g.osy + oy0)) * g.Wout + (size_t)(ix * g.osx + ox0)) * g.Cout + co0;
            float nz = 0.f;
            const float* drow = nullptr;
            if (ep.enabled && inside) {
                if (ep.noise) nz = ep.noise[(ep.noise_per_sample ? (size_t)nn * g.Hout * g.Wout : 0) + (size_t)(iy * g.osy + oy0) * g.Wout + (size_t)(ix * g.osx + ox0)];
                if (ep.dcoef) drow = ep.dcoef + (size_t)nn * g.Cout + co0;
            }
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + buf * kAccStride + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);   // warp-collective
                tmem_ld_wait();
                if (!inside) continue;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                    if (ep.enabled) {
                        const float4 dv = drow ? *reinterpret_cast<const float4*>(drow + c + j) : make_float4(1.f, 1.f, 1.f, 1.f);
                        const float4 bv = ep.bias ? *reinterpret_cast<const float4*>(ep.bias + co0 + c + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        o.x = fmaf(o.x, dv.x, nz) + bv.x; o.y = fmaf(o.y, dv.y, nz) + bv.y; o.z = fmaf(o.z, dv.z, nz) + bv.z; o.w = fmaf(o.w, dv.w, nz) + bv.w;
                        if (ep.act == 3) {
                            o.x = (o.x > 0.f) ? o.x : o.x * ep.alpha; o.y = (o.y > 0.f) ? o.y : o.y * ep.alpha;
                            o.z = (o.z > 0.f) ? o.z : o.z * ep.alpha; o.w = (o.w > 0.f) ? o.w : o.w * ep.alpha;
                        }
                        o.x *= ep.gain; o.y *= ep.gain; o.z *= ep.gain; o.w *= ep.gain;
                        if (ep.clamp > 0.f) {
                            o.x = fminf(fmaxf(o.x, -ep.clamp), ep.clamp); o.y = fminf(fmaxf(o.y, -ep.clamp), ep.clamp);
                            o.z = fminf(fmaxf(o.z, -ep.clamp), ep.clamp); o.w = fminf(fmaxf(o.w, -ep.clamp), ep.clamp);
                        }
                    }
                    float4* dst = reinterpret_cast<float4*>(yrow + c + j);
                    if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                    *dst = o;
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);       // this warp's quarter of the buffer is free again
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kAccStride);
}

template <int BN, int TERMS>
int launch_conv(const CUtensorMap& tmX, const CUtensorMap& tmW, const CUtensorMap& tmXl, const CUtensorMap& tmWl, float* y, const ConvGeom& g,
                int accumulate, cudaStream_t s, const ConvEpi& ep) {
    constexpr int CSTAGES = (TERMS == 3) ? (BN == 256 ? 2 : 3) : (TERMS == 2 && BN == 256) ? 3 : 4;
    constexpr uint32_t kStage = (CBM + BN) * CBK * 2 * (TERMS == 3 ? 2 : 1) + (TERMS == 2 ? CBM * CBK * 2 : 0);
    const size_t smem = 1024 + (size_t)CSTAGES * kStage + 256 + 4 * 32 * 36 * sizeof(float);      // ring + barriers + epilogue staging
    auto kern = conv_nhwc_bf16_kernel<BN, TERMS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("conv2d_nhwc_bf16: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int64_t num_tiles = (int64_t)g.tiles_n * g.tiles_y * g.tiles_x * g.tiles_co * g.nphases;
    int sms = GP3D_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t grid = num_tiles < sms ? num_tiles : sms;        // persistent: one CTA per SM
    kern<<<(unsigned)grid, kConvThreads, smem, s>>>(tmX, tmW, tmXl, tmWl, y, g, accumulate, (int)num_tiles, ep);
    return 0;
}

}  // namespace tc

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// 256-wide tiles for the three-term form (tuning switch: gp3d_conv_set_wide3; on by default when measured faster)
static int g_conv_wide3 = 1;
extern "C" int gp3d_conv_set_wide3(int on) { const int old = g_conv_wide3; g_conv_wide3 = on ? 1 : 0; return old; }

// General tap convolution.  taps: ntaps x (dy, dx, slab).  num_slabs = weight slabs per output channel.
static int conv_impl(const void* x, const void* xl, const void* w, const void* wl, float* y, int N, int H, int W, int Cin, int Cout,
                     int num_slabs, int ntaps, const int* taps, int in_stride, int HoP, int WoP, int Hout, int Wout,
                     int osy, int osx, int oy0, int ox0, int accumulate, void* stream, const char* who, const gp3d_conv_epilogue* epi = nullptr,
                     int w_format = 0, int x_format = 0, int nphases = 1, const int* phases = nullptr) {
    // phases (nphases > 1): nphases x (ntaps, first tap, oy0, ox0, HoP, WoP); the taps of all phases are concatenated in `taps`, HoP / WoP passed to this
    // function are the LARGEST phase domain (the tile grid), oy0 / ox0 are ignored
    GP3D_CHECK_ARG(x && w && y, "%s: null pointer", who);
    GP3D_CHECK_ARG((w_format == 0 || w_format == 1) && (x_format == 0 || x_format == 1), "%s: operand formats are 0 (bf16) or 1 (fp16)", who);
    GP3D_CHECK_ARG(w_format == x_format, "%s: both operands of a tcgen05 kind::f16 product must have the same element format (got x %d, w %d)", who, x_format, w_format);
    GP3D_CHECK_ARG(w_format == 0 || wl == nullptr, "%s: fp16 weights have no low-order half", who);
    GP3D_CHECK_ARG(!(xl != nullptr && wl == nullptr) || w_format == 1, "%s: an activation pair with a single weight operand is the fp16 two-term form", who);
    tc::ConvEpi ep{};
    if (epi) {
        GP3D_CHECK_ARG(!accumulate, "%s: the fused epilogue cannot accumulate into y", who);
        GP3D_CHECK_ARG(epi->act == 1 || epi->act == 3, "%s: fused epilogue activation must be linear (1) or lrelu (3), got %d", who, epi->act);
        GP3D_CHECK_ARG((!epi->dcoef || gp3d_aligned16(epi->dcoef)) && (!epi->bias || gp3d_aligned16(epi->bias)), "%s: epilogue vectors must be 16-byte aligned", who);
        ep.dcoef = epi->dcoef; ep.noise = epi->noise; ep.bias = epi->bias; ep.enabled = 1; ep.noise_per_sample = epi->noise_per_sample;
        ep.act = epi->act; ep.alpha = epi->alpha; ep.gain = epi->gain; ep.clamp = epi->clamp;
    }
    GP3D_CHECK_ARG(!(wl != nullptr && xl == nullptr), "%s: a low-order weight half needs the low-order activation half", who);
    GP3D_CHECK_ARG(N > 0 && H > 0 && W > 0 && HoP > 0 && WoP > 0, "%s: empty tensor", who);
    GP3D_CHECK_ARG(ntaps >= 1 && ntaps <= 25 && (in_stride == 1 || in_stride == 2), "%s: bad tap list / stride", who);
    if (Cin % 64 != 0 || !(Cout % 128 == 0 || Cout == 96 || Cout == 64)) {
        gp3d_set_error("%s: need Cin %% 64 == 0 and Cout %% 128 == 0 (or Cout in {64, 96}); got Cin=%d Cout=%d", who, Cin, Cout);
        return GP3D_E_UNSUPPORTED;
    }
    GP3D_CHECK_ARG(nphases >= 1 && nphases <= 4 && (nphases == 1 || phases != nullptr), "%s: 1 .. 4 phases", who);
    GP3D_CHECK_ARG(nphases > 1 || ((HoP - 1) * osy + oy0 < Hout && (WoP - 1) * osx + ox0 < Wout), "%s: output lattice exceeds the output tensor", who);
    tc::ConvGeom g{};
    g.nphases = nphases;
    if (nphases == 1) { g.ph_ntaps[0] = ntaps; g.ph_tap0[0] = 0; g.ph_oy0[0] = oy0; g.ph_ox0[0] = ox0; g.ph_HoP[0] = HoP; g.ph_WoP[0] = WoP; }
    else {
        int covered = 0;
        for (int i = 0; i < nphases; i++) {
            g.ph_ntaps[i] = phases[6 * i]; g.ph_tap0[i] = phases[6 * i + 1]; g.ph_oy0[i] = phases[6 * i + 2]; g.ph_ox0[i] = phases[6 * i + 3];
            g.ph_HoP[i] = phases[6 * i + 4]; g.ph_WoP[i] = phases[6 * i + 5];
            GP3D_CHECK_ARG(g.ph_ntaps[i] >= 1 && g.ph_tap0[i] == covered && g.ph_HoP[i] >= 1 && g.ph_HoP[i] <= HoP && g.ph_WoP[i] >= 1 && g.ph_WoP[i] <= WoP,
                           "%s: bad phase %d", who, i);
            GP3D_CHECK_ARG((g.ph_HoP[i] - 1) * osy + g.ph_oy0[i] < Hout && (g.ph_WoP[i] - 1) * osx + g.ph_ox0[i] < Wout, "%s: phase %d exceeds the output tensor", who, i);
            covered += g.ph_ntaps[i];
        }
        GP3D_CHECK_ARG(covered == ntaps, "%s: the phases must partition the tap list", who);
    }
    g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.ntaps = ntaps; g.in_stride = in_stride;
    g.x_fp16 = x_format; g.w_fp16 = w_format;
    for (int t = 0; t < ntaps; t++) { g.tdy[t] = taps[3 * t]; g.tdx[t] = taps[3 * t + 1]; g.tslab[t] = taps[3 * t + 2];
        GP3D_CHECK_ARG(g.tslab[t] >= 0 && g.tslab[t] < num_slabs, "%s: weight slab out of range", who); }
    g.HoP = HoP; g.WoP = WoP; g.Hout = Hout; g.Wout = Wout; g.osy = osy; g.osx = osx; g.oy0 = oy0; g.ox0 = ox0;
    g.TW = pow2ceil(WoP) < 16 ? pow2ceil(WoP) : 16;
    g.TH = pow2ceil(HoP) < (128 / g.TW) ? pow2ceil(HoP) : (128 / g.TW);
    g.TN = 128 / (g.TW * g.TH);
    // launches are bound by the L2 -> SM operand stream (single-term: 96 B/clk/SM at 128x128 tiles; three-term: ~48 B/clk/SM): 256-wide tiles reuse
    // each activation tile for twice the MMAs.  The three-term form then runs a two-stage ring of 96 KB stages (24 MMAs of 128x256x16 each).
    const bool wide3 = xl && wl && Cout % 256 == 0 && g_conv_wide3;
    const int BN = ((!xl || wide3) && Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : Cout;
    g.tiles_x = (WoP + g.TW - 1) / g.TW; g.tiles_y = (HoP + g.TH - 1) / g.TH; g.tiles_n = (N + g.TN - 1) / g.TN; g.tiles_co = Cout / BN;
    GP3D_CHECK_ARG((int64_t)g.tiles_x * g.tiles_y * g.tiles_n * g.tiles_co * nphases < 2147483647LL, "%s: grid too large", who);
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    CUtensorMap tmX, tmW, tmXl, tmWl;
    for (int part = 0; part < (xl ? 2 : 1); part++) {
        const void* xp = part ? xl : x;
        CUtensorMap* tm = part ? &tmXl : &tmX;
        cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(g.TW * in_stride), (cuuint32_t)(g.TH * in_stride), (cuuint32_t)g.TN};
        cuuint32_t estr[4] = {1, (cuuint32_t)in_stride, (cuuint32_t)in_stride, 1};
        CUresult r = enc(tm, x_format == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { gp3d_set_error("%s: activation tensor map encode failed (CUresult %d)", who, (int)r); return GP3D_E_BADARG; }
    }
    for (int part = 0; part < (wl ? 2 : 1); part++) {
        const void* wp = part ? wl : w;
        CUtensorMap* tm = part ? &tmWl : &tmW;
        const cuuint64_t Kt = (cuuint64_t)num_slabs * Cin;
        cuuint64_t dims[2] = {Kt, (cuuint64_t)Cout};
        cuuint64_t strides[1] = {Kt * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(tm, w_format == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { gp3d_set_error("%s: weight tensor map encode failed (CUresult %d)", who, (int)r); return GP3D_E_BADARG; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (!xl) {
        tmXl = tmX; tmWl = tmW;
        rc = (BN == 256) ? tc::launch_conv<256, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
           : (BN == 128) ? tc::launch_conv<128, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
           : (BN == 96)  ? tc::launch_conv<96, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
                         : tc::launch_conv<64, 1>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep);
    } else if (!wl) {
        tmWl = tmW;
        rc = (BN == 128) ? tc::launch_conv<128, 2>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
           : (BN == 96)  ? tc::launch_conv<96, 2>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
                         : tc::launch_conv<64, 2>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep);
    } else {
        rc = (BN == 256) ? tc::launch_conv<256, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
           : (BN == 128) ? tc::launch_conv<128, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
           : (BN == 96)  ? tc::launch_conv<96, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep)
                         : tc::launch_conv<64, 3>(tmX, tmW, tmXl, tmWl, y, g, accumulate, s, ep);
    }
    if (rc) return rc;
    GP3D_RETURN_LAUNCH();
}

static int conv_same(const void* x, const void* xl, const void* w, const void* wl, float* y, int N, int H, int W, int Cin, int Cout,
                     int ksize, int accumulate, void* stream, const char* who, const gp3d_conv_epilogue* epi = nullptr) {
    GP3D_CHECK_ARG(ksize == 1 || ksize == 3 || ksize == 5, "%s: kernel size must be 1, 3 or 5 (got %d)", who, ksize);
    int taps[75]; int nt = 0;
    for (int ky = 0; ky < ksize; ky++) for (int kx = 0; kx < ksize; kx++) { taps[3 * nt] = ky - ksize / 2; taps[3 * nt + 1] = kx - ksize / 2; taps[3 * nt + 2] = ky * ksize + kx; nt++; }
    return conv_impl(x, xl, w, wl, y, N, H, W, Cin, Cout, ksize * ksize, nt, taps, 1, H, W, H, W, 1, 1, 0, 0, accumulate, stream, who, epi);
}

extern "C" int gp3d_conv2d_nhwc_bf16(const void* x, const void* w, float* y, int N, int H, int W, int Cin, int Cout,
                                     int ksize, int accumulate, void* stream) {
    return conv_same(x, nullptr, w, nullptr, y, N, H, W, Cin, Cout, ksize, accumulate, stream, "conv2d_nhwc_bf16");
}

extern "C" int gp3d_conv2d_nhwc_bf16x3(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                                       int Cin, int Cout, int ksize, int accumulate, void* stream) {
    GP3D_CHECK_ARG(xl && wl, "conv2d_nhwc_bf16x3: null low-order operand");
    return conv_same(xh, xl, wh, wl, y, N, H, W, Cin, Cout, ksize, accumulate, stream, "conv2d_nhwc_bf16x3");
}

extern "C" int gp3d_conv2d_nhwc_bf16x3_act(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                                           int Cin, int Cout, int ksize, const gp3d_conv_epilogue* epi, void* stream) {
    GP3D_CHECK_ARG(xl && wl, "conv2d_nhwc_bf16x3_act: null low-order operand");
    GP3D_CHECK_ARG(epi != nullptr, "conv2d_nhwc_bf16x3_act: null epilogue description");
    return conv_same(xh, xl, wh, wl, y, N, H, W, Cin, Cout, ksize, 0, stream, "conv2d_nhwc_bf16x3_act", epi);
}

extern "C" int gp3d_conv2d_nhwc_act(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                                    int Cin, int Cout, int ksize, const gp3d_conv_epilogue* epi, void* stream) {
    GP3D_CHECK_ARG(epi != nullptr, "conv2d_nhwc_act: null epilogue description");
    return conv_same(xh, xl, wh, wl, y, N, H, W, Cin, Cout, ksize, 0, stream, "conv2d_nhwc_act", epi);
}

extern "C" int gp3d_conv_taps_nhwc(const void* xh, const void* xl, const void* wh, const void* wl, float* y,
                                   int N, int H, int W, int Cin, int Cout, int num_slabs, int ntaps, const int* h_taps, int in_stride,
                                   int HoP, int WoP, int Hout, int Wout, int osy, int osx, int oy0, int ox0, int accumulate, void* stream) {
    GP3D_CHECK_ARG(h_taps != nullptr, "conv_taps_nhwc: null tap list");
    return conv_impl(xh, xl, wh, wl, y, N, H, W, Cin, Cout, num_slabs, ntaps, h_taps, in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0,
                     accumulate, stream, "conv_taps_nhwc");
}

// One descriptor-based entry point covering every form above (and the two-term bf16-pair x fp16-weight form).
extern "C" int gp3d_conv_nhwc(const gp3d_conv_desc* d, void* stream) {
    GP3D_CHECK_ARG(d != nullptr, "conv_nhwc: null descriptor");
    GP3D_CHECK_ARG(d->taps != nullptr, "conv_nhwc: null tap list");
    return conv_impl(d->xh, d->xl, d->wh, d->wl, d->y, d->N, d->H, d->W, d->Cin, d->Cout, d->num_slabs, d->ntaps, d->taps, d->in_stride,
                     d->HoP, d->WoP, d->Hout, d->Wout, d->osy, d->osx, d->oy0, d->ox0, d->accumulate, stream, "conv_nhwc", d->epi, d->w_format, d->x_format);
}

// Stride-2 transposed 3x3 convolution (padding 0): y[n][2i+ky][2j+kx][co] += x[n][i][j][ci] * w[co][ky*3+kx][ci], y = [N][2H+1][2W+1][Cout], as the four
// polyphase tap convolutions on the input grid in ONE launch (phases of 1, 2, 2 and 4 taps; conv2d_resample.py:113-126 runs this as conv_transpose2d).
extern "C" int gp3d_conv_transpose_s2_nhwc(const void* xh, const void* xl, const void* wh, const void* wl, int w_format, int x_format, float* y,
                                           int N, int H, int W, int Cin, int Cout, void* stream) {
    int taps[27], phases[24], nt = 0, np = 0;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            const int first = nt;
            // output row 2i + a: a == 0 -> (input offset 0, ky 0), (-1, ky 2); a == 1 -> (0, ky 1); same along x
            const int nky = a == 0 ? 2 : 1, nkx = b == 0 ? 2 : 1;
            for (int iy = 0; iy < nky; iy++)
                for (int ix = 0; ix < nkx; ix++) {
                    const int dy = (a == 0 && iy == 1) ? -1 : 0, ky = a == 0 ? (iy == 0 ? 0 : 2) : 1;
                    const int dx = (b == 0 && ix == 1) ? -1 : 0, kx = b == 0 ? (ix == 0 ? 0 : 2) : 1;
                    taps[3 * nt] = dy; taps[3 * nt + 1] = dx; taps[3 * nt + 2] = ky * 3 + kx; nt++;
                }
            phases[6 * np] = nt - first; phases[6 * np + 1] = first; phases[6 * np + 2] = a; phases[6 * np + 3] = b;
            phases[6 * np + 4] = H + 1 - a; phases[6 * np + 5] = W + 1 - b; np++;
        }
    return conv_impl(xh, xl, wh, wl, y, N, H, W, Cin, Cout, 9, nt, taps, 1, H + 1, W + 1, 2 * H + 1, 2 * W + 1, 2, 2, 0, 0, 0, stream,
                     "conv_transpose_s2_nhwc", nullptr, w_format, x_format, 4, phases);
}
