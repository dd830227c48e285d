// 4x4 FIR resampling of NCHW (W-minor) planes with the input window staged by TMA (sm_100a): upsample2d (up 2, pad 2) and downsample2d (down 2, pad 1)
// of float16 / float32 tensors -- BASELINE configs[3]'s "super-res stack" ops (upfirdn2d.py:330-389: upsample2d / downsample2d with f = [1,3,3,1]).
//
// Why another kernel: the register-tiled LDG version (upfirdn2d.cu: upfirdn2d_wminor4_kernel) spends ~53 instructions per output on staging and
// index arithmetic and is ISSUE-bound at 0.13-0.22 of HBM (profiles/r1_hot_kernels_details.txt: issue slots 82 % busy, DRAM 14 %).  Here
//   * ONE 3-D TMA box {IW, IH, 1 plane} lands the haloed window in shared memory in the tensor's own element type; rows / columns outside the image
//     are zero-filled by the TMA unit (= the padding rule): no staging instructions, no bounds logic;
//   * the innermost START coordinate of a tiled TMA load must be 16-byte aligned (c0 * elemsize % 16 == 0; anything else raises an illegal-instruction
//     fault on sm_100a -- measured, tools/tma_probe.cu / profiles/r2_tma_box_start_probe.txt; negative aligned starts and boxes larger than the tensor
//     are fine), so the window starts A = 16 / elemsize columns left of the first output's column instead of 1: a thread reads three (up) / four (down)
//     aligned 4-element vectors per window row and uses elements 3 .. 8 (3 .. 12) of them, converted to float in registers;
//   * a thread produces 8 x 2 (up) or 4 x 2 (down) outputs from a 3 x 6 / 6 x 10 register window with compile-time tap indices
//     (7 / 26 instructions per output), stored as one 16-byte / 8-byte vector per row.
// Taps accumulate in the order of the index contract in upfirdn2d.cu (rows ascending, then columns; gain applied last): results are bit-identical to
// the generic kernels.
#include "tc_common.cuh"

namespace {
using namespace tc;

struct FirNchwParams {
    void* y; const float* f; int flip; float gain;
    int planes, outH, outW;
};

template <class T> struct rowvec;
template <> struct rowvec<__half> {
    // n consecutive halves starting at an address aligned to n * 2 bytes
    static __device__ __forceinline__ void ld4(const __half* p, float* v) {
        const uint2 r = *reinterpret_cast<const uint2*>(p);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    static __device__ __forceinline__ void ld2(const __half* p, float* v) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(p));
        v[0] = a.x; v[1] = a.y;
    }
    static __device__ __forceinline__ void ld8(const __half* p, float* v) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) { const float2 a = __half22float2(h[i]); v[2 * i] = a.x; v[2 * i + 1] = a.y; }
    }
    static __device__ __forceinline__ void st8(__half* p, const float* v) {
        uint4 r; __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = r;
    }
    static __device__ __forceinline__ void st4(__half* p, const float* v) {
        uint2 r; __half2* h = reinterpret_cast<__half2*>(&r);
        h[0] = __floats2half2_rn(v[0], v[1]); h[1] = __floats2half2_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = r;
    }
};
template <> struct rowvec<float> {
    static __device__ __forceinline__ void ld4(const float* p, float* v) { const float4 r = *reinterpret_cast<const float4*>(p); v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
    static __device__ __forceinline__ void ld2(const float* p, float* v) { const float2 r = *reinterpret_cast<const float2*>(p); v[0] = r.x; v[1] = r.y; }
    static __device__ __forceinline__ void ld8(const float* p, float* v) { ld4(p, v); ld4(p + 4, v + 4); }
    static __device__ __forceinline__ void st8(float* p, const float* v) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    static __device__ __forceinline__ void st4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};

// sf[ky * 4 + kx] = coefficient of tap (ky, kx) (upfirdn2d.cu::stage_filter)
__device__ __forceinline__ void stage_filter16(float* sf, const float* f, int flip) {
    if (threadIdx.x < 16) {
        const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
        sf[threadIdx.x] = f[(flip ? ky : 3 - ky) * 4 + (flip ? kx : 3 - kx)];
    }
}

// ---- up = 2, pad0 = 2: output (2j + px, 2i + py) = sum_{a, b in {0,1}} in[i - 1 + py + a][j - 1 + px + b] * sf[2a + py][2b + px].
// Tile 128 x 32 outputs of one plane; window origin (ox0 / 2 - A, oy0 / 2 - 1); thread (xg, yg): outputs x = 8 xg .. 8 xg + 7, rows 2 yg, 2 yg + 1.
constexpr int U_TOW = 128, U_TOH = 32, U_IH = 18;
template <class T> struct win {
    static constexpr int A = 16 / (int)sizeof(T);          // alignment quantum of the innermost TMA start coordinate, in elements
    static constexpr int BO = A - 4;                       // a thread's first 4-element vector starts BO columns right of its group origin
    static constexpr int U_IW = (4 * 15 + BO + 12 + A - 1) / A * A;      // 72 (float) / 80 (half)
    static constexpr int D_IW = (8 * 15 + BO + 16 + A - 1) / A * A;      // 136 (float) / 144 (half)
};

template <class T>
__global__ void __launch_bounds__(256) fir4_nchw_up2_kernel(const __grid_constant__ CUtensorMap tmX, FirNchwParams p) {
    constexpr int U_IW = win<T>::U_IW;
    __shared__ __align__(128) T tile[U_IH * U_IW];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float sf[16];
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * U_TOW, oy0 = blockIdx.y * U_TOH, plane = blockIdx.z;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    stage_filter16(sf, p.f, p.flip);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(U_IH * U_IW * sizeof(T)));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(&bar)), "r"(ox0 / 2 - win<T>::A), "r"(oy0 / 2 - 1), "r"(plane) : "memory");
    }
    const int xg = tid & 15, yg = tid >> 4;
    float cf[16];
#pragma unroll
    for (int k = 0; k < 16; k++) cf[k] = sf[k];
    mbar_wait(&bar, 0);
    float v[3][6];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const T* row = tile + (yg + r) * U_IW + 4 * xg + win<T>::BO;
        float t12[12];
        rowvec<T>::ld4(row, t12); rowvec<T>::ld4(row + 4, t12 + 4); rowvec<T>::ld4(row + 8, t12 + 8);
#pragma unroll
        for (int k = 0; k < 6; k++) v[r][k] = t12[3 + k];
    }
    T* yb = reinterpret_cast<T*>(p.y) + (size_t)plane * p.outH * p.outW;
#pragma unroll
    for (int py = 0; py < 2; py++) {
        const int oy = oy0 + 2 * yg + py;
        float o[8];
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int px = 0; px < 2; px++) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int b = 0; b < 2; b++) acc = fmaf(v[py + a][jj + px + b], cf[(2 * a + py) * 4 + 2 * b + px], acc);
                o[2 * jj + px] = acc * p.gain;
            }
        if (oy < p.outH) {
            const int ox = ox0 + 8 * xg;
            T* dst = yb + (size_t)oy * p.outW + ox;
            if (ox + 8 <= p.outW && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) rowvec<T>::st8(dst, o);
            else {
#pragma unroll
                for (int j = 0; j < 8; j++) if (ox + j < p.outW) io_traits<T>::st(dst + j, o[j]);
            }
        }
    }
}

// ---- down = 2, pad0 = 1: output (x, y) = sum_{ky, kx} in[2y - 1 + ky][2x - 1 + kx] * sf[ky][kx].
// Tile 64 x 32 outputs; window origin (2 ox0 - A, 2 oy0 - 1); thread (xg, yg): outputs x = 4 xg .. 4 xg + 3, rows 2 yg, 2 yg + 1.
constexpr int D_TOW = 64, D_TOH = 32, D_IH = 66;

template <class T>
__global__ void __launch_bounds__(256) fir4_nchw_down2_kernel(const __grid_constant__ CUtensorMap tmX, FirNchwParams p) {
    constexpr int D_IW = win<T>::D_IW;
    extern __shared__ __align__(128) unsigned char dsm[];
    T* tile = reinterpret_cast<T*>(dsm);
    __shared__ __align__(8) uint64_t bar;
    __shared__ float sf[16];
    const int tid = threadIdx.x;
    const int ox0 = blockIdx.x * D_TOW, oy0 = blockIdx.y * D_TOH, plane = blockIdx.z;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    stage_filter16(sf, p.f, p.flip);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(D_IH * D_IW * sizeof(T)));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(smem_u32(&bar)), "r"(2 * ox0 - win<T>::A), "r"(2 * oy0 - 1), "r"(plane) : "memory");
    }
    const int xg = tid & 15, yg = tid >> 4;
    float cf[16];
#pragma unroll
    for (int k = 0; k < 16; k++) cf[k] = sf[k];
    mbar_wait(&bar, 0);
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
#pragma unroll
    for (int r = 0; r < 6; r++) {                          // window rows 4 yg + r feed output row i with tap ky = r - 2 i
        float t16[16], v[10];
        const T* row = tile + (4 * yg + r) * D_IW + 8 * xg + win<T>::BO;
        rowvec<T>::ld4(row, t16); rowvec<T>::ld4(row + 4, t16 + 4); rowvec<T>::ld4(row + 8, t16 + 8); rowvec<T>::ld4(row + 12, t16 + 12);
#pragma unroll
        for (int k = 0; k < 10; k++) v[k] = t16[3 + k];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int ky = r - 2 * i;
            if (ky < 0 || ky >= 4) continue;
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int kx = 0; kx < 4; kx++) acc[i][j] = fmaf(v[2 * j + kx], cf[ky * 4 + kx], acc[i][j]);
        }
    }
    T* yb = reinterpret_cast<T*>(p.y) + (size_t)plane * p.outH * p.outW;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int oy = oy0 + 2 * yg + i, ox = ox0 + 4 * xg;
        if (oy >= p.outH) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; j++) o[j] = acc[i][j] * p.gain;
        T* dst = yb + (size_t)oy * p.outW + ox;
        if (ox + 4 <= p.outW && ((reinterpret_cast<uintptr_t>(dst) & (sizeof(T) * 4 - 1)) == 0)) rowvec<T>::st4(dst, o);
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (ox + j < p.outW) io_traits<T>::st(dst + j, o[j]);
        }
    }
}

template <class T>
int encode_plane_map(CUtensorMap* tm, const void* x, int planes, int H, int W, int IW, int IH) {
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) return GP3D_E_UNSUPPORTED;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(T), (cuuint64_t)H * W * sizeof(T)};
    cuuint32_t box[3] = {(cuuint32_t)IW, (cuuint32_t)IH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : GP3D_E_UNSUPPORTED;
}

}  // namespace

// Called by upfirdn2d.cu's dispatcher for dense NCHW planes.  Returns 0 when the launch was issued, GP3D_E_UNSUPPORTED when the shape is not covered
// (the caller falls through to the register-tiled kernels).  mode: 0 = up 2 (pad0 = 2), 1 = down 2 (pad0 = 1).
int gp3d_fir4_nchw_launch(const void* x, const float* f, void* y, int is_half, int mode, int planes, int H, int W, int outH, int outW, int flip, float gain,
                          cudaStream_t st) {
    const size_t es = is_half ? 2 : 4;
    if (((size_t)W * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(x) & 15u) != 0 || planes > 2147483647 / 2) return GP3D_E_UNSUPPORTED;
    FirNchwParams p{y, f, flip, gain, planes, outH, outW};
    CUtensorMap tm;
    if (mode == 0) {
        const dim3 grid((outW + U_TOW - 1) / U_TOW, (outH + U_TOH - 1) / U_TOH, planes);
        if (grid.y > 65535 || planes > 65535) return GP3D_E_UNSUPPORTED;
        if (is_half) { if (encode_plane_map<__half>(&tm, x, planes, H, W, win<__half>::U_IW, U_IH)) return GP3D_E_UNSUPPORTED; fir4_nchw_up2_kernel<__half><<<grid, 256, 0, st>>>(tm, p); }
        else { if (encode_plane_map<float>(&tm, x, planes, H, W, win<float>::U_IW, U_IH)) return GP3D_E_UNSUPPORTED; fir4_nchw_up2_kernel<float><<<grid, 256, 0, st>>>(tm, p); }
        return 0;
    }
    const dim3 grid((outW + D_TOW - 1) / D_TOW, (outH + D_TOH - 1) / D_TOH, planes);
    if (grid.y > 65535 || planes > 65535) return GP3D_E_UNSUPPORTED;
    if (is_half) {
        if (encode_plane_map<__half>(&tm, x, planes, H, W, win<__half>::D_IW, D_IH)) return GP3D_E_UNSUPPORTED;
        fir4_nchw_down2_kernel<__half><<<grid, 256, (size_t)D_IH * win<__half>::D_IW * 2, st>>>(tm, p);
    } else {
        if (encode_plane_map<float>(&tm, x, planes, H, W, win<float>::D_IW, D_IH)) return GP3D_E_UNSUPPORTED;
        fir4_nchw_down2_kernel<float><<<grid, 256, (size_t)D_IH * win<float>::D_IW * 4, st>>>(tm, p);
    }
    return 0;
}
