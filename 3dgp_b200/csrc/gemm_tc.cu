// Dense contraction engine on tcgen05 / TMEM (sm_100a):  D[M][N] (+)= A[M][K] * B[N][K]^T, bf16 operands, fp32 accumulate.
//
// One CTA per 128 x 128 output tile, 256 threads, warp-specialised:
//   warp 0 : TMA producer  -- one elected lane issues cp.async.bulk.tensor loads of the A and B k-blocks
//            ([128 rows][64 bf16], 128-byte swizzle) into a 4-stage shared-memory ring, signalling `full` mbarriers
//   warp 1 : MMA issuer    -- one elected lane issues 4 x tcgen05.mma (M128 N128 K16) per k-block into a 128-column TMEM
//            accumulator, then tcgen05.commit -> `empty` mbarrier of the stage (and the accumulator-ready barrier at the end)
//   warp 2 : TMEM allocator / deallocator
//   warps 4-7 : epilogue   -- tcgen05.ld 32 lanes x 32 columns per warp, fp32 rows written straight to D
// The implicit-GEMM convolution (conv_tc.cu) reuses this pipeline with a 4-D activation tensor map for the A operand.
#include "tc_common.cuh"

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

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4;
constexpr int kGemmThreads = 256;
constexpr uint32_t kStageBytes = (BM + BN) * BK * 2;     // 32 KB
constexpr uint32_t kTmemCols = 128;

struct GemmSmem {
    static constexpr size_t bytes() { return 1024 /*align slack*/ + (size_t)STAGES * kStageBytes + 256; }
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    float* __restrict__ D, int M, int N, int K, int accumulate) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tiles = smem;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * kStageBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = N / BN;
    const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x - tile_m * tiles_n;
    const int num_kb = K / BK;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; kb++) {
                const int st = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[st], ph ^ 1);
                unsigned char* sa = tiles + (size_t)st * kStageBytes;
                unsigned char* sb = sa + BM * BK * 2;
                mbar_expect_tx(&full_bar[st], kStageBytes);
                tma_load_2d(sa, &tmA, &full_bar[st], kb * BK, tile_m * BM);
                tma_load_2d(sb, &tmB, &full_bar[st], kb * BK, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16_f32(BM, BN);
            for (int kb = 0; kb < num_kb; kb++) {
                const int st = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full_bar[st], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(tiles + (size_t)st * kStageBytes);
                const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
                for (int k = 0; k < BK / 16; k++) {
                    const uint64_t da = make_desc_k_sw128(sa + k * 32);
                    const uint64_t db = make_desc_k_sw128(sb + k * 32);
                    umma_bf16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[st]);          // frees the smem stage when these MMAs have read it
            }
            umma_commit(accum_bar);                   // accumulator complete
        }
    } else if (warp >= 4) {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int row = tile_m * BM + q * 32 + lane;
        float* drow = D + (size_t)row * N + (size_t)tile_n * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                float4* dst = reinterpret_cast<float4*>(drow + c + j);
                if (accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                *dst = v;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc

static int encode_2d_bf16(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    gp3d_encode_tiled_fn enc = gp3d_get_encode_tiled();
    if (!enc) { gp3d_set_error("cuTensorMapEncodeTiled is not available from this driver"); return GP3D_E_UNSUPPORTED; }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { gp3d_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return GP3D_E_BADARG; }
    return 0;
}

extern "C" int gp3d_gemm_bf16_tn(const void* A, const void* B, float* D, int M, int N, int K, int accumulate, void* stream) {
    GP3D_CHECK_ARG(A && B && D, "gemm_bf16_tn: null pointer");
    GP3D_CHECK_ARG(M > 0 && N > 0 && K > 0 && M % tc::BM == 0 && N % tc::BN == 0 && K % tc::BK == 0,
                   "gemm_bf16_tn: need M %% 128 == 0, N %% 128 == 0, K %% 64 == 0 (got %d %d %d)", M, N, K);
    GP3D_CHECK_ARG(gp3d_aligned16(A) && gp3d_aligned16(B) && gp3d_aligned16(D), "gemm_bf16_tn: pointers must be 16-byte aligned");
    CUtensorMap tmA, tmB;
    int rc = encode_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, tc::BM, tc::BK);
    if (rc) return rc;
    rc = encode_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, tc::BN, tc::BK);
    if (rc) return rc;
    const size_t smem = tc::GemmSmem::bytes();
    cudaError_t e = cudaFuncSetAttribute(tc::gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("gemm_bf16_tn: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int grid = (M / tc::BM) * (N / tc::BN);
    tc::gemm_bf16_tn_kernel<<<grid, tc::kGemmThreads, smem, (cudaStream_t)stream>>>(tmA, tmB, D, M, N, K, accumulate);
    GP3D_RETURN_LAUNCH();
}
