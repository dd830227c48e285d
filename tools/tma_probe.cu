// GPU probe: which tiled-TMA box shapes load correctly on this part?  One configuration per process (a device fault is sticky):
//   tma_probe <rank 3|4> <elem bytes 2|4> <box inner elems> <box rows> <tensor W> <tensor H> [c0 c1]   (box start coordinates, default -1 -1)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe_kernel(const __grid_constant__ CUtensorMap tm, unsigned char* out, int bytes, int c0, int c1) {
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(s32(tile)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(1) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(tile)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(1), "r"(0) : "memory");
    }
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; spin++)
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(done) : "r"(s32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = done ? tile[i] : 0xEE;
}

int main(int argc, char** argv) {
    if (argc < 7) { printf("usage\n"); return 2; }
    const int rank = atoi(argv[1]), es = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]), W = atoi(argv[5]), H = atoi(argv[6]);
    const int planes = 3;
    const int c0 = argc > 7 ? atoi(argv[7]) : -1, c1 = argc > 8 ? atoi(argv[8]) : -1;
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)fnp;
    const size_t n = (size_t)planes * H * W;
    std::vector<unsigned char> h(n * es);
    for (size_t i = 0; i < n; i++) { if (es == 2) ((uint16_t*)h.data())[i] = (uint16_t)(i * 7 + 1); else ((uint32_t*)h.data())[i] = (uint32_t)(i * 7 + 1); }
    unsigned char *dx, *dout; cudaMalloc(&dx, n * es); cudaMemcpy(dx, h.data(), n * es, cudaMemcpyHostToDevice);
    const int bytes = bw * bh * es; cudaMalloc(&dout, bytes);
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes, 1};
    cuuint64_t strides[3] = {(cuuint64_t)W * es, (cuuint64_t)H * W * es, (cuuint64_t)planes * H * W * es};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("rank %d es %d box %dx%d W %d H %d: ENCODE FAILED %d\n", rank, es, bw, bh, W, H, (int)r); return 0; }
    if (rank == 3) { cudaFuncSetAttribute(probe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); probe_kernel<3><<<1, 128, bytes>>>(tm, dout, bytes, c0, c1); }
    else { cudaFuncSetAttribute(probe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); probe_kernel<4><<<1, 128, bytes>>>(tm, dout, bytes, c0, c1); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank %d es %d box %dx%d W %d H %d start (%d, %d): FAULT %s\n", rank, es, bw, bh, W, H, c0, c1, cudaGetErrorString(e)); return 0; }
    std::vector<unsigned char> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) {
        const int gy = y + c1, gx = x + c0;
        uint32_t want = 0;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) want = (uint32_t)(((size_t)1 * H * W + (size_t)gy * W + gx) * 7 + 1);
        uint32_t got = es == 2 ? ((uint16_t*)o.data())[y * bw + x] : ((uint32_t*)o.data())[y * bw + x];
        if (es == 2) want &= 0xffff;
        bad += (got != want);
    }
    printf("rank %d es %d box %dx%d W %d H %d start (%d, %d): OK, %d mismatches\n", rank, es, bw, bh, W, H, c0, c1, bad);
    return 0;
}
