// Entry points declared in include/gp3d_b200.h whose kernels are not built yet fail loudly here.
#include "common.cuh"

extern "C" int gp3d_gemm_bf16_tn(const void*, const void*, float*, int, int, int, int, void*) {
    gp3d_set_error("gemm_bf16_tn: tcgen05 GEMM is not built in this revision");
    return GP3D_E_UNSUPPORTED;
}
extern "C" int gp3d_conv2d_nhwc_bf16(const void*, const void*, float*, int, int, int, int, int, int, int, void*) {
    gp3d_set_error("conv2d_nhwc_bf16: tcgen05 implicit-GEMM conv is not built in this revision");
    return GP3D_E_UNSUPPORTED;
}
