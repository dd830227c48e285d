// Shared helpers for lib3dgp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/gp3d_b200.h"

#ifndef GP3D_NUM_SMS
#define GP3D_NUM_SMS 148   // B200: 2 dies x 74 SMs
#endif

void gp3d_set_error(const char* fmt, ...);

#define GP3D_CHECK_ARG(cond, ...)                          \
    do {                                                   \
        if (!(cond)) {                                     \
            gp3d_set_error(__VA_ARGS__);                   \
            return GP3D_E_BADARG;                          \
        }                                                  \
    } while (0)

#define GP3D_RETURN_LAUNCH()                                                  \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) {                                             \
            gp3d_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(e__)); \
            return (int)e__;                                                  \
        }                                                                     \
        return GP3D_OK;                                                       \
    } while (0)

template <class T> struct io_traits;
template <> struct io_traits<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct io_traits<__half> {
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};
template <> struct io_traits<__nv_bfloat16> {
    static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 16-byte vector of T (4 floats or 8 halves) with float unpack/pack.
template <class T> struct vec16;
template <> struct vec16<float> {
    static constexpr int N = 4;
    float4 v;
    __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
    __device__ __forceinline__ void unpack(float* f) const { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ __forceinline__ void pack(const float* f) { v = make_float4(f[0], f[1], f[2], f[3]); }
};
template <> struct vec16<__half> {
    static constexpr int N = 8;
    uint4 v;
    __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
    __device__ __forceinline__ void unpack(float* f) const {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    }
};
template <> struct vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    uint4 v;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void zero() { v = make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
    __device__ __forceinline__ void unpack(float* f) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    __device__ __forceinline__ void pack(const float* f) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; i++) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
};

static inline bool gp3d_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Grid sized as a multiple of the SM count (persistent-style grid-stride kernels).
static inline int gp3d_grid_for(int64_t work_items, int threads, int max_ctas_per_sm) {
    int64_t blocks = (work_items + threads - 1) / threads;
    int64_t cap = (int64_t)GP3D_NUM_SMS * max_ctas_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
