// Optimiser step of the training loop as ONE pass over a module's flat parameter storage (sm_100a):
//   g   = nan_to_num(g_sum / world, 0, +1e5, -1e5)                 training_loop.py:340-341
//   Adam(betas, eps), torch.optim.Adam arithmetic, op for op       training_loop.py:190-205, 346 (opt.step())
//   p_ema = p + beta * (p_ema - p)                                 training_loop.py:357-364 (G only)
// HBM-bound: 4 B x (g, p, v [, m] [, ema]) read + (p, v [, m] [, ema]) written per parameter, each exactly once.
// The flat buffers are carved in 1024-element blocks (one CTA each), every parameter tensor starting on a block boundary,
// so a CTA never straddles two tensors and per-tensor step counts / "no gradient this phase" flags are one table look-up.
#include "common.cuh"

namespace {

struct AdamArgs {
    float* p; const float* g; float* m; float* v; float* ema;
    float grad_scale, posinf, neginf, beta1, beta2, omb1, omb2, eps, step_size, bc2_sqrt, ema_beta;
    const int* blk_seg; const float* seg_desc;
};

__device__ __forceinline__ float sanitize(float x, float scale, float posinf, float neginf) {
    x *= scale;
    if (isnan(x)) return 0.f;
    if (isinf(x)) return x > 0.f ? posinf : neginf;
    return x;
}

template <bool HAS_M, bool HAS_EMA>
__global__ void __launch_bounds__(256) adam_ema_kernel(AdamArgs a) {
    const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    float step_size = a.step_size, bc2_sqrt = a.bc2_sqrt;
    bool active = true;
    if (a.blk_seg) {
        const int seg = a.blk_seg[blockIdx.x];
        if (seg < 0) return;                                   // alignment padding
        const float4 d = reinterpret_cast<const float4*>(a.seg_desc)[seg];
        step_size = d.x; bc2_sqrt = d.y; active = d.z != 0.f;
    }
    float4 p4 = *reinterpret_cast<const float4*>(a.p + i);
    float pp[4] = {p4.x, p4.y, p4.z, p4.w};
    if (active) {
        const float4 g4 = *reinterpret_cast<const float4*>(a.g + i);
        const float4 v4 = *reinterpret_cast<const float4*>(a.v + i);
        float gg[4] = {g4.x, g4.y, g4.z, g4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, mm[4];
        if (HAS_M) { const float4 m4 = *reinterpret_cast<const float4*>(a.m + i); mm[0] = m4.x; mm[1] = m4.y; mm[2] = m4.z; mm[3] = m4.w; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float g = sanitize(gg[k], a.grad_scale, a.posinf, a.neginf);
            // exp_avg.lerp_(grad, 1 - beta1): ATen's lerp uses start + w*(end-start) for w < 0.5, end - (end-start)*(1-w) otherwise
            float m;
            if (HAS_M) {
                const float w = a.omb1, diff = g - mm[k];
                m = (w < 0.5f) ? __fmaf_rn(w, diff, mm[k]) : __fmaf_rn(-diff, 1.f - w, g);
                mm[k] = m;
            } else m = g;                                       // beta1 == 0: exp_avg == grad exactly, no state needed
            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
            const float v = __fmaf_rn(a.omb2 * g, g, vv[k] * a.beta2);
            vv[k] = v;
            // denom = sqrt(v) / sqrt(bc2) + eps ; p += -step_size * m / denom
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), a.eps);
            pp[k] = __fmaf_rn(-step_size, __fdiv_rn(m, denom), pp[k]);
        }
        *reinterpret_cast<float4*>(a.p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<float4*>(a.v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (HAS_M) *reinterpret_cast<float4*>(a.m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    }
    if (HAS_EMA) {                                              // p.lerp(p_ema, beta)
        const float4 e4 = *reinterpret_cast<const float4*>(a.ema + i);
        float ee[4] = {e4.x, e4.y, e4.z, e4.w};
        const float w = a.ema_beta;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float diff = ee[k] - pp[k];
            ee[k] = (w < 0.5f) ? __fmaf_rn(w, diff, pp[k]) : __fmaf_rn(-diff, 1.f - w, ee[k]);
        }
        *reinterpret_cast<float4*>(a.ema + i) = make_float4(ee[0], ee[1], ee[2], ee[3]);
    }
}

}  // namespace

extern "C" int gp3d_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t numel,
                                  float grad_scale, float posinf, float neginf, float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                                  float step_size, float bc2_sqrt, float ema_beta,
                                  const int* blk_seg, const float* seg_desc, void* stream) {
    GP3D_CHECK_ARG(p && g && v && numel >= 0, "adam_ema_step: null pointer");
    GP3D_CHECK_ARG(numel % 1024 == 0, "adam_ema_step: flat storage must be carved in 1024-element blocks (numel = %lld)", (long long)numel);
    GP3D_CHECK_ARG((beta1 == 0.f) == (m == nullptr), "adam_ema_step: exp_avg storage is required iff beta1 != 0");
    GP3D_CHECK_ARG((blk_seg == nullptr) == (seg_desc == nullptr), "adam_ema_step: blk_seg and seg_desc go together");
    GP3D_CHECK_ARG(numel / 1024 < 2147483647LL, "adam_ema_step: too many blocks");
    if (numel == 0) return GP3D_OK;
    AdamArgs a{p, g, m, v, ema, grad_scale, posinf, neginf, beta1, beta2, one_minus_beta1, one_minus_beta2, eps, step_size, bc2_sqrt, ema_beta, blk_seg, seg_desc};
    const int grid = (int)(numel / 1024);
    cudaStream_t st = (cudaStream_t)stream;
    if (m) { if (ema) adam_ema_kernel<true, true><<<grid, 256, 0, st>>>(a); else adam_ema_kernel<true, false><<<grid, 256, 0, st>>>(a); }
    else   { if (ema) adam_ema_kernel<false, true><<<grid, 256, 0, st>>>(a); else adam_ema_kernel<false, false><<<grid, 256, 0, st>>>(a); }
    GP3D_RETURN_LAUNCH();
}
