// Fused tri-plane ray-march, forward, second generation (sm_100a): kernel + launcher.  The per-CTA pipeline lives in raymarch2.cuh.
#include "raymarch2.cuh"

namespace rm2 {

template <class PT, int MODE>
__global__ void __launch_bounds__(kThreads, 3) raymarch_fwd2_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem s;
    s.carve(smem_raw, p.o.N);
    const int N = p.o.N, NP = s.NP, R = p.o.R, M2 = 2 * N;
    const int tid = threadIdx.x;
    const int blocks_per_img = (R + TR - 1) / TR;
    const int b = blockIdx.x / blocks_per_img;
    const int r0 = (blockIdx.x - b * blocks_per_img) * TR;
    const int nrays = min(TR, R - r0);
    const int64_t ray_base = (int64_t)b * R + r0;
    const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;
    const float t0 = p.o.ray_start, t1 = p.o.ray_end;

    stage_weights(s, p);
    for (int t = tid; t < nrays * 3; t += kThreads) { s.ro[t] = p.ray_o[ray_base * 3 + t]; s.rd[t] = p.ray_d[ray_base * 3 + t]; }
    __syncthreads();

    forward_phases2<PT, MODE, false>(s, p, img, ray_base, nrays, nullptr);
    auto depth_of = [&](int rl, int code) { return s_to_t(code < N ? s.s_co[rl * NP + code] : s.bufA[rl * NP + code - N], t0, t1); };
    auto value_of = [&](int rl, int code) { return code < N ? s.out_co[rl * (N + 1) + code] : s.out_fi[rl * (N + 1) + code - N]; };
    {                                                                       // D4: weighted sums, 16 lanes per ray
        const int rl = tid >> 4, l16 = tid & 15;
        float cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, ws = 0.f;
        if (rl < nrays) {
            for (int m = l16; m < M2; m += 16) {
                const int code = s.ord[rl * M2 + m];
                const float w = s.wm[rl * (M2 + 1) + m];
                const float4 v = value_of(rl, code);
                cr += w * v.x; cg += w * v.y; cb += w * v.z; dep += w * depth_of(rl, code); ws += w;
            }
        }
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {
            cr += __shfl_xor_sync(0xffffffffu, cr, off); cg += __shfl_xor_sync(0xffffffffu, cg, off); cb += __shfl_xor_sync(0xffffffffu, cb, off);
            dep += __shfl_xor_sync(0xffffffffu, dep, off); ws += __shfl_xor_sync(0xffffffffu, ws, off);
        }
        if (rl < nrays && l16 == 0) {
            const float wagg = ws;
            if (p.o.last_back) {
                const int code = s.ord[rl * M2 + M2 - 1];
                const float4 v = value_of(rl, code);
                const float extra = 1.f - wagg;
                cr += extra * v.x; cg += extra * v.y; cb += extra * v.z; dep += extra * depth_of(rl, code); ws += extra;
            }
            if (p.o.white_back_end_idx > 0) {
                const float add = 1.f - wagg;
                cr += add;
                if (p.o.white_back_end_idx > 1) cg += add;
                if (p.o.white_back_end_idx > 2) cb += add;
            }
            const int64_t ri = ray_base + rl;
            p.rgb[ri * 3 + 0] = cr; p.rgb[ri * 3 + 1] = cg; p.rgb[ri * 3 + 2] = cb;
            p.depth[ri] = dep; p.wsum[ri] = ws; p.tfinal[ri] = s.wm[rl * (M2 + 1) + M2];
        }
    }
}

template <class PT, int MODE>
int launch(const Params& p, cudaStream_t st) {
    const size_t smem = Smem::bytes(p.o.N);
    auto kern = raymarch_fwd2_kernel<PT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_forward(v2): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int blocks = p.o.B * ((p.o.R + TR - 1) / TR);
    kern<<<blocks, kThreads, smem, st>>>(p);
    return 0;
}

}  // namespace rm2

int gp3d_raymarch_forward_v2(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st) {
    if (planes_dtype == GP3D_F32) return mode == 2 ? rm2::launch<float, 2>(p, st) : rm2::launch<float, 1>(p, st);
    return mode == 2 ? rm2::launch<__half, 2>(p, st) : rm2::launch<__half, 1>(p, st);
}
