// Fused tri-plane ray-march, forward, third generation (sm_100a): persistent kernel + launcher + ray generator.  Pipeline in raymarch3.cuh.
#include "raymarch3.cuh"

namespace rm3 {

template <class PT, int MODE, bool DENSE>
__global__ void __launch_bounds__(kT3, 2) raymarch_fwd3_kernel(Params p, int ntiles, int tiles_x, int tiles_per_img) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem3 s;
    s.carve(smem_raw, p.o.N);
    const int N = p.o.N, NP = s.NP, R = p.o.R, M2 = 2 * N;
    const int tid = threadIdx.x;
    const float t0 = p.o.ray_start, t1 = p.o.ray_end;

    stage_weights3(s, p);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img;
        const int tl = tile - b * tiles_per_img;
        const int64_t img_ray_base = (int64_t)b * R;
        const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;
        if (tid < TRAYS) {
            int rid, px = 0, py = 0;
            if (p.img_w > 0) {                      // 4 x 4 pixel tile
                const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
                py = ty * 4 + (tid >> 2); px = tx * 4 + (tid & 3);
                rid = (py < p.img_h && px < p.img_w) ? py * p.img_w + px : -1;
            } else {
                const int r = tl * TRAYS + tid;
                rid = r < R ? r : -1;
            }
            s.rid[tid] = rid;
            float o[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 1.f};
            if (rid >= 0) {
                if (p.cam_c2w != nullptr) generate_ray(p, b, px, py, o, d);
                else {
#pragma unroll
                    for (int k = 0; k < 3; k++) { o[k] = p.ray_o[(img_ray_base + rid) * 3 + k]; d[k] = p.ray_d[(img_ray_base + rid) * 3 + k]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) { s.ro[tid * 3 + k] = o[k]; s.rd[tid * 3 + k] = d[k]; }
        }
        __syncthreads();     // also orders stage_weights3 before the first use, and the previous tile's D4 reads before this tile's writes

        forward_phases3<PT, MODE, DENSE>(s, p, img, img_ray_base);

        {                                                                       // D4: weighted sums, 16 lanes per ray
            const int rl = tid >> 4, l16 = tid & 15;
            const int rid = s.rid[rl];
            auto depth_of = [&](int code) { return s_to_t(code < N ? s.s_co[rl * NP + code] : s.bufA[rl * NP + code - N], t0, t1); };
            auto value_of = [&](int code) { return code < N ? s.out_co[rl * (N + 1) + code] : s.out_fi[rl * (N + 1) + code - N]; };
            float cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f, ws = 0.f;
            if (rid >= 0) {
                for (int m = l16; m < M2; m += 16) {
                    const int code = s.ord[rl * M2 + m];
                    const float w = s.wm[rl * (M2 + 1) + m];
                    const float4 v = value_of(code);
                    cr += w * v.x; cg += w * v.y; cb += w * v.z; dep += w * depth_of(code); ws += w;
                }
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) {
                cr += __shfl_xor_sync(0xffffffffu, cr, off); cg += __shfl_xor_sync(0xffffffffu, cg, off); cb += __shfl_xor_sync(0xffffffffu, cb, off);
                dep += __shfl_xor_sync(0xffffffffu, dep, off); ws += __shfl_xor_sync(0xffffffffu, ws, off);
            }
            if (rid >= 0 && l16 == 0) {
                const float wagg = ws;
                if (p.o.last_back) {
                    const int code = s.ord[rl * M2 + M2 - 1];
                    const float4 v = value_of(code);
                    const float extra = 1.f - wagg;
                    cr += extra * v.x; cg += extra * v.y; cb += extra * v.z; dep += extra * depth_of(code); ws += extra;
                }
                if (p.o.white_back_end_idx > 0) {
                    const float add = 1.f - wagg;
                    cr += add;
                    if (p.o.white_back_end_idx > 1) cg += add;
                    if (p.o.white_back_end_idx > 2) cb += add;
                }
                const int64_t ri = img_ray_base + rid;
                p.rgb[ri * 3 + 0] = cr; p.rgb[ri * 3 + 1] = cg; p.rgb[ri * 3 + 2] = cb;
                p.depth[ri] = dep; p.wsum[ri] = ws; p.tfinal[ri] = s.wm[rl * (M2 + 1) + M2];
            }
        }
        __syncthreads();
    }
}

__global__ void generate_rays_kernel(Params p, float* ray_o, float* ray_d) {
    const int64_t n = (int64_t)p.o.B * p.o.R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / p.o.R), r = (int)(i - (int64_t)b * p.o.R);
        float o[3], d[3];
        generate_ray(p, b, r % p.img_w, r / p.img_w, o, d);
#pragma unroll
        for (int k = 0; k < 3; k++) { ray_o[i * 3 + k] = o[k]; ray_d[i * 3 + k] = d[k]; }
    }
}

template <class PT, int MODE, bool DENSE>
int launch3(const Params& p, cudaStream_t st) {
    const size_t smem = Smem3::bytes(p.o.N);
    auto kern = raymarch_fwd3_kernel<PT, MODE, DENSE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_forward(v3): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    int tiles_x = 0, tiles_per_img;
    if (p.img_w > 0) { tiles_x = (p.img_w + 3) / 4; tiles_per_img = tiles_x * ((p.img_h + 3) / 4); }
    else tiles_per_img = (p.o.R + TRAYS - 1) / TRAYS;
    const int ntiles = p.o.B * tiles_per_img;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kT3, smem);
    if (per_sm < 1) per_sm = 1;
    int sms = GP3D_NUM_SMS, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, kT3, smem, st>>>(p, ntiles, tiles_x, tiles_per_img);
    return 0;
}

}  // namespace rm3

// True when the third-generation kernel covers this call: 8-channel (32-byte fp32 / 16-byte fp16) vector loads need every stride to be a multiple of 8.
bool gp3d_raymarch_v3_ok(const rm::Params& p, int planes_dtype) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p.planes);
    return (p.psX % 8 == 0) && (p.psY % 8 == 0) && (p.psP % 8 == 0) && (p.psB % 8 == 0) && (a % (planes_dtype == GP3D_F32 ? 32 : 16) == 0) && p.o.N <= 64;
}

int gp3d_raymarch_forward_v3(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st) {
    const bool dense = (p.psX == 3 * rm::kC);
    if (planes_dtype == GP3D_F32) {
        if (dense) return mode == 2 ? rm3::launch3<float, 2, true>(p, st) : rm3::launch3<float, 1, true>(p, st);
        return mode == 2 ? rm3::launch3<float, 2, false>(p, st) : rm3::launch3<float, 1, false>(p, st);
    }
    if (dense) return mode == 2 ? rm3::launch3<__half, 2, true>(p, st) : rm3::launch3<__half, 1, true>(p, st);
    return mode == 2 ? rm3::launch3<__half, 2, false>(p, st) : rm3::launch3<__half, 1, false>(p, st);
}

extern "C" int gp3d_generate_rays(const float* c2w, const float* fov, const float* patch_scales, const float* patch_offsets,
                                  int B, int img_h, int img_w, float* ray_o, float* ray_d, void* stream) {
    GP3D_CHECK_ARG(c2w && fov && ray_o && ray_d, "generate_rays: null pointer");
    GP3D_CHECK_ARG((patch_scales == nullptr) == (patch_offsets == nullptr), "generate_rays: patch scales and offsets go together");
    GP3D_CHECK_ARG(B >= 1 && img_h >= 2 && img_w >= 2, "generate_rays: need B >= 1 and an image of at least 2 x 2 rays");
    rm::Params p{};
    p.cam_c2w = c2w; p.cam_fov = fov; p.patch_scale = patch_scales; p.patch_offset = patch_offsets; p.img_w = img_w; p.img_h = img_h;
    p.o.B = B; p.o.R = img_h * img_w;
    const int64_t n = (int64_t)B * img_h * img_w;
    const int grid = gp3d_grid_for(n, 256, 8);
    rm3::generate_rays_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, ray_o, ray_d);
    GP3D_RETURN_LAUNCH();
}
