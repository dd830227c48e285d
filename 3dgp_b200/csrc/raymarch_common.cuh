// Device-side building blocks shared by the fused ray-march forward and backward kernels.
// Math follows SURVEY.md Appendix B / the reference (file:line cited at each step).
#pragma once
#include "common.cuh"

namespace rm {

constexpr int kC = 32;        // feature channels per plane (cfg.tri_plane.feat_dim)
constexpr int kH = 64;        // MLP hidden width (cfg.tri_plane.mlp.hid_dim)
constexpr int kThreads = 128;
constexpr int kWarps = 4;
constexpr int kMaxN = 64;

struct Params {
    const void* planes; int64_t psB, psP, psY, psX;   // channel stride is 1 (checked on the host)
    const float* ray_o; const float* ray_d;
    const float* w1; const float* b1; const float* w2; const float* b2;
    const float* u_coarse; const float* u_fine; const float* sn_coarse; const float* sn_fine;
    float* rgb; float* depth; float* wsum; float* tfinal;
    // backward only
    const float* g_rgb; const float* g_depth;
    float* g_planes; float* g_w1; float* g_b1; float* g_w2; float* g_b2; float* g_ray_o; float* g_ray_d;
    gp3d_raymarch_opts o;
    // in-kernel ray generation (third-generation forward, gp3d_generate_rays): pinhole camera per image instead of ray_o / ray_d
    const float* cam_c2w;      // [B][4][4] row-major cam2world (rendering_utils.py:194-218)
    const float* cam_fov;      // [B] degrees
    const float* patch_scale;  // [B][2] or NULL (full frame)
    const float* patch_offset; // [B][2] or NULL
    int img_w, img_h;          // rays form an img_h x img_w grid, row-major (ray = y * img_w + x); 0 = unknown (1-D tiles)
};

// One pinhole-camera ray (tri_plane_renderer.py:487-527): NDC grid x in linspace(-1, 1, w), y in linspace(1, -1, h) (ATen's symmetric linspace),
// optional patch transform (:511-512), z = -1 / tan(fov / 2), normalise, rotate by cam2world; origin = cam2world translation.
__device__ __forceinline__ float linspace_sym(float start, float end, int steps, int i) {
    const float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}
__device__ __forceinline__ void generate_ray(const Params& p, int b, int x, int y, float (&o)[3], float (&d)[3]) {
    const float* M = p.cam_c2w + (size_t)b * 16;
    float xs = linspace_sym(-1.f, 1.f, p.img_w, x), ys = linspace_sym(1.f, -1.f, p.img_h, y);
    if (p.patch_scale != nullptr) {
        xs = (xs + 1.0f) * p.patch_scale[b * 2 + 0] - 1.0f + p.patch_offset[b * 2 + 0] * 2.0f;
        ys = (ys + 1.0f) * p.patch_scale[b * 2 + 1] - 1.0f + p.patch_offset[b * 2 + 1] * 2.0f;
    }
    const float fov_rad = p.cam_fov[b] / 360.0f * 2.0f * 3.14159265358979323846f;
    const float z = -1.0f / tanf(fov_rad * 0.5f);
    const float nrm = sqrtf(xs * xs + ys * ys + z * z);
    const float dx = xs / nrm, dy = ys / nrm, dz = z / nrm;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        d[i] = M[i * 4 + 0] * dx + M[i * 4 + 1] * dy + M[i * 4 + 2] * dz;
        o[i] = M[i * 4 + 3];
    }
}

// ---------------------------------------------------------------------------------------------
// Counter-based RNG (Philox4x32-10) for the production mode (u_* == NULL).  Stream layout:
// counter = (block_of_4, pass, ray_lo, ray_hi) + offset, key = seed.  pass: 0 u_coarse 1 u_fine 2 sn_coarse 3 sn_fine.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ uint32_t rng_u32(const gp3d_raymarch_opts& o, uint64_t ray, int i, int pass) {
    uint64_t blk = (uint64_t)(i >> 2) + o.offset;
    uint4 ctr = make_uint4((uint32_t)blk, (uint32_t)pass, (uint32_t)ray, (uint32_t)(ray >> 32) ^ (uint32_t)(blk >> 32));
    uint4 r = philox4x32_10(ctr, make_uint2((uint32_t)o.seed, (uint32_t)(o.seed >> 32)));
    int k = i & 3;
    return k == 0 ? r.x : k == 1 ? r.y : k == 2 ? r.z : r.w;
}
__device__ __forceinline__ float rng_uniform(const gp3d_raymarch_opts& o, uint64_t ray, int i, int pass) {
    return (float)(rng_u32(o, ray, i, pass) >> 8) * (1.0f / 16777216.0f);   // [0,1), 24 bits like torch.rand
}
__device__ __forceinline__ float rng_normal(const gp3d_raymarch_opts& o, uint64_t ray, int i, int pass) {
    uint32_t a = rng_u32(o, ray, i, pass);
    uint32_t b = rng_u32(o, ray ^ 0x8000000000000000ull, i, pass);
    float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);             // (0,1]
    float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

// ---------------------------------------------------------------------------------------------
// torch.linspace(0, 1, N)[i] as ATen computes it (symmetric formula) -- tri_plane_renderer.py:224.
__device__ __forceinline__ float linspace01(int i, int N) {
    const float step = 1.0f / (float)(N - 1);
    return (i < N / 2) ? step * (float)i : 1.0f - step * (float)(N - 1 - i);
}
// Stratified coarse depth in s-space (tri_plane_renderer.py:226-230).
__device__ __forceinline__ float coarse_s(int i, int N, float u) {
    const float g = linspace01(i, N);
    const float lo = (i == 0) ? g : 0.5f * (g + linspace01(i - 1, N));
    const float hi = (i == N - 1) ? g : 0.5f * (linspace01(i + 1, N) + g);
    return lo + (hi - lo) * u;
}
// s -> t (tri_plane_renderer.py:132)
__device__ __forceinline__ float s_to_t(float s, float t0, float t1) { return s * t1 + (1.0f - s) * t0; }

// F.softplus(x) with beta=1, threshold=20 (tri_plane_renderer.py:360) / relu (:362)
__device__ __forceinline__ float density_act(float x, int clamp_mode) {
    if (clamp_mode == 1) return fmaxf(x, 0.f);
    return (x > 20.f) ? x : log1pf(expf(x));
}
__device__ __forceinline__ float density_act_grad(float x, int clamp_mode) {
    if (clamp_mode == 1) return x > 0.f ? 1.f : 0.f;
    return (x > 20.f) ? 1.f : 1.f / (1.f + expf(-x));
}

// ---------------------------------------------------------------------------------------------
// Bilinear footprint of one sample on one plane (F.grid_sample, bilinear, align_corners=True, zeros;
// tri_plane_renderer.py:584).  The footprint is re-expressed on an always-in-bounds 2x2 texel block
// (x0c, y0c)..(x0c+1, y0c+1) with per-axis weights (wa on the low texel, wb on the high texel); taps that fall
// outside the plane get weight 0, which is exactly the zero-padding rule.
struct Axis { int i0; float wa, wb; float da, db; };   // da,db: d(wa)/d(pix), d(wb)/d(pix)
__device__ __forceinline__ Axis axis_footprint(float coord, int P) {
    const float pix = (coord + 1.0f) * 0.5f * (float)(P - 1);
    const float fl = floorf(pix);
    const float t = pix - fl;                    // weight of the high tap
    const float omt = (fl + 1.0f) - pix;         // weight of the low tap (as ATen: ix_se - ix)
    Axis a;
    // Guard against huge / NaN coordinates before the float->int conversion.
    if (!(pix > -2.0f && pix < (float)P + 1.0f)) { a.i0 = 0; a.wa = a.wb = 0.f; a.da = a.db = 0.f; return a; }
    const int i = (int)fl;
    if (i >= 0 && i <= P - 2) { a.i0 = i; a.wa = omt; a.wb = t; a.da = -1.f; a.db = 1.f; }
    else if (i == -1)         { a.i0 = 0; a.wa = t; a.wb = 0.f; a.da = 1.f; a.db = 0.f; }
    else if (i == P - 1)      { a.i0 = P - 2; a.wa = 0.f; a.wb = omt; a.da = 0.f; a.db = -1.f; }
    else                      { a.i0 = 0; a.wa = a.wb = 0.f; a.da = a.db = 0.f; }
    return a;
}

struct Footprint {      // one sample, three planes
    int base[3];        // element offset of texel (y0c, x0c) of plane p relative to the image base
    float wxa[3], wxb[3], wya[3], wyb[3];
};

// coords = (o + t d) / box_half; plane0 <- (x,y), plane1 <- (x,z), plane2 <- (y,z) where the first
// component indexes width (ix) and the second height (iy)  (tri_plane_renderer.py:576-582).
__device__ __forceinline__ void sample_footprint(Footprint& fp, float px, float py, float pz, const Params& p) {
    const int P = p.o.P;
    const Axis ax = axis_footprint(px, P), ay = axis_footprint(py, P), az = axis_footprint(pz, P);
    const Axis* U[3] = {&ax, &ax, &ay};
    const Axis* V[3] = {&ay, &az, &az};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        fp.base[k] = (int)(k * p.psP + (int64_t)V[k]->i0 * p.psY + (int64_t)U[k]->i0 * p.psX);
        fp.wxa[k] = U[k]->wa; fp.wxb[k] = U[k]->wb; fp.wya[k] = V[k]->wa; fp.wyb[k] = V[k]->wb;
    }
}

// 16-byte (f32) / 8-byte (f16) load of 4 consecutive channels through the read-only path.
template <class T> __device__ __forceinline__ float4 ld_tex4(const T* p);
template <> __device__ __forceinline__ float4 ld_tex4<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 ld_tex4<__half>(const __half* p) {
    uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

}  // namespace rm
