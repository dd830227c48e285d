// Fused tri-plane ray-march, forward (sm_100a).  See raymarch_block.cuh for the per-CTA pipeline (passes A-C);
// this file adds pass D (depth merge + final compositing) and the C-ABI entry point.
// Algorithmic HBM traffic: every touched plane texel once + 24 B/ray in + 2*N*4 B/ray of injected variates
// (parity mode only) + 24 B/ray out.  No per-sample tensor ever reaches HBM.
#include "raymarch_block.cuh"
#include <stdlib.h>

namespace rm {

template <class PT, int TR>
__global__ void __launch_bounds__(kThreads) raymarch_fwd_kernel(Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Block<TR> s;
    s.carve(smem_raw, p.o.N);
    const int N = p.o.N, R = p.o.R;
    const int tid = threadIdx.x;
    const int blocks_per_img = (R + TR - 1) / TR;
    const int b = blockIdx.x / blocks_per_img;
    const int r0 = (blockIdx.x - b * blocks_per_img) * TR;
    const int nrays = min(TR, R - r0);
    const int64_t ray_base = (int64_t)b * R + r0;
    const PT* img = reinterpret_cast<const PT*>(p.planes) + (int64_t)b * p.psB;

    stage_mlp<TR>(s, p);
    for (int t = tid; t < nrays * 3; t += kThreads) { s.ro[t] = p.ray_o[ray_base * 3 + t]; s.rd[t] = p.ray_d[ray_base * 3 + t]; }
    __syncthreads();

    forward_passes<PT, TR>(s, p, img, ray_base, nrays);

    // ---- D: merge + final compositing in t-space (tri_plane_renderer.py:163-166, 196-206, 353-405)
    if (tid < nrays) {
        const int rl = tid;
        Merge<TR> mg(s, p, rl);
        float T = 1.f, wsum = 0.f, cr = 0.f, cg = 0.f, cb = 0.f, dep = 0.f;
        float tcur, tnext = 0.f; float4 cur, nxt;
        mg.pop(tcur, cur);
        nxt = cur;
        for (int m = 0; m < 2 * N; m++) {
            const bool last = (m == 2 * N - 1);
            if (!last) mg.pop(tnext, nxt);
            const float delta = last ? (p.o.use_inf_depth ? 1e10f : 1e-3f) : (tnext - tcur);
            const float sig = density_act(cur.w, p.o.clamp_mode);
            const float alpha = 1.f - expf(-delta * sig);
            const float w = alpha * T;
            T *= (1.f - alpha + 1e-10f);
            wsum += w;
            cr += w * cur.x; cg += w * cur.y; cb += w * cur.z; dep += w * tcur;
            if (!last) { cur = nxt; tcur = tnext; }
        }
        const float wagg = wsum;
        if (p.o.last_back) {                       // weights[:, :, -1] += 1 - weights_agg  (:386-387)
            const float extra = 1.f - wagg;
            cr += extra * cur.x; cg += extra * cur.y; cb += extra * cur.z; dep += extra * tcur;
            wsum += extra;
        }
        if (p.o.white_back_end_idx > 0) {          // :392-395
            const float add = 1.f - wagg;
            cr += add;
            if (p.o.white_back_end_idx > 1) cg += add;
            if (p.o.white_back_end_idx > 2) cb += add;
        }
        const int64_t ro_i = ray_base + rl;
        p.rgb[ro_i * 3 + 0] = cr; p.rgb[ro_i * 3 + 1] = cg; p.rgb[ro_i * 3 + 2] = cb;
        p.depth[ro_i] = dep; p.wsum[ro_i] = wsum; p.tfinal[ro_i] = T;
    }
}

constexpr int kTR = 16;

template <class PT>
int launch_fwd(const Params& p, cudaStream_t s) {
    const size_t smem = Block<kTR>::bytes(p.o.N);
    auto kern = raymarch_fwd_kernel<PT, kTR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { gp3d_set_error("raymarch_forward: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    const int blocks = p.o.B * ((p.o.R + kTR - 1) / kTR);
    kern<<<blocks, kThreads, smem, s>>>(p);
    return 0;
}

}  // namespace rm

int gp3d_raymarch_forward_v2(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st);   // raymarch_fwd2.cu

int gp3d_raymarch_check(const void* planes, int planes_dtype, int64_t psB, int64_t psP, int64_t psC, int64_t psY,
                        int64_t psX, const gp3d_raymarch_opts* o, const char* who) {
    GP3D_CHECK_ARG(o != nullptr, "%s: opts is null", who);
    GP3D_CHECK_ARG(planes != nullptr, "%s: planes is null", who);
    GP3D_CHECK_ARG(o->B >= 1 && o->R >= 1, "%s: empty batch / ray set", who);
    GP3D_CHECK_ARG(o->N >= 3 && o->N <= rm::kMaxN, "%s: samples per pass N=%d outside [3, %d]", who, o->N, rm::kMaxN);
    GP3D_CHECK_ARG(o->P >= 2, "%s: plane resolution must be >= 2", who);
    GP3D_CHECK_ARG(planes_dtype == GP3D_F32 || planes_dtype == GP3D_F16, "%s: planes must be float32 or float16", who);
    GP3D_CHECK_ARG(o->box_half > 0.f, "%s: box_half must be positive", who);
    if (o->C != rm::kC || o->H != rm::kH) {
        gp3d_set_error("%s: only feat_dim=32 / hid_dim=64 / n_layers=2 tri-plane MLPs are built (got C=%d H=%d)", who, o->C, o->H);
        return GP3D_E_UNSUPPORTED;
    }
    if (psC != 1 || psX % 4 || psY % 4 || psP % 4 || psB % 4 || (reinterpret_cast<uintptr_t>(planes) & 15u)) {
        gp3d_set_error("%s: planes must be channel-minor (stride_c == 1) with 16-byte aligned texels; got strides "
                       "B=%lld P=%lld C=%lld Y=%lld X=%lld", who, (long long)psB, (long long)psP, (long long)psC,
                       (long long)psY, (long long)psX);
        return GP3D_E_UNSUPPORTED;
    }
    const int64_t span = 2 * psP + (int64_t)(o->P - 1) * (psY + psX) + o->C;
    GP3D_CHECK_ARG(span < 2147483647LL, "%s: one image's planes span more than 2^31 elements", who);
    return GP3D_OK;
}

int gp3d_raymarch_forward_v3(const rm::Params& p, int planes_dtype, int mode, cudaStream_t st);   // raymarch_fwd3.cu
bool gp3d_raymarch_v3_ok(const rm::Params& p, int planes_dtype);

static bool use_legacy_v2() {
    static const int flag = [] { const char* e = getenv("GP3D_RAYMARCH_V2"); return (e != nullptr && e[0] == '1') ? 1 : 0; }();
    return flag != 0;
}

static int raymarch_forward_impl(const void* planes, int planes_dtype,
                                 int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                                 const float* ray_o, const float* ray_d, const gp3d_raymarch_cam* cam,
                                 const float* w1, const float* b1, const float* w2, const float* b2,
                                 const float* u_coarse, const float* u_fine, const float* sn_coarse, const float* sn_fine,
                                 float* rgb, float* depth, float* wsum, float* tfinal,
                                 const gp3d_raymarch_opts* opts, void* stream, const char* who) {
    int rc = gp3d_raymarch_check(planes, planes_dtype, psB, psP, psC, psY, psX, opts, who);
    if (rc != GP3D_OK) return rc;
    GP3D_CHECK_ARG(w1 && b1 && w2 && b2 && rgb && depth && wsum && tfinal, "%s: null pointer", who);
    rm::Params p{};
    p.planes = planes; p.psB = psB; p.psP = psP; p.psY = psY; p.psX = psX;
    p.ray_o = ray_o; p.ray_d = ray_d; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2;
    p.u_coarse = u_coarse; p.u_fine = u_fine; p.sn_coarse = sn_coarse; p.sn_fine = sn_fine;
    p.rgb = rgb; p.depth = depth; p.wsum = wsum; p.tfinal = tfinal;
    p.o = *opts;
    if (cam != nullptr) {
        GP3D_CHECK_ARG(cam->c2w && cam->fov, "%s: camera needs cam2world matrices and fields of view", who);
        GP3D_CHECK_ARG((cam->patch_scales == nullptr) == (cam->patch_offsets == nullptr), "%s: patch scales and offsets go together", who);
        GP3D_CHECK_ARG(cam->img_h >= 2 && cam->img_w >= 2 && (int64_t)cam->img_h * cam->img_w == opts->R, "%s: img_h * img_w must equal R (got %d x %d, R = %d)", who,
                       cam->img_h, cam->img_w, opts->R);
        p.cam_c2w = cam->c2w; p.cam_fov = cam->fov; p.patch_scale = cam->patch_scales; p.patch_offset = cam->patch_offsets;
        p.img_w = cam->img_w; p.img_h = cam->img_h;
    } else {
        GP3D_CHECK_ARG(ray_o && ray_d, "%s: null ray pointers", who);
    }
    cudaStream_t s = (cudaStream_t)stream;
    GP3D_CHECK_ARG(opts->mlp_mode >= 0 && opts->mlp_mode <= 2, "%s: mlp_mode must be 0 (fp32 SIMT), 1 (TF32) or 2 (3xTF32)", who);
    int r;
    if (opts->mlp_mode != 0 && gp3d_raymarch_v3_ok(p, planes_dtype) && !(use_legacy_v2() && cam == nullptr)) r = gp3d_raymarch_forward_v3(p, planes_dtype, opts->mlp_mode, s);
    else {
        if (cam != nullptr) {
            gp3d_set_error("%s: in-kernel ray generation needs the third-generation kernel (mlp_mode 1 / 2, strides that are multiples of 8 elements)", who);
            return GP3D_E_UNSUPPORTED;
        }
        if (opts->mlp_mode == 0) r = (planes_dtype == GP3D_F32) ? rm::launch_fwd<float>(p, s) : rm::launch_fwd<__half>(p, s);
        else r = gp3d_raymarch_forward_v2(p, planes_dtype, opts->mlp_mode, s);
    }
    if (r != 0) return r;
    GP3D_RETURN_LAUNCH();
}

extern "C" int gp3d_raymarch_forward(const void* planes, int planes_dtype,
                                     int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                                     const float* ray_o, const float* ray_d,
                                     const float* w1, const float* b1, const float* w2, const float* b2,
                                     const float* u_coarse, const float* u_fine,
                                     const float* sn_coarse, const float* sn_fine,
                                     float* rgb, float* depth, float* wsum, float* tfinal,
                                     const gp3d_raymarch_opts* opts, void* stream) {
    return raymarch_forward_impl(planes, planes_dtype, psB, psP, psC, psY, psX, ray_o, ray_d, nullptr, w1, b1, w2, b2, u_coarse, u_fine, sn_coarse, sn_fine,
                                 rgb, depth, wsum, tfinal, opts, stream, "raymarch_forward");
}

extern "C" int gp3d_raymarch_forward_cam(const void* planes, int planes_dtype,
                                         int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                                         const gp3d_raymarch_cam* cam,
                                         const float* w1, const float* b1, const float* w2, const float* b2,
                                         const float* u_coarse, const float* u_fine,
                                         const float* sn_coarse, const float* sn_fine,
                                         float* rgb, float* depth, float* wsum, float* tfinal,
                                         const gp3d_raymarch_opts* opts, void* stream) {
    GP3D_CHECK_ARG(cam != nullptr, "raymarch_forward_cam: camera is null");
    return raymarch_forward_impl(planes, planes_dtype, psB, psP, psC, psY, psX, nullptr, nullptr, cam, w1, b1, w2, b2, u_coarse, u_fine, sn_coarse, sn_fine,
                                 rgb, depth, wsum, tfinal, opts, stream, "raymarch_forward_cam");
}
