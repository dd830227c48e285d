/*
 * gp3d_b200.h -- C ABI of lib3dgp_b200.so, the sm_100a implementation of 3DGP's per-image hot path.
 *
 * Every entry point replaces one function of the reference's native plugin layer
 * (pybind11 modules built by src/torch_utils/custom_ops.py::get_plugin) or one
 * library call site on the hot path.  The reference's boundary passes torch::Tensor; this ABI is
 * the layer directly below it: plain device pointers, sizes, strides and a cudaStream_t.  The Python
 * shim in 3dgp_b200/torch_utils/custom_ops.py re-creates the reference's plugin objects
 * (`bias_act_plugin.bias_act(...)`, `upfirdn2d_plugin.upfirdn2d(...)`, ...) on top of these calls.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name starts with h_;
 *   - `stream` is a cudaStream_t (pass 0 for the legacy default stream); launches are asynchronous;
 *   - return value: 0 = ok, <0 = invalid argument (GP3D_E_*), >0 = cudaError_t of the failed launch;
 *     gp3d_last_error() returns a static human-readable string for the last failure on this thread;
 *   - dtype codes: 0 = float32, 1 = float16, 2 = bfloat16;
 *   - strides are in ELEMENTS.
 *   - callee never allocates: outputs are caller-allocated (the reference's callee-allocates contract,
 *     bias_act.cpp:55 / upfirdn2d.cpp:38, is restored by the Python shim with torch.empty).
 */
#ifndef GP3D_B200_H_
#define GP3D_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP3D_OK            0
#define GP3D_E_BADARG     -1
#define GP3D_E_UNSUPPORTED -2
#define GP3D_E_TOOLARGE   -3

#define GP3D_F32  0
#define GP3D_F16  1
#define GP3D_BF16 2

const char* gp3d_last_error(void);
int         gp3d_version(void);            /* ABI version, currently 1 */
int         gp3d_built_arch(void);         /* 100 => sm_100a */

/* ------------------------------------------------------------------------------------------------
 * bias_act  -- replaces bias_act_plugin.bias_act (reference src/torch_utils/ops/bias_act.cpp:32-90,
 * kernel bias_act.cu:23-147).  y = clamp(act(x + b[(i / stepB) % sizeB]) * gain) for grad == 0;
 * grad == 1 / 2: first / second derivative forms (x := dy resp. d_dx; xref, yref, dy as in
 * bias_act.py:179,198).  b / xref / yref / dy may be NULL ("absent", the reference's empty tensor).
 * act: 1 linear 2 relu 3 lrelu 4 tanh 5 sigmoid 6 elu 7 selu 8 softplus 9 swish (bias_act.py:22-30).
 * clamp < 0 disables clamping.  numel <= INT32_MAX (bias_act.cpp:50).
 */
int gp3d_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy,
                  void* y, int dtype, int64_t numel, int64_t sizeB, int64_t stepB,
                  int grad, int act, float alpha, float gain, float clamp, void* stream);

/* ------------------------------------------------------------------------------------------------
 * upfirdn2d -- replaces upfirdn2d_plugin.upfirdn2d (reference upfirdn2d.cpp:16-100, kernels
 * upfirdn2d.cu:29-200).  Per channel: zero-insert upsample (upx,upy) -> pad/crop (pad*) -> FIR with the
 * float32 filter f[fh][fw] (true convolution unless flip != 0) -> keep every (downx,downy)-th sample.
 * Output extent is computed by the CALLER with gp3d_upfirdn2d_out_size() (integer formula of
 * upfirdn2d.cpp:35-36) and must be >= 1.
 * x: [N,C,inH,inW] addressed through strides (NCHW or channels-last); y likewise.
 */
int gp3d_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int fsize);
int gp3d_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                   int N, int C, int inH, int inW,
                   int64_t xsN, int64_t xsC, int64_t xsH, int64_t xsW,
                   int fh, int fw, int upx, int upy, int downx, int downy,
                   int padx0, int padx1, int pady0, int pady1, int flip, float gain,
                   int outH, int outW,
                   int64_t ysN, int64_t ysC, int64_t ysH, int64_t ysW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * filtered_lrelu_act_ -- replaces filtered_lrelu_plugin.filtered_lrelu_act_ (filtered_lrelu.cpp:213-298,
 * kernel filtered_lrelu.cu:1105-1211).  In place on x [N,C,H,W] (NCHW contiguous):
 *   write_signs=1: s = sign/clamp code of x (bit0: x<0, bit1: |gain*lrelu(x)|>clamp), 2 bits per element
 *                  packed 4 per byte along W into si[N,C,sH,sW4] at offset (sx,sy); x = clamp(lrelu(x)*gain)
 *   write_signs=0, si != NULL: x *= (code==0 ? gain : code==1 ? gain*slope : 0) using stored codes
 *   si == NULL: plain x = clamp(lrelu(x)*gain).
 */
int gp3d_filtered_lrelu_act(void* x, uint8_t* si, int dtype, int N, int C, int H, int W,
                            int sH, int sW4, int sx, int sy, float gain, float slope, float clamp,
                            int write_signs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * filtered_lrelu -- replaces filtered_lrelu_plugin.filtered_lrelu (filtered_lrelu.cpp:16-209, kernels filtered_lrelu.cu:139-1099):
 * bias -> zero-insert up-sample + pad + FIR(fu) * up^2 -> gain * lrelu -> clamp -> FIR(fd) + decimate in ONE kernel; the up-sampled
 * intermediate exists only in shared memory.  SEPARABLE filters (fu, fd given as 1-D taps, applied along both axes); any other
 * configuration returns GP3D_E_UNSUPPORTED and the caller takes the generic route (the reference's return code -1, filtered_lrelu.cpp:50-55).
 *   x [N,C,xH,xW], y [N,C,yH,yW]: NCHW contiguous, dtype f32 / f16; b: [C] in x's dtype or NULL; px0 / py0: leading padding w.r.t. the up-sampled
 *   signal (negative = crop); yW = (xW*up + px0 + px1 - (fu_taps-1) - (fd_taps-1) + down-1) / down (the caller computes it, filtered_lrelu.cpp:72-79).
 *   s [N,C,sH,sW4] uint8: four 2-bit codes per byte (0 positive, 1 negative, 2 clamped) of intermediate element (u, v) at sign coordinate
 *   (u + sx, v + sy); write_signs: produced (bytes up to sw_limit per row); read_signs: consumed instead of evaluating the activation
 *   (x gain, x gain*slope, x 0) -- the backward pass (filtered_lrelu.py:252-263).  flip_filter: 0 convolution, 1 correlation.
 */
int gp3d_filtered_lrelu(const void* x, const float* fu, const float* fd, const void* b, uint8_t* s, void* y, int dtype,
                        int N, int C, int xH, int xW, int yH, int yW, int fu_taps, int fd_taps, int up, int down, int px0, int py0,
                        int sH, int sW4, int sx, int sy, int sw_limit, float gain, float slope, float clamp, int flip_filter,
                        int write_signs, int read_signs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused tri-plane ray-march (forward).  Replaces the ~25 torch ops of ImportanceRenderer.forward
 * (reference src/training/tri_plane_renderer.py:126-170) together with simple_tri_plane_renderer
 * (:560-588, ATen grid_sampler_2d), TriPlaneMLP.forward (networks_epigraf.py:46-68),
 * sample_stratified (:208-235), sample_importance/sample_pdf (:237-295), unify_samples (:196-206) and
 * ClassicalRayMarcher.forward (:353-405).  No per-sample tensor is written to HBM.
 *
 * planes   : [B][3 planes][C=32 ch][P][P] addressed by element strides (psB, psP, psC, psY, psX).
 *            Fast path: psC == 1 (channel-minor / channels-last storage).  dtype f32 or f16.
 * ray_o/d  : [B][R][3] float32 contiguous.
 * w1,b1    : first FC layer raw parameters [H=64][32], [64]; w2,b2: [4][64],[4] (float32, row-major);
 *            the reference's runtime weight gains 1/sqrt(fan_in) (layers.py:39,47) are applied inside.
 * u_coarse : [B][R][N] uniform(0,1) stratification jitter (tri_plane_renderer.py:225), or NULL => Philox.
 * u_fine   : [B][R][N] uniform(0,1) inverse-CDF variates (:279), or NULL => Philox.
 * sn_coarse/sn_fine : optional [B][R][N] standard-normal density noise (:185-186), scaled by noise_std.
 * outputs  : rgb [B][R][3], depth [B][R], wsum [B][R], tfinal [B][R]  (float32).
 */
typedef struct gp3d_raymarch_opts {
    int   B, R, N;            /* batch, rays per image, samples per pass (coarse == fine == N, N <= 64) */
    int   P;                  /* plane resolution */
    int   C;                  /* feature channels per plane (32) */
    int   H;                  /* MLP hidden width (64) */
    float ray_start, ray_end; /* cfg.camera.ray.{start,end} */
    float box_half;           /* box_size / 2 = cfg.camera.cube_scale; coords are divided by it (tri_plane_renderer.py:576) */
    float noise_std;          /* rendering_options['density_noise'] */
    int   use_inf_depth;      /* last delta 1e10 (1) or 1e-3 (0), tri_plane_renderer.py:356 */
    int   last_back;          /* :386-387 */
    int   white_back_end_idx; /* :392-395 */
    int   clamp_mode;         /* 0 softplus, 1 relu (:359-364) */
    int   mlp_mode;           /* 0 = fp32 SIMT, 1 = TF32 mma.sync, 2 = 3xTF32 error-compensated mma.sync */
    uint64_t seed, offset;    /* Philox stream when u_* are NULL */
} gp3d_raymarch_opts;

int gp3d_raymarch_forward(const void* planes, int planes_dtype,
                          int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                          const float* ray_o, const float* ray_d,
                          const float* w1, const float* b1, const float* w2, const float* b2,
                          const float* u_coarse, const float* u_fine,
                          const float* sn_coarse, const float* sn_fine,
                          float* rgb, float* depth, float* wsum, float* tfinal,
                          const gp3d_raymarch_opts* opts, void* stream);

/* The same render with the rays GENERATED IN THE KERNEL from a pinhole camera per image -- replaces compute_cam2world_matrix's consumers
 * sample_rays (tri_plane_renderer.py:487-527) + ImportanceRenderer.forward in one launch; no ray tensors in HBM.  Rays form an img_h x img_w grid
 * (R == img_h * img_w, ray = y * img_w + x): x in linspace(-1, 1, img_w), y in linspace(1, -1, img_h), optional patch transform
 * x' = (x + 1) * scale_x - 1 + 2 * offset_x (:511-512), z = -1 / tan(fov / 2), direction = cam2world[:3,:3] . normalize(x, y, z), origin = cam2world[:3, 3].
 * CTAs own 4 x 4 pixel tiles.  Needs mlp_mode 1 / 2 and plane strides that are multiples of 8 elements. */
typedef struct gp3d_raymarch_cam {
    const float* c2w;            /* [B][4][4] row-major cam2world (rendering_utils.py:194-218) */
    const float* fov;            /* [B] field of view, degrees */
    const float* patch_scales;   /* [B][2] or NULL */
    const float* patch_offsets;  /* [B][2] or NULL */
    int img_h, img_w;
} gp3d_raymarch_cam;

int gp3d_raymarch_forward_cam(const void* planes, int planes_dtype,
                              int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                              const gp3d_raymarch_cam* cam,
                              const float* w1, const float* b1, const float* w2, const float* b2,
                              const float* u_coarse, const float* u_fine,
                              const float* sn_coarse, const float* sn_fine,
                              float* rgb, float* depth, float* wsum, float* tfinal,
                              const gp3d_raymarch_opts* opts, void* stream);

/* The ray generator on its own (sample_rays, tri_plane_renderer.py:487-527): ray_o, ray_d [B][img_h * img_w][3].  Used by the backward pass, which
 * takes explicit rays. */
int gp3d_generate_rays(const float* c2w, const float* fov, const float* patch_scales, const float* patch_offsets,
                       int B, int img_h, int img_w, float* ray_o, float* ray_d, void* stream);

/* Backward of the above.  Recomputes the forward per ray block (nothing was saved), then back-propagates
 * g_rgb [B][R][3] and g_depth [B][R] into
 *   g_planes (same strides as planes, float32, ACCUMULATED with red.global.add -- caller zero-fills),
 *   g_w1,g_b1,g_w2,g_b2 (accumulated, caller zero-fills), g_ray_o / g_ray_d [B][R][3] (overwritten; may be NULL).
 * The importance-sampled depths are constants (tri_plane_renderer.py:241,254 no_grad + detach).
 * Requires the SAME u_* / sn_* inputs (or the same Philox seed/offset) as the forward call.
 */
int gp3d_raymarch_backward(const void* planes, int planes_dtype,
                           int64_t psB, int64_t psP, int64_t psC, int64_t psY, int64_t psX,
                           const float* ray_o, const float* ray_d,
                           const float* w1, const float* b1, const float* w2, const float* b2,
                           const float* u_coarse, const float* u_fine,
                           const float* sn_coarse, const float* sn_fine,
                           const float* g_rgb, const float* g_depth,
                           float* g_planes, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                           float* g_ray_o, float* g_ray_d,
                           const gp3d_raymarch_opts* opts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Elementwise fusions around the modulated convolution (reference networks_stylegan2.py:67-76, 142-144):
 *   gp3d_modulate   : y[n,c,h,w] = x[n,c,h,w] * s[n,c]                         (x * styles, :68)
 *   gp3d_demod_act  : y = clamp(lrelu(x * d[n,c] + noise[n?,h,w] * nstr + b[c]) * gain)
 *                     == fma.fma (:71) followed by bias_act.bias_act (:144) in one pass.
 * NCHW-contiguous or channels-last (cl != 0), float32/16/bf16.
 */
int gp3d_modulate(const void* x, const void* s, void* y, int dtype, int N, int C, int HW, int cl, void* stream);
/* uint8 conversion of generated images (metric_utils.py:313 `(img * 127.5 + 128).clamp(0, 255).to(torch.uint8)`, training_loop.py:23-49): the first
 * Cy channels of float32 x [N, Cx, H, W] (element strides sN, sC, sH, sW: NCHW or channels-last) -> NCHW-contiguous uint8 y [N, Cy, H, W];
 * y = (uint8) clamp(x * scale + shift, 0, 255), truncating like torch's cast.  W % 4 == 0. */
int gp3d_to_uint8(const float* x, uint8_t* y, int N, int Cx, int Cy, int H, int W, int64_t sN, int64_t sC, int64_t sH, int64_t sW,
                  float scale, float shift, void* stream);
int gp3d_demod_act(const void* x, const void* d, const void* noise, int noise_per_sample, const void* b,
                   void* y, int dtype, int N, int C, int HW, int cl,
                   int act, float alpha, float gain, float clamp, void* stream);

/* Backward halves of the fused modulated-conv layer, channel-minor float32 [N][HW][C] (C % 4 == 0):
 *   gp3d_demod_act_bwd : given dy and the saved OUTPUT y of gp3d_demod_act (no clamp): dt = dy*gain*act'(y); dc = dt*d[n,c];
 *                        g_b[c] += sum dt; g_d[n,c] += sum_hw dt*c (c rebuilt from y); g_ns += sum dt*noise  (noise = UNSCALED image, *noise_scale = its device-side strength).  (fma.py:33-53 +
 *                        bias_act.py:157-172 in one pass; g_* are ACCUMULATED, caller zero-fills; any of d/noise/b/g_* may be NULL)
 *   gp3d_modulate_bwd  : dx = dxs * s[n,c];  g_s[n,c] += sum_hw dxs * x           (backward of x * styles, networks_stylegan2.py:68)
 */
int gp3d_demod_act_bwd(const float* dy, const float* y, const float* d, const float* noise, const float* noise_scale, int noise_per_sample, const float* b,
                       float* dc, float* g_d, float* g_b, float* g_ns, int N, int HW, int C, int act, float alpha, float gain, void* stream);
/* gp3d_demod_act_bwd writing dc either as float32 (dc != NULL) or directly as the bf16 (hi, lo) operand pair of the tensor-core
 * input-gradient / weight-gradient kernels (dc_hi, dc_lo != NULL, [N][HW][C_pad] with channels [C, C_pad) zero-filled; C_pad - C <= C). */
int gp3d_demod_act_bwd_split(const float* dy, const float* y, const float* d, const float* noise, const float* noise_scale, int noise_per_sample,
                             const float* b, float* dc, void* dc_hi, void* dc_lo, int C_pad, float* g_d, float* g_b, float* g_ns,
                             int N, int HW, int C, int act, float alpha, float gain, void* stream);
/* same with bias_act's clamp (gradient passes only where |y| < clamp; clamp <= 0: none) and an optional low-order half
 * (dc_lo == NULL: single-term bf16 operand) -- the backward of Conv2dLayer's bias_act in the discriminator. */
int gp3d_act_bwd_split(const float* dy, const float* y, float* dc, void* dc_hi, void* dc_lo, int C_pad, float* g_b,
                       int N, int HW, int C, int act, float alpha, float gain, float clamp, void* stream);
int gp3d_modulate_bwd(const float* dxs, const float* x, const float* s, float* dx, float* g_s, int N, int HW, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient all-reduce epilogue (reference training_loop.py:340-341): in place
 *   g = nan_to_num(g / world, nan=0, posinf=1e5, neginf=-1e5)    over a flat float32 buffer.
 */
int gp3d_grad_epilogue(float* g, int64_t numel, float inv_world, float posinf, float neginf, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step fused with the all-reduce epilogue and the G_ema update (reference training_loop.py:340-341 nan_to_num,
 * :346 opt.step() = torch.optim.Adam built at :190-205, :357-364 p_ema.copy_(p.lerp(p_ema, beta))): one pass over a
 * module's flat float32 storage.  p, g, v (and m, ema when given) are parallel buffers of `numel` floats, numel % 1024 == 0,
 * every parameter tensor starting on a 1024-element boundary.
 *   g' = nan_to_num(g * grad_scale, 0, posinf, neginf);  m = lerp(m, g', 1-beta1) (m == NULL iff beta1 == 0);
 *   v = v*beta2 + (1-beta2) g'^2;  p -= step_size * m / (sqrt(v)/bc2_sqrt + eps);  ema = lerp(p, ema, ema_beta) if ema != NULL
 * with step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t), one_minus_beta* = 1 - beta* computed by the caller in double
 * precision and rounded once (torch.optim.Adam's python scalars).
 * blk_seg / seg_desc (both or neither): blk_seg[numel/1024] = tensor index of each block (-1: padding), seg_desc[4*i] =
 * {step_size_i, bc2_sqrt_i, active_i, 0}: per-tensor step counts and "received no gradient this phase => untouched"
 * (torch.optim skips parameters whose .grad is None); the EMA update still applies to inactive tensors.
 */
int gp3d_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, int64_t numel,
                       float grad_scale, float posinf, float neginf, float beta1, float beta2, float one_minus_beta1, float one_minus_beta2, float eps,
                       float step_size, float bc2_sqrt, float ema_beta,
                       const int* blk_seg, const float* seg_desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 / TMEM (sm_100a): implicit-GEMM convolutions -- bf16 / fp16 operands, fp32 accumulate in TMEM, TMA-fed
 * 128B-swizzled smem tiles (the conv front-end lays the im2col tiles out through TMA; no im2col buffer exists). */

/* 3x3 / 1x1 stride-1 "same" convolution as an implicit GEMM on tcgen05 (NHWC bf16 activations,
 * weights [Cout][kh][kw][Cin] bf16, fp32 NHWC output; ksize in {1, 3, 5}).  Replaces the cuDNN call of
 * conv2d_gradfix.py:113 for the hot shapes.  Cin % 64 == 0, Cout % 128 == 0 (or Cout in {64, 96}); any N, H, W.
 */
int gp3d_conv2d_nhwc_bf16(const void* x, const void* w, float* y, int N, int H, int W, int Cin, int Cout,
                          int ksize, int accumulate, void* stream);

/* Error-compensated variant for fp32 parity ("bf16x3"): x = xh + xl, w = wh + wl are bf16 pairs produced by
 * gp3d_split_bf16; y (+)= xh*wh + xh*wl + xl*wh, all three products accumulated in the same fp32 TMEM tile
 * (relative error ~2^-16 instead of 2^-9).  Same shapes / constraints as gp3d_conv2d_nhwc_bf16.
 */
int gp3d_conv2d_nhwc_bf16x3(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                            int Cin, int Cout, int ksize, int accumulate, void* stream);

/* bf16x3 'same' convolution with the rest of the modulated-conv layer fused into the TMEM -> HBM epilogue
 * (reference networks_stylegan2.py:71 `fma(x, dcoefs, noise)` and :144 `bias_act(x, b, act, gain)`, no clamp):
 *   y[n][oy][ox][co] = act(conv * dcoef[n][co] + noise[n?][oy][ox] + bias[co]) * gain
 * dcoef / noise / bias may each be NULL; noise is already multiplied by noise_strength, [N][H][W] if noise_per_sample else [H][W].
 */
typedef struct gp3d_conv_epilogue {
    const float* dcoef;
    const float* noise;
    const float* bias;
    int noise_per_sample;
    int act;            /* 1 linear, 3 lrelu */
    float alpha, gain;
    float clamp;        /* > 0: clamp the result to [-clamp, clamp] (bias_act's clamp, layers.py:239); <= 0: none */
} gp3d_conv_epilogue;

/* Descriptor form of the tap convolution (every form of conv2d_resample.py:93-141 on one kernel family), including the two-term precision:
 *   xl == NULL, wl == NULL, w_format 0 : single bf16 product                                (1 MMA  / product)
 *   xl, wl given,           w_format 0 : bf16x3  xh*wh + xh*wl + xl*wh                      (3 MMAs / product, ~2^-16)
 *   xl given, wl == NULL,   x_format = w_format = 1 : x2w16   (xh + xl) * w16, fp16 activation pair x fp16 weight
 *                                                                                            (2 MMAs / product, weight rounding 2^-12)
 * x_format must equal w_format: tcgen05.mma.kind::f16 faults (illegal instruction) on sm_100a when the A and B element formats differ. */
typedef struct gp3d_conv_desc {
    const void* xh; const void* xl;      /* activations, bf16 [N][H][W][Cin] (+ low-order half) */
    const void* wh; const void* wl;      /* weights [Cout][num_slabs][Cin]: bf16 (+ low-order half), or fp16 when w_format == 1 */
    int w_format;                        /* 0 bf16, 1 fp16 */
    int x_format;                        /* 0 bf16, 1 fp16 (== w_format): fp16 x fp16 is the arithmetic class of the reference's fp16 discriminator blocks */
    float* y;                            /* [N][Hout][Wout][Cout] */
    int N, H, W, Cin, Cout, num_slabs;
    int ntaps; const int* taps;          /* ntaps x (dy, dx, slab), host memory */
    int in_stride, HoP, WoP, Hout, Wout, osy, osx, oy0, ox0, accumulate;
    const gp3d_conv_epilogue* epi;       /* optional fused epilogue */
} gp3d_conv_desc;
int gp3d_conv_nhwc(const gp3d_conv_desc* desc, void* stream);
/* Stride-2 transposed 3x3 convolution, padding 0 (what conv2d_resample.py:113-126 runs as conv_transpose2d for the up-sampling layers):
 *   y[n][2i+ky][2j+kx][co] += x[n][i][j][ci] * w[co][ky*3+kx][ci],   y = [N][2H+1][2W+1][Cout] float32 (fully written),
 * evaluated as its four polyphase tap convolutions (1, 2, 2 and 4 taps) by ONE launch: the phases of a pixel window are scheduled back to back, so the
 * activation tensor is read from DRAM once instead of four times.  Operands / precision as gp3d_conv_desc (xl / wl / formats). */
int gp3d_conv_transpose_s2_nhwc(const void* xh, const void* xl, const void* wh, const void* wl, int w_format, int x_format, float* y,
                                int N, int H, int W, int Cin, int Cout, void* stream);
/* Tuning switch: 256-wide output-channel tiles for the three-term (bf16x3) form when Cout % 256 == 0 (two-stage ring of 96 KB stages).
 * Returns the previous setting. */
int gp3d_conv_set_wide3(int on);
int gp3d_conv2d_nhwc_bf16x3_act(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                                int Cin, int Cout, int ksize, const gp3d_conv_epilogue* epi, void* stream);
/* same epilogue for either precision: xl / wl both NULL = single-term bf16 (the discriminator's low-precision blocks: Conv2dLayer's
 * conv + bias_act, layers.py:228-241, dcoef / noise NULL), both given = bf16x3. */
int gp3d_conv2d_nhwc_act(const void* xh, const void* xl, const void* wh, const void* wl, float* y, int N, int H, int W,
                         int Cin, int Cout, int ksize, const gp3d_conv_epilogue* epi, void* stream);

/* 4x4 FIR, up = down = 1, dense channel-minor float32 [N][H][W][C] (C % 32 == 0), input window staged by TMA: the filter after the
 * up-sampling convolution (conv2d_resample.py:119-126) and its adjoint.  out = (H + pady0 + pady1 - 3) x (W + padx0 + padx1 - 3).
 * f: DEVICE pointer to the 4x4 filter as upfirdn2d.setup_filter stores it; flip / gain as in gp3d_upfirdn2d.
 * Exactly one output form:  y (float32; with epi != NULL the modulated-conv epilogue act(v*dcoef + noise + bias)*gain is applied to the
 * filtered value v, networks_stylegan2.py:71,144)  or  (hi, lo) = bf16 pair of v for the tensor-core gradient kernels. */
int gp3d_fir4_nhwc(const float* x, const float* f, int flip, float gain, int N, int H, int W, int C,
                   int padx0, int padx1, int pady0, int pady1, float* y, void* hi, void* lo,
                   const gp3d_conv_epilogue* epi, void* stream);

/* General tap convolution on the same tcgen05 pipeline -- the building block of the strided forms:
 *   y[n][iy*osy+oy0][ix*osx+ox0][co] (+)= sum_t sum_ci x[n][iy*in_stride+dy_t][ix*in_stride+dx_t][ci] * w[co][slab_t][ci]
 * for (iy, ix) in [0,HoP) x [0,WoP); out-of-range input pixels read as zero.  h_taps is a HOST array of ntaps x (dy, dx, slab).
 *   - stride-2 transposed conv (G's up-sampling conv0, conv2d_resample.py:112-126; the input gradient of D's stride-2 convs) =
 *     four polyphase launches (1, 2, 2 and 4 taps) with osy = osx = 2 and (oy0, ox0) in {0,1}^2: no zero-stuffed FLOPs;
 *   - stride-2 conv (D's down-sampling conv1, conv2d_resample.py:106-109; the input gradient of G's conv0) = one launch with in_stride = 2.
 * xl / wl may be NULL (single-term bf16) or both given (bf16x3).
 */
int gp3d_conv_taps_nhwc(const void* xh, const void* xl, const void* wh, const void* wl, float* y,
                        int N, int H, int W, int Cin, int Cout, int num_slabs, int ntaps, const int* h_taps, int in_stride,
                        int HoP, int WoP, int Hout, int Wout, int osy, int osx, int oy0, int ox0, int accumulate, void* stream);

/* Weight gradient of the tap convolutions as a tcgen05 GEMM over pixels (replaces aten.convolution_backward's weight output,
 * conv2d_gradfix.py:141-151):
 *   dW[co][slab_t][ci] += sum_{n,iy,ix} dy[n][iy*sa+ay_t][ix*sa+ax_t][co] * x[n][iy*sb+by_t][ix*sb+bx_t][ci],  (iy,ix) in [0,HoP)x[0,WoP)
 * dy [N][Hd][Wd][Cout], x [N][Hx][Wx][Cin] bf16 NHWC (hi, optional lo pair for bf16x3); h_taps: HOST array ntaps x (ay, ax, by, bx, slab);
 * dW float32 [Cout][num_slabs][Cin], ACCUMULATED (caller zero-fills).  Cin % 64 == 0, Cout % 64 == 0; up to 25 taps.
 */
int gp3d_wgrad_taps_nhwc(const void* dyh, const void* dyl, const void* xh, const void* xl, float* dW,
                         int N, int Hd, int Wd, int Cout, int Hx, int Wx, int Cin, int num_slabs,
                         int ntaps, const int* h_taps, int sa, int sb, int HoP, int WoP, void* stream);
/* Same with the element format of the single-term form selectable (0 bf16, 1 fp16; dy_format must equal x_format). */
int gp3d_wgrad_taps_nhwc_fmt(const void* dyh, const void* dyl, const void* xh, const void* xl, int dy_format, int x_format, float* dW,
                             int N, int Hd, int Wd, int Cout, int Hx, int Wx, int Cin, int num_slabs,
                             int ntaps, const int* h_taps, int sa, int sb, int HoP, int WoP, void* stream);

/* fp32 / fp16 -> bf16 hi (+ lo) split with optional per-(n, c) modulation (x * styles, networks_stylegan2.py:68), channel-minor
 * (NHWC) tensors: hi = bf16(x * s[n][c]); lo = bf16(x * s[n][c] - hi) (lo may be NULL).  s may be NULL.  src_dtype: GP3D_F32 / F16.
 */
int gp3d_split_bf16(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, void* stream);
/* same, with the outputs zero-padded to C_out >= C channels (96-channel toRGB tensors -> 128 so that they fill whole 64-channel TMA blocks) */
int gp3d_split_bf16_pad(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, int C_out, void* stream);
/* Same with the element format of `hi` selectable: hi_format 0 = bf16, 1 = fp16; lo (optional) has the same format. */
int gp3d_split_pad(const void* x, int src_dtype, const float* s, void* hi, void* lo, int N, int HW, int C, int C_out, int hi_format, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GP3D_B200_H_ */
