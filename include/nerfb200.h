/*
 * nerfb200.h — C ABI of libnerfb200.so, the B200 (sm_100a) implementation of the NeRF
 * train/render hot path of sarphiv/nerf-experiments.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - tensors are dense row-major fp32 unless stated; sizes are element counts;
 *   - every call is stream-ordered on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing, never synchronises and never throws; it returns 0 on success or a
 *     NERFB200_ERR_* code, with the text retrievable through nerfb200_last_error();
 *   - workspaces (activation / gradient stashes, sign-bit masks) are caller-owned; the fused-field
 *     entry points state their sizes (nerfb200_mlp_workspace_bytes, nerfb200_garf_workspace_bytes).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef NERFB200_H_
#define NERFB200_H_

#include <stddef.h>
#include <stdint.h>

#include "nerfb200_mlp.h"  /* tile-program structs passed across the ABI */
#include "nerfb200_garf.h" /* tile programs of the GARF networks */

#ifdef __cplusplus
extern "C" {
#endif

#define NERFB200_ABI_VERSION 6   /* 3: adam_step_dev replaces adam_step_sched; GARF fused field; render_rays; trans_cdf / prop_loss. 4: NbWgradItem z duty. 5: gauss_width_grad. 6: NbWgradItem.x2_slab */

enum {
  NERFB200_OK = 0,
  NERFB200_ERR_INVALID_ARGUMENT = 1,
  NERFB200_ERR_CUDA = 2,
  NERFB200_ERR_UNSUPPORTED = 3
};

/* Text of the last error raised on the calling thread ("" if none). */
const char* nerfb200_last_error(void);
int nerfb200_abi_version(void);
/* Number of kernels this library has launched since load (process-wide counter). */
long long nerfb200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * a1. Uniform t-sampling.
 * Replaces NerfInterpolation._sample_t_stratified_uniform + _get_intervals
 * (barf/model_interpolation.py:135-180, :114-132).
 *   s_k     = linspace(near, far - D, S)[k],  D = (far - near) / S
 *   s_k    += jitter[r,k] * D                 if jitter != NULL   ("stratified_uniform")
 *   s_k    += offset_u[r] * D * offset_size   if offset_u != NULL and offset_size != 0
 *   t_start = s ; t_end[k] = s[k+1] ; t_end[S-1] = far
 * jitter: (B,S) uniforms in [0,1) or NULL ("equidistant"); offset_u: (B,) uniforms or NULL.
 * near/far/offset_size are doubles because the reference evaluates them as Python floats
 * before ATen narrows to fp32 (bit-exact linspace).
 */
int nerfb200_sample_uniform(double near_t, double far_t, int B, int S, const float* jitter,
                            const float* offset_u, double offset_size, float* t_start,
                            float* t_end, void* stream);

/* ------------------------------------------------------------------------------------------
 * a10. Alpha compositing along rays.
 * Replaces NerfInterpolation._render_rays (barf/model_interpolation.py:316-353) when
 * flavour == NERFB200_COMPOSITE_BARF:
 *   b = ((-sigma*delta)*3)*fl32(1/3); alpha = 1-exp(b); T_i = exp(sum_{j<i} b_j);
 *   w = T*alpha; rgb_out = sum_i w_i c_i
 * and the nerfacc.rendering arithmetic used by GarfModel.forward (garf/model_garf.py:223-236)
 * when flavour == NERFB200_COMPOSITE_NERFACC: b = -sigma*delta, plus opacity = sum w and
 * depth = sum w*t_mid / max(opacity, FLT_EPSILON).
 *   sigma, delta: (B,S); rgb: (B,S,3); t_mid: (B,S) or NULL (depth then not produced)
 *   out_rgb: (B,3); out_w: (B,S) or NULL; out_opacity, out_depth: (B,) or NULL.
 */
enum { NERFB200_COMPOSITE_BARF = 0, NERFB200_COMPOSITE_NERFACC = 1 };

int nerfb200_composite_fwd(const float* sigma, const float* delta, const float* rgb,
                           const float* t_mid, int B, int S, int flavour, float* out_rgb,
                           float* out_w, float* out_opacity, float* out_depth, void* stream);

/* Backward of the above (autograd of barf/model_interpolation.py:340-353).
 *   g_rgb: (B,3) dL/d out_rgb; g_w: (B,S) dL/d out_w or NULL;
 *   g_opacity, g_depth: (B,) or NULL (NERFACC flavour only; t_mid required for g_depth)
 *   d_sigma: (B,S); d_rgb: (B,S,3).  delta receives no gradient in the reference's use.
 */
int nerfb200_composite_bwd(const float* sigma, const float* delta, const float* rgb,
                           const float* t_mid, const float* g_rgb, const float* g_w,
                           const float* g_opacity, const float* g_depth, int B, int S,
                           int flavour, float* d_sigma, float* d_rgb, void* stream);

/* ------------------------------------------------------------------------------------------
 * a11. Deterministic pdf resampling.
 * Replaces NerfInterpolation._sample_t_pdf_weighted (barf/model_interpolation.py:193-277).
 *   t_coarse, weights, delta_coarse: (B,Sc); outputs t_start, t_end: (B,Sf), Sf > Sc.
 *   counts_out: (B,Sc) int32 or NULL — the per-bin sample counts n_i (for parity checks).
 *   fail_flag: device int32 (caller zeroes it): set to 1 if any ray violates the reference's
 *     postcondition (n<0 or sum n != Sf, barf/model_interpolation.py:235).  The reference
 *     then replaces the WHOLE batch by equidistant samples with offset -1 (:275);
 *     nerfb200_resample_fallback applies exactly that on the device, without a host sync,
 *     and is a no-op when the flag is 0.
 */
int nerfb200_resample_alloc(const float* t_coarse, const float* weights,
                            const float* delta_coarse, int B, int Sc, int Sf, double far_t,
                            float* t_start, float* t_end, int32_t* counts_out,
                            int32_t* fail_flag, void* stream);

int nerfb200_resample_fallback(const int32_t* fail_flag, double near_t, double far_t, int B,
                               int Sf, const float* offset_u, float* t_start, float* t_end,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * a12. Inverse-CDF importance resampling (nerfacc importance_sampling as used by
 * PropNetEstimator.sampling, call sites garf/model_garf.py:210-220; the nerfacc source is not
 * part of the reference tree — algorithm restated in oracle/ref_nerfacc.py).
 *   edges: (B,Sc+1) interval edges in s-space; cdf: (B,Sc+1) non-decreasing, cdf[0]=0;
 *   u_ray: (B,) per-ray jitter in [0,1) or NULL (=> 0.5, the non-stratified case);
 *   out_edges: (B,Sf+1) new interval edges; out_idx: (B,Sf) int32 bin index of every new
 *   sample centre (bit-exact vs oracle) or NULL.
 */
int nerfb200_resample_icdf(const float* edges, const float* cdf, const float* u_ray, int B,
                           int Sc, int Sf, float* out_edges, int32_t* out_idx, void* stream);

/* ------------------------------------------------------------------------------------------
 * a12 (rest of the chain). What GarfModel.forward obtains from nerfacc besides the importance
 * sampling above (garf/model_garf.py:210-230,257; restated in oracle/ref_garf.py — parity unpinned):
 *   lindisp_intervals: t = 1 / (s/far + (1-s)/near) for normalised edges s (B,E); any of t_edges (B,E),
 *     t_start / t_end / delta / t_mid (B,E-1) may be NULL.
 *   trans_cdf_fwd: trans (B,S) [or NULL] = exp(-exclusive cumsum(sigma * (t_end - t_start))),
 *     cdf (B,S+1) = 1 - [trans, 0]  (render_transmittance_from_density at the proposal level).
 *   trans_cdf_bwd: d_sigma (B,S) from g_cdf (B,S+1) and / or g_trans (B,S) (either may be NULL).
 *   prop_loss: *loss += scale * sum clip(w - w_outer, 0)^2 / (w + eps) over the query bins
 *     (PropNetEstimator.compute_loss, Mip-NeRF 360 eq. 13; w from cdf_query (B,Sq+1), the outer bound
 *     from cdf_key (B,Sk+1) at the searchsorted bounds of t_query in t_key); d_cdf_key (B,Sk+1) or
 *     NULL receives scale * d(sum)/d(cdf_key). The caller zeroes *loss; scale = 1 / (B * Sq) for the mean.
 */
int nerfb200_lindisp_intervals(const float* s_edges, float near, float far, int B, int E,
                               float* t_edges, float* t_start, float* t_end, float* delta,
                               float* t_mid, void* stream);
int nerfb200_trans_cdf_fwd(const float* sigma, const float* t_start, const float* t_end, int B, int S,
                           float* out_trans, float* out_cdf, void* stream);
int nerfb200_trans_cdf_bwd(const float* sigma, const float* t_start, const float* t_end,
                           const float* g_cdf, const float* g_trans, int B, int S, float* d_sigma,
                           void* stream);
int nerfb200_prop_loss(const float* t_query, const float* cdf_query, const float* t_key,
                       const float* cdf_key, int B, int Sq, int Sk, float eps, float scale,
                       float* loss, float* d_cdf_key, void* stream);

/* Gaussian-width gradients of the fused GARF networks without the pre-activation stash (a6 inside a7 / a8:
 * the parameter gradient of barf/gaussian.py:8-63's autograd): for a layer z = W x + b followed by
 * y = exp(-z^2 (s^2 + 1e-6)), sum over samples of z_n dz_n equals sum_c W[n,c] dW[n,c] + b_n db_n, so
 *   d_params[g_off + n] += sign * (W[n,:] . dW[n,:] + b_n db_n) * s_n / (s_n^2 + 1e-6)
 * from the weight / bias gradients already in d_params. The fused field calls it with sign = -1 before its
 * weight-gradient kernel and +1 after it, so that whatever the buffer held before cancels (the map is
 * linear). n_features = sum of out_f over the layers. */
typedef struct {
  int64_t w_off, b_off, g_off;   /* float offsets of weight (out_f, in_f), bias (out_f), inverse std (out_f) */
  int32_t in_f, out_f;
} NbGaussLayer;
int nerfb200_gauss_width_grad(const NbGaussLayer* layers_dev, int n_layers, long long n_features,
                              const float* params, float* d_params, float sign, void* stream);

/* ------------------------------------------------------------------------------------------
 * a13. Camera extrinsics (per-image so(3) rotation + translation).
 * Replaces CameraExtrinsics.forward / forward_origins / so3_to_SO3
 * (barf/model_camera_extrinsics.py:22-85): R_i = exp([w_i]x), o' = o + t_i, d' = R_i d.
 *   rotation, translation: (n_images,3); img_idx: (B,) int32; o, d: (B,3)
 *   out_o, out_d: (B,3); out_R: (B,3,3) or NULL; out_t: (B,3) or NULL.
 */
int nerfb200_pose_fwd(const float* rotation, const float* translation, const int32_t* img_idx,
                      const float* o, const float* d, int B, int n_images, float* out_o,
                      float* out_d, float* out_R, float* out_t, void* stream);

/* Backward: accumulates (+=) into d_rotation, d_translation (n_images,3), which the caller
 * zero-initialises; g_o, g_d: (B,3) gradients w.r.t. out_o, out_d. */
int nerfb200_pose_bwd(const float* rotation, const int32_t* img_idx, const float* d,
                      const float* g_o, const float* g_d, int B, int n_images,
                      float* d_rotation, float* d_translation, void* stream);

/* so3_to_SO3 alone (barf/model_camera_extrinsics.py:22-43): so3 (n,3) -> R (n,3,3). */
int nerfb200_so3_to_SO3(const float* so3, int n, float* out_R, void* stream);

/* ------------------------------------------------------------------------------------------
 * Stand-alone positional encodings (fp32 in/out); PeCfg is NbPeCfg of csrc/mlp.h.
 * Replaces forward() of FourierFeatures / BarfPositionalEncoding / IntegratedFourierFeatures /
 * IntegratedBarfFourierFeatures (barf/positional_encodings.py:43-57, :124-148, :170-240, :266-282).
 *   pos, dir: (N,3); pixel_width, t_start, t_end: (N,) (integrated encoding only, else NULL)
 *   alpha: device scalar of the BARF mask or NULL; out: (N, output_dim).
 */
int nerfb200_pe_fwd(const NbPeCfg* cfg_host, const float* alpha, const float* pos,
                    const float* dir, const float* pixel_width, const float* t_start,
                    const float* t_end, long long N, float* out, void* stream);
int nerfb200_pe_bwd(const NbPeCfg* cfg_host, const float* alpha, const float* pos,
                    const float* dir, const float* pixel_width, const float* t_start,
                    const float* t_end, const float* g_out, long long N, float* d_pos,
                    float* d_dir, void* stream);

/* ------------------------------------------------------------------------------------------
 * a2-a9. Fused field: query positions + positional encodings + radiance / proposal MLP.
 * Replaces NerfInterpolation._compute_positions + NerfModel.forward and their autograd
 * (barf/model_interpolation.py:288-312, barf/model_interpolation_architecture.py:96-141).
 * The network is described by a tile program (NbProgram, csrc/mlp.h) compiled on the host; the
 * structs are plain C and are passed from HOST memory.
 *
 * nerfb200_mlp_pack   refreshes the packed bf16 weight images / padded fp32 biases from the flat
 *                     fp32 parameter buffer (once per optimiser step).
 * nerfb200_mlp_fwd    sigma (N,), rgb (N,3). With stash/masks non-NULL (training) it also saves
 *                     every layer input as bf16 slabs (stash: n_tiles * stash_slabs_per_tile *
 *                     16 KiB) and the ReLU sign bits (masks: n_tiles * mask_words_per_tile * 128
 *                     uint32).
 * nerfb200_mlp_bwd    data gradients: walks the layers in reverse, writes every dY as bf16 slabs
 *                     (dy_stash: n_tiles * program.stash_slabs_per_tile * 16 KiB) and, if
 *                     pos_grad_cols/dir_grad_cols are non-zero, accumulates (+=, caller zeroes)
 *                     the gradients w.r.t. ray origins / directions or per-sample positions /
 *                     directions.
 * nerfb200_mlp_wgrad  weight gradients dW += dY^T X and bias gradients db += column sums of dY
 *                     over the work items (NbWgradItem, DEVICE array) into d_params.
 */
int nerfb200_mlp_pack(const float* params, const NbPackChunk* chunks_dev, int n_chunks,
                      void* wpack, const NbPackBias* biases_dev, int n_biases,
                      float* bias_out, void* stream);
int nerfb200_mlp_fwd(const void* program_host, const void* wpack, const float* bias,
                     const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                     const NbPeCfg* pe_dir_host, const float* alpha_pos,
                     const float* alpha_dir, float sigma_bias, float* out_sigma, float* out_rgb,
                     void* stash, uint32_t* masks, int n_bias_floats, void* stream);
/* The same forward with two 128-sample tiles in flight per CTA (the epilogue of one runs under the
 * MMAs of the other). Takes the forward weights packed as per-K-step images (NbPackChunk.img_rows
 * > 0) and the float offset of the density weight row inside the packed biases (-1: the network
 * has no "extra" density column). Programs it cannot run (6-slab shape, density row riding in a
 * main image) are rejected with NERFB200_ERR_ARG: use nerfb200_mlp_fwd. Same stash / mask layout. */
int nerfb200_mlp_fwd2(const void* program, const void* wpack_k16, const float* bias,
                      const NbMlpInputs* in, const NbPeCfg* pe_pos, const NbPeCfg* pe_dir,
                      const float* alpha_pos, const float* alpha_dir, float sigma_bias,
                      float* out_sigma, float* out_rgb, void* stash, uint32_t* masks,
                      int n_bias_floats, int density_w_off, void* stream);

int nerfb200_mlp_bwd(const void* program_host, const void* wpack_t,
                     const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                     const NbPeCfg* pe_dir_host, const float* alpha_pos,
                     const float* alpha_dir, const float* sigma, const float* rgb,
                     const float* g_sigma, const float* g_rgb, const uint32_t* masks,
                     int fwd_mask_words_per_tile, void* dy_stash, int head_sigma_col3,
                     int pos_grad_cols, int dir_grad_cols, float* d_ray_o, float* d_ray_d,
                     float* d_pos, float* d_dir, void* stream);
int nerfb200_mlp_wgrad(const NbWgradItem* items_dev, int n_items, const void* x_stash,
                       int x_slabs_per_tile, const void* dy_stash, int dy_slabs_per_tile,
                       const void* z_stash, int z_slabs_per_tile, const float* params,
                       float* d_params, void* stream);
/* Sizes of the caller-owned workspaces of nerfb200_mlp_fwd (training) for n_samples samples:
 * the activation stash and the ReLU sign-bit masks. (The dY stash of nerfb200_mlp_bwd is
 * ceil(n / 128) * backward_program.stash_slabs_per_tile * 16384 bytes.) */
int nerfb200_mlp_workspace_bytes(const void* program_host, long long n_samples,
                                 long long* stash_bytes, long long* mask_bytes);

/* ------------------------------------------------------------------------------------------
 * The inner fused op of BASELINE.json's north_star — forward(rays_o, rays_d, near, far) -> rgb, depth,
 * weights — as one call: uniform t-sampling (barf/model_interpolation.py:135-180), fused field (:288-312,
 * barf/model_interpolation_architecture.py:96-141) and compositing (:316-353) enqueued back to back.
 * Inference only (no activation stash). rays_o, rays_d (B,3); pixel_width (B) or NULL; t_start / t_end
 * (B,S): the sample bins, or both NULL for uniform sampling with jitter (B,S) / offset_u (B) or NULL as
 * in nerfb200_sample_uniform; t_mode 0 = "left", 1 = "middle"; flavour NERFB200_COMPOSITE_*.
 * out_rgb (B,3); out_depth (B) = sum w t_mid / max(sum w, eps), out_weights (B,S), out_opacity (B)
 * = sum w: each of the three may be NULL. workspace:
 * nerfb200_render_rays_workspace_bytes(B, S) bytes, caller-owned.
 */
int nerfb200_render_rays_workspace_bytes(int B, int S, long long* bytes);
int nerfb200_render_rays(const void* program_host, const void* wpack, const float* bias,
                         int n_bias_floats, const NbPeCfg* pe_pos_host, const NbPeCfg* pe_dir_host,
                         const float* alpha_pos, const float* alpha_dir, float sigma_bias,
                         const float* rays_o, const float* rays_d, const float* pixel_width, int B, int S,
                         float near_t, float far_t, const float* t_start, const float* t_end,
                         const float* jitter, const float* offset_u,
                         float offset_size, int t_mode, int flavour, void* workspace, float* out_rgb,
                         float* out_depth, float* out_weights, float* out_opacity, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7, a8 (+ a6 inside). Fused GARF field: RadianceNetwork.forward (garf/model_radiance.py:84-96,
 * == barf/model_garf_radiance.py:101-113) and ProposalNetwork.forward (garf/model_proposal.py:55-56)
 * with the Gaussian activation (barf/gaussian.py:10-31) and their autograd, on per-ray or
 * per-sample inputs (NbMlpInputs; positions o + t d are formed in registers, replacing
 * GarfModel._get_positions / repeat_interleave, garf/model_garf.py:105,141). The tile programs
 * (include/nerfb200_garf.h) come from the host mirror (nerf_experiments_b200/garf_program.py).
 *   wpack / floats: packed bf16 weight images and fp32 values written by nerfb200_mlp_pack;
 *   params: the flat fp32 master parameters (the first layer is evaluated from them in fp32);
 *   fwd: out_sigma (N) [, out_rgb (N,3)]; y_stash / z_stash: NULL for inference, else workspaces of
 *        nerfb200_garf_workspace_bytes() receiving the activations / pre-activations (bf16 slabs);
 *   bwd: data gradients; dy_stash (backward_program.y_slabs_per_tile slabs per tile) receives dz of
 *        every layer; parameter gradients come from nerfb200_mlp_wgrad over (y_stash, dy_stash,
 *        z_stash): weight blocks as MMAs, bias and Gaussian-width gradients as column sums;
 *        want_input_grads: also d(ray origin / direction) or d(position / direction) (+=).
 */
int nerfb200_garf_workspace_bytes(const void* program_host, long long n_samples,
                                  long long* y_stash_bytes, long long* z_stash_bytes);
int nerfb200_garf_fwd(const void* program_host, const void* wpack, const float* floats,
                      const float* params, const NbMlpInputs* in_host, float* out_sigma,
                      float* out_rgb, void* y_stash, void* z_stash, void* stream);
int nerfb200_garf_bwd(const void* program_host, const void* wpack_t, const float* floats,
                      const float* params, const NbMlpInputs* in_host, const float* sigma,
                      const float* rgb, const float* g_sigma, const float* g_rgb,
                      const void* z_stash, void* dy_stash, int want_input_grads, float* d_ray_o,
                      float* d_ray_d, float* d_pos, float* d_dir, void* stream);

/* ------------------------------------------------------------------------------------------
 * a6, a9. Learnable per-feature activations of the GARF / SARF / Gabor networks over (N, F)
 * row-major fp32 activations. Replaces GaussActivation.forward/backward (barf/gaussian.py:8-34,
 * == garf/gaussian.py), SarfAct.forward and its autograd (sarf/activation.py:63-65) and
 * GaborActivation.forward/backward (gaborf/gabor.py:8-29).
 *   p0: (F,) inverse standard deviation (GAUSS, GABOR: v = p0^2 + 1e-6) or frequency (SARF)
 *   p1: (F,) spread (GABOR only, else NULL)
 *   bwd: dx (N, F) is written; dp0 / dp1 (F,) accumulate (+=, caller zeroes) the gradients
 *        w.r.t. the PARAMETERS p0 / p1 (the v = p0^2 + 1e-6 chain rule is applied inside);
 *        dsum (F,) or NULL accumulates the column sums of dx, i.e. the bias gradient of the
 *        Linear layer that produced x (saves the separate reduction autograd would launch).
 *   out_bf16 != 0: y / dx are written as bf16 (N, F) — the operand type of the GEMM that
 *        consumes them — instead of fp32; inputs, sums and parameter gradients stay fp32.
 */
enum { NERFB200_ACT_GAUSS = 0, NERFB200_ACT_SARF = 1, NERFB200_ACT_GABOR = 2 };
int nerfb200_act_fwd(int kind, const float* x, const float* p0, const float* p1, long long N,
                     int F, void* y, int out_bf16, void* stream);
int nerfb200_act_bwd(int kind, const float* x, const float* p0, const float* p1, const float* g,
                     long long N, int F, void* dx, float* dp0, float* dp1, float* dsum,
                     int out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------
 * GPU-resident ray batcher. Replaces
 * ImagePoseDataset.__getitem__ + DataLoader collate + H2D copy (barf/dataset.py:613-637; ray
 * generation :407-482) and ImagePoseDataModule.get_blurred_pixel_colors
 * (barf/data_module.py:276-369).
 *   ray_index (B) int64: flat index image * H*W + pixel;  c2w_raw / c2w_noisy (N,4,4) row-major;
 *   images (N, H*W, n_sigmas, 3) fp32;  image_id_map (N) int32 or NULL (identity).
 *   blur_low < 0: colors (B, n_sigmas, 3) = the stored pyramid;
 *   otherwise colors (B,2,3): [:,0] = c[blur_low]*blur_coef + c[blur_high]*(1-blur_coef),
 *                             [:,1] = c[n_sigmas-1] (the unblurred pixel).
 *   Outputs: o_raw, o_noisy, d_raw, d_noisy (B,3) fp32, img_idx (B) int64, pixel_width_out (B).
 */
int nerfb200_ray_batch(const long long* ray_index, int B, const float* c2w_raw,
                       const float* c2w_noisy, const float* images, const int* image_id_map,
                       int n_images, int H, int W, int n_sigmas, float focal, float pixel_width,
                       int blur_low, int blur_high, float blur_coef, float* o_raw, float* o_noisy,
                       float* d_raw, float* d_noisy, float* colors, long long* img_idx,
                       float* pixel_width_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Pose alignment on the device. Replaces
 * CameraCalibrationModel.kabsch_algorithm and compute_pose_error
 * (barf/model_camera_calibration.py:69-156, :340-346).
 *   from, to (n,3) fp32, n <= 2048.  R (3,3), t (3), c (1):  to ~= c * R from + t.
 *   remove_outliers: refit on the points closer than the 0.9 distance quantile of a first fit.
 *   out_err (1) or NULL: mean |c R from + t - to| over all points with the final fit.
 */
int nerfb200_kabsch(const float* from, const float* to, int n, int remove_outliers, float* out_R,
                    float* out_t, float* out_c, float* out_err, void* stream);

/* ------------------------------------------------------------------------------------------
 * K9. Fused Adam over the flat fp32 parameter buffer (torch.optim.Adam arithmetic, per-group
 * learning rate / weight decay), replacing the optimiser the reference configures at
 * barf/model_interpolation.py:543-564.  group_* are HOST arrays of n_groups entries
 * ([begin,end) float ranges of the flat buffer); step counts from 1; grads are multiplied by
 * grad_scale first (1/world_size after a sum all-reduce).
 */
#define NERFB200_MAX_ADAM_GROUPS 8
int nerfb200_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                       long long n, int n_groups, const long long* group_begin_host,
                       const long long* group_end_host, const float* group_lr_host,
                       const float* group_wd_host, float beta1, float beta2, float eps,
                       long long step, float grad_scale, void* stream);
/* The whole optimiser step driven from DEVICE memory, so that a launch captured in a CUDA graph
 * needs no host work per step: `state` = 3 int64 (zero-initialised by the caller): [0] steps taken,
 * [1] steps skipped, [2] scratch. The kernel evaluates every group's learning-rate schedule in
 * closed form for step state[0] + 1 — SchedulerLeNice (barf/model_interpolation.py:30-67) or
 * ExponentialLR, lr0 * gamma^(step - 1) (garf/model_garf.py:365-428) — and Adam's
 * bias corrections for state[0] + 1 - state[1], then advances the counters. When `loss_flag` (one
 * float, may be NULL) is NaN or infinite the step is a no-op for parameters and moments, as in the
 * reference, whose NaN loss is replaced by a fresh leaf so that no parameter receives a gradient
 * (barf/model_interpolation.py:522-524); the skipped step still advances the schedules. */
enum { NERFB200_LR_LE_NICE = 0, NERFB200_LR_EXPONENTIAL = 1 };
typedef struct {
  long long begin, end;      /* [begin, end) floats of the flat buffer                         */
  long long n_steps;         /* LE_NICE: lr0 * exp(log_factor * min(step, n_steps)); EXPONENTIAL: unused */
  float lr0;                 /* learning_rate_start                                             */
  float log_factor;          /* LE_NICE: (ln stop - ln start) / n; EXPONENTIAL: ln gamma        */
  float weight_decay;
  int32_t mode;              /* NERFB200_LR_*                                                   */
} NbAdamGroup;
int nerfb200_adam_step_dev(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                           long long n, const NbAdamGroup* groups_host, int n_groups, float beta1,
                           float beta2, float eps, float grad_scale, const float* loss_flag,
                           long long* state, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_H_ */
