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
 *   - workspaces are caller-owned and sized by the matching *_workspace_bytes() call.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef NERFB200_H_
#define NERFB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFB200_ABI_VERSION 1

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

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_H_ */
