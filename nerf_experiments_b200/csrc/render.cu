// The inner fused op of the hot path as ONE entry point (BASELINE.json north_star:
// "forward(rays_o, rays_d, near, far) returns rgb, depth and weights"): uniform t-sampling
// (reference barf/model_interpolation.py:135-180), the fused field (positions + encodings + NerfModel,
// :288-312 and barf/model_interpolation_architecture.py:96-141) and alpha compositing (:316-353) with
// the expected depth sum w t_mid and the opacity sum w, enqueued back to back on the caller's stream
// over a caller-owned workspace. Inference only (no stash); the training path keeps the separate
// entry points because autograd sits between them.
#include "common.cuh"
#include "mlp.h"

extern "C" int nerfb200_sample_uniform(double, double, int, int, const float*, const float*, double, float*, float*, void*);
extern "C" int nerfb200_mlp_fwd(const void*, const void*, const float*, const NbMlpInputs*, const NbPeCfg*, const NbPeCfg*,
                                const float*, const float*, float, float*, float*, void*, uint32_t*, int, void*);
extern "C" int nerfb200_composite_fwd(const float*, const float*, const float*, const float*, int, int, int, float*, float*,
                                      float*, float*, void*);

namespace nerfb200 {
namespace {

__global__ void __launch_bounds__(256)
interval_kernel(const float* __restrict__ t0, const float* __restrict__ t1, long long n,
                float* __restrict__ delta, float* __restrict__ t_mid) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = t0[i], b = t1[i];
    delta[i] = b - a;
    t_mid[i] = (a + b) * 0.5f;
  }
}

// workspace layout (floats): t_start | t_end | delta | t_mid | sigma | rgb (3x)
inline long long ws_floats(long long B, long long S) { return B * S * 8; }

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_render_rays_workspace_bytes(int B, int S, long long* bytes) {
  NB_CHECK_ARG(B >= 0 && S >= 1 && bytes, "render_rays_workspace_bytes: bad arguments");
  *bytes = ws_floats(B, S) * 4;
  return NERFB200_OK;
}

extern "C" int nerfb200_render_rays(const void* program_host, const void* wpack, const float* bias,
                                    int n_bias_floats, const NbPeCfg* pe_pos_host, const NbPeCfg* pe_dir_host,
                                    const float* alpha_pos, const float* alpha_dir, float sigma_bias,
                                    const float* rays_o, const float* rays_d, const float* pixel_width, int B,
                                    int S, float near_t, float far_t, const float* t_start_in,
                                    const float* t_end_in, const float* jitter, const float* offset_u,
                                    float offset_size, int t_mode, int flavour, void* workspace, float* out_rgb,
                                    float* out_depth, float* out_weights, float* out_opacity, void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "render_rays: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(program_host && wpack && bias && pe_pos_host && pe_dir_host && rays_o && rays_d && workspace && out_rgb,
               "render_rays: null pointer");
  NB_CHECK_ARG(far_t > near_t, "render_rays: far must exceed near");
  NB_CHECK_ARG((t_start_in == nullptr) == (t_end_in == nullptr), "render_rays: t_start and t_end go together");
  NB_CHECK_ARG(t_mode == 0 || t_mode == 1, "render_rays: t_mode must be 0 (left) or 1 (middle)");
  if (B == 0) return NERFB200_OK;
  const long long n = (long long)B * S;
  float* ws = reinterpret_cast<float*>(workspace);
  const float* t0 = t_start_in ? t_start_in : ws;
  const float* t1 = t_end_in ? t_end_in : ws + n;
  float* delta = ws + 2 * n;
  float* t_mid = ws + 3 * n;
  float* sigma = ws + 4 * n;
  float* rgb = ws + 5 * n;
  int rc = NERFB200_OK;
  if (t_start_in == nullptr) {
    rc = nerfb200_sample_uniform(near_t, far_t, B, S, jitter, offset_u, offset_size, ws, ws + n, stream);
    if (rc != NERFB200_OK) return rc;
  }
  int blocks = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  interval_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(t0, t1, n, delta, t_mid);
  count_launch();
  NB_CHECK_LAUNCH();
  NbMlpInputs in;
  memset(&in, 0, sizeof(in));
  in.N = n;
  in.S = S;
  in.t_mode = t_mode;
  in.ray_o = rays_o;
  in.ray_d = rays_d;
  in.t_start = t0;
  in.t_end = t1;
  in.pixel_width = pixel_width;
  in.pixel_width_per_sample = 0;
  rc = nerfb200_mlp_fwd(program_host, wpack, bias, &in, pe_pos_host, pe_dir_host, alpha_pos, alpha_dir, sigma_bias,
                        sigma, rgb, nullptr, nullptr, n_bias_floats, stream);
  if (rc != NERFB200_OK) return rc;
  return nerfb200_composite_fwd(sigma, delta, rgb, t_mid, B, S, flavour, out_rgb, out_weights, out_opacity, out_depth,
                                stream);
}
