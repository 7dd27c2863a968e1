// Stand-alone positional-encoding kernels (fp32 in / fp32 out), for callers that use the
// encoder modules on their own (reference barf/positional_encodings.py forward() of each class).
// The fused MLP kernels evaluate the same device functions (pe.cuh) in registers instead.
#include "common.cuh"
#include "mlp.h"
#include "pe.cuh"

namespace nerfb200 {
namespace {

__global__ void __launch_bounds__(128)
pe_fwd_kernel(NbPeCfg cfg, const float* __restrict__ alpha, const float* __restrict__ pos,
              const float* __restrict__ dir, const float* __restrict__ pixel_width,
              const float* __restrict__ t0, const float* __restrict__ t1, long long N, int out_dim,
              float* __restrict__ out) {
  __shared__ float mask[kMaxLevels];
  if (threadIdx.x == 0) pe_fill_mask(cfg, alpha, mask);
  __syncthreads();
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (long long)gridDim.x * blockDim.x) {
    PeSample s;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.x[c] = pos[n * 3 + c];
      s.dir[c] = dir ? dir[n * 3 + c] : 0.f;
    }
    s.pixel_width = pixel_width ? pixel_width[n] : 0.f;
    s.t0 = t0 ? t0[n] : 0.f;
    s.t1 = t1 ? t1[n] : 0.f;
    float* o = out + n * out_dim;
    pe_encode(cfg, mask, s, [&](int col, float v) { o[col] = v; });
  }
}

__global__ void __launch_bounds__(128)
pe_bwd_kernel(NbPeCfg cfg, const float* __restrict__ alpha, const float* __restrict__ pos,
              const float* __restrict__ dir, const float* __restrict__ pixel_width,
              const float* __restrict__ t0, const float* __restrict__ t1,
              const float* __restrict__ g_out, long long N, int out_dim, float* __restrict__ d_pos,
              float* __restrict__ d_dir) {
  __shared__ float mask[kMaxLevels];
  if (threadIdx.x == 0) pe_fill_mask(cfg, alpha, mask);
  __syncthreads();
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N;
       n += (long long)gridDim.x * blockDim.x) {
    PeSample s;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.x[c] = pos[n * 3 + c];
      s.dir[c] = dir ? dir[n * 3 + c] : 0.f;
    }
    s.pixel_width = pixel_width ? pixel_width[n] : 0.f;
    s.t0 = t0 ? t0[n] : 0.f;
    s.t1 = t1 ? t1[n] : 0.f;
    const float* g = g_out + n * out_dim;
    float dx[3], ds;
    pe_backward(cfg, mask, s, [&](int col) { return g[col]; }, dx, ds);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      d_pos[n * 3 + c] = dx[c];
      if (d_dir) d_dir[n * 3 + c] = ds * dx[c];
    }
  }
}

int pe_out_dim(const NbPeCfg& c) {
  if (c.kind == NB_PE_IDENTITY) return 3;
  return (c.levels * 2 + (c.include_identity ? 1 : 0)) * 3;
}

int grid_for(long long N) {
  long long b = (N + 127) / 128;
  const int cap = sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_pe_fwd(const NbPeCfg* cfg_host, const float* alpha, const float* pos,
                               const float* dir, const float* pixel_width, const float* t_start,
                               const float* t_end, long long N, float* out, void* stream) {
  NB_CHECK_ARG(cfg_host && pos && out && N >= 0, "pe_fwd: bad arguments");
  NB_CHECK_ARG(cfg_host->levels >= 0 && cfg_host->levels <= kMaxLevels, "pe_fwd: levels=%d unsupported", cfg_host->levels);
  NB_CHECK_ARG(cfg_host->kind != NB_PE_INTEGRATED || (dir && pixel_width && t_start && t_end),
               "pe_fwd: the integrated encoding needs dir, pixel_width, t_start, t_end");
  if (N == 0) return NERFB200_OK;
  pe_fwd_kernel<<<grid_for(N), 128, 0, (cudaStream_t)stream>>>(*cfg_host, alpha, pos, dir, pixel_width,
                                                               t_start, t_end, N, pe_out_dim(*cfg_host), out);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_pe_bwd(const NbPeCfg* cfg_host, const float* alpha, const float* pos,
                               const float* dir, const float* pixel_width, const float* t_start,
                               const float* t_end, const float* g_out, long long N, float* d_pos,
                               float* d_dir, void* stream) {
  NB_CHECK_ARG(cfg_host && pos && g_out && d_pos && N >= 0, "pe_bwd: bad arguments");
  NB_CHECK_ARG(cfg_host->levels >= 0 && cfg_host->levels <= kMaxLevels, "pe_bwd: levels=%d unsupported", cfg_host->levels);
  NB_CHECK_ARG(cfg_host->kind != NB_PE_INTEGRATED || (dir && pixel_width && t_start && t_end),
               "pe_bwd: the integrated encoding needs dir, pixel_width, t_start, t_end");
  if (N == 0) return NERFB200_OK;
  pe_bwd_kernel<<<grid_for(N), 128, 0, (cudaStream_t)stream>>>(*cfg_host, alpha, pos, dir, pixel_width,
                                                               t_start, t_end, g_out, N,
                                                               pe_out_dim(*cfg_host), d_pos, d_dir);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
