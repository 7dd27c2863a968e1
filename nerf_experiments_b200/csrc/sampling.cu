// Uniform t-sampling (a1) and the equidistant fallback of the pdf sampler.
// Mirrors NerfInterpolation._sample_t_stratified_uniform + _get_intervals
// (reference barf/model_interpolation.py:135-180, :114-132).
#include "common.cuh"

namespace nerfb200 {

// th.linspace(start, end, steps)[k] in fp32, as ATen computes it: step=(end-start)/(steps-1),
// the lower half counted up from start and the upper half counted down from end.
__device__ __forceinline__ float linspace_at(float start, float end, float step, int steps,
                                             int k) {
  const int halfway = steps / 2;
  if (k < halfway) return __fadd_rn(start, __fmul_rn(step, (float)k));
  return __fsub_rn(end, __fmul_rn(step, (float)(steps - k - 1)));
}

struct UniformArgs {
  float start, end, step, delta, far_t, offset_size;
};

__device__ __forceinline__ float uniform_t(const UniformArgs& a, int S, int k, float jit,
                                           bool has_jit, float off, bool has_off) {
  float t = linspace_at(a.start, a.end, a.step, S, k);
  if (has_jit) t = __fadd_rn(t, __fmul_rn(jit, a.delta));
  if (has_off) t = __fadd_rn(t, __fmul_rn(__fmul_rn(off, a.delta), a.offset_size));
  return t;
}

namespace {

__global__ void __launch_bounds__(256)
sample_uniform_kernel(UniformArgs a, int B, int S, const float* __restrict__ jitter,
                      const float* __restrict__ offset_u, const int32_t* __restrict__ gate,
                      float* __restrict__ t_start, float* __restrict__ t_end) {
  if (gate != nullptr && *gate == 0) return;  // fallback variant: only when the flag is set
  const long long n = (long long)B * S;
  const bool has_off = (offset_u != nullptr) && (a.offset_size != 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / S);
    const int k = (int)(i - (long long)r * S);
    const float off = has_off ? __ldg(offset_u + r) : 0.f;
    const float t0 = uniform_t(a, S, k, jitter ? jitter[i] : 0.f, jitter != nullptr, off, has_off);
    float t1 = a.far_t;
    if (k + 1 < S)
      t1 = uniform_t(a, S, k + 1, jitter ? jitter[i + 1] : 0.f, jitter != nullptr, off, has_off);
    t_start[i] = t0;
    t_end[i] = t1;
  }
}

int launch_uniform(double near_t, double far_t, int B, int S, const float* jitter,
                   const float* offset_u, double offset_size, const int32_t* gate,
                   float* t_start, float* t_end, void* stream) {
  // Host arithmetic follows the reference's Python-float (double) expressions
  // (barf/model_interpolation.py:159-166) before ATen narrows them to fp32.
  const double interval = (far_t - near_t) / (double)S;
  UniformArgs a;
  a.start = (float)near_t;
  a.end = (float)(far_t - interval);
  a.step = (S > 1) ? (a.end - a.start) / (float)(S - 1) : 0.f;
  a.delta = (float)interval;
  a.far_t = (float)far_t;
  a.offset_size = (float)offset_size;
  const long long n = (long long)B * S;
  int blocks = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sample_uniform_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, B, S, jitter, offset_u, gate,
                                                                  t_start, t_end);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_sample_uniform(double near_t, double far_t, int B, int S,
                                       const float* jitter, const float* offset_u,
                                       double offset_size, float* t_start, float* t_end,
                                       void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "sample_uniform: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(t_start && t_end, "sample_uniform: null output");
  if (B == 0) return NERFB200_OK;
  return launch_uniform(near_t, far_t, B, S, jitter, offset_u, offset_size, nullptr, t_start,
                        t_end, stream);
}

extern "C" int nerfb200_resample_fallback(const int32_t* fail_flag, double near_t, double far_t,
                                          int B, int Sf, const float* offset_u, float* t_start,
                                          float* t_end, void* stream) {
  NB_CHECK_ARG(B >= 0 && Sf >= 1, "resample_fallback: bad shape B=%d Sf=%d", B, Sf);
  NB_CHECK_ARG(fail_flag && t_start && t_end, "resample_fallback: null pointer");
  if (B == 0) return NERFB200_OK;
  // reference: _sample_t_stratified_uniform(batch, n_samples, "equidistant", -1)
  // (barf/model_interpolation.py:275)
  return launch_uniform(near_t, far_t, B, Sf, nullptr, offset_u, -1.0, fail_flag, t_start, t_end,
                        stream);
}
