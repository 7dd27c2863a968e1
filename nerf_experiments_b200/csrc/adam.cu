// Fused multi-group Adam on the flat fp32 parameter buffer (torch.optim.Adam arithmetic, the
// optimiser the reference configures: Adam(eps=1e-5) with per-group learning rates,
// barf/model_interpolation.py:543-564). One launch replaces ~60 foreach kernels; the per-group
// learning rate is the closed form of SchedulerLeNice (:65-67) evaluated by the caller.
#include "common.cuh"

namespace nerfb200 {
namespace {

struct AdamGroups {
  int n_groups;
  long long begin[NERFB200_MAX_ADAM_GROUPS];
  long long end[NERFB200_MAX_ADAM_GROUPS];
  float lr[NERFB200_MAX_ADAM_GROUPS];
  float weight_decay[NERFB200_MAX_ADAM_GROUPS];
};

// sched (optional, DEVICE): [bias_c1, sqrt(bias_c2), lr[0..n_groups)] of this step — read from memory
// instead of the launch parameters, so that a launch captured in a CUDA graph follows the schedule
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, AdamGroups groups, float beta1, float beta2,
            float eps, float bias_c1, float bias_c2_sqrt, float grad_scale,
            const float* __restrict__ sched) {
  if (sched != nullptr) {
    bias_c1 = sched[0];
    bias_c2_sqrt = sched[1];
#pragma unroll
    for (int q = 0; q < NERFB200_MAX_ADAM_GROUPS; ++q)
      if (q < groups.n_groups) groups.lr[q] = sched[2 + q];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float lr = 0.f, wd = 0.f;
    bool found = false;
#pragma unroll
    for (int q = 0; q < NERFB200_MAX_ADAM_GROUPS; ++q) {
      if (q < groups.n_groups && i >= groups.begin[q] && i < groups.end[q]) {
        lr = groups.lr[q];
        wd = groups.weight_decay[q];
        found = true;
      }
    }
    if (!found) continue;
    const float w = p[i];
    float grad = g[i] * grad_scale;
    if (wd != 0.f) grad = grad + wd * w;
    const float mi = m[i] + (grad - m[i]) * (1.f - beta1);         // exp_avg.lerp_(grad, 1-beta1)
    const float vi = v[i] * beta2 + (1.f - beta2) * grad * grad;   // mul_(beta2).addcmul_(g, g, 1-beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
    p[i] = w - (lr / bias_c1) * (mi / denom);
  }
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_adam_step(float* params, const float* grads, float* exp_avg,
                                  float* exp_avg_sq, long long n, int n_groups,
                                  const long long* group_begin_host, const long long* group_end_host,
                                  const float* group_lr_host, const float* group_wd_host,
                                  float beta1, float beta2, float eps, long long step,
                                  float grad_scale, void* stream) {
  NB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 0, "adam_step: null pointer");
  NB_CHECK_ARG(n_groups >= 1 && n_groups <= NERFB200_MAX_ADAM_GROUPS, "adam_step: n_groups=%d", n_groups);
  NB_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
  if (n == 0) return NERFB200_OK;
  AdamGroups gr;
  gr.n_groups = n_groups;
  for (int q = 0; q < NERFB200_MAX_ADAM_GROUPS; ++q) {
    gr.begin[q] = q < n_groups ? group_begin_host[q] : 0;
    gr.end[q] = q < n_groups ? group_end_host[q] : 0;
    gr.lr[q] = q < n_groups ? group_lr_host[q] : 0.f;
    gr.weight_decay[q] = q < n_groups ? group_wd_host[q] : 0.f;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  int blocks = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, gr, beta1,
                                                        beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale, nullptr);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_adam_step_sched(float* params, const float* grads, float* exp_avg,
                                        float* exp_avg_sq, long long n, int n_groups,
                                        const long long* group_begin_host, const long long* group_end_host,
                                        const float* group_wd_host, const float* sched_dev, float beta1,
                                        float beta2, float eps, float grad_scale, void* stream) {
  NB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && sched_dev && n >= 0, "adam_step_sched: null pointer");
  NB_CHECK_ARG(n_groups >= 1 && n_groups <= NERFB200_MAX_ADAM_GROUPS, "adam_step_sched: n_groups=%d", n_groups);
  if (n == 0) return NERFB200_OK;
  AdamGroups gr;
  gr.n_groups = n_groups;
  for (int q = 0; q < NERFB200_MAX_ADAM_GROUPS; ++q) {
    gr.begin[q] = q < n_groups ? group_begin_host[q] : 0;
    gr.end[q] = q < n_groups ? group_end_host[q] : 0;
    gr.lr[q] = 0.f;
    gr.weight_decay[q] = q < n_groups ? group_wd_host[q] : 0.f;
  }
  int blocks = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, gr, beta1,
                                                        beta2, eps, 1.f, 1.f, grad_scale, sched_dev);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
