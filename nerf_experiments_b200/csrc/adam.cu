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

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, AdamGroups groups, float beta1, float beta2,
            float eps, float bias_c1, float bias_c2_sqrt, float grad_scale) {
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
                                                        beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

// ---------------------------------------------------------------------------------------------
// The whole optimiser step driven from DEVICE memory: step counter, learning-rate schedules in closed
// form, bias corrections and the non-finite-loss guard. Nothing of a step is a launch parameter, so a
// launch captured in a CUDA graph follows the schedules across replays with no host work at all.
// ---------------------------------------------------------------------------------------------
namespace nerfb200 {
namespace {

struct AdamDevGroups {
  int n_groups;
  NbAdamGroup g[NERFB200_MAX_ADAM_GROUPS];
};

// state: [0] optimiser steps launched so far, [1] steps skipped by the guard, [2] block ticket
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, long long n, AdamDevGroups groups, float beta1, float beta2,
                float eps, float grad_scale, const float* __restrict__ loss_flag,
                long long* __restrict__ state) {
  __shared__ float s_lr[NERFB200_MAX_ADAM_GROUPS];
  __shared__ float s_bc[2];
  __shared__ int s_skip;
  const long long step = state[0] + 1;              // 1-based index of this optimiser step
  if (threadIdx.x == 0) {
    // reference: a NaN loss is replaced by a fresh leaf, so no parameter receives a gradient and
    // Adam skips every one of them (barf/model_interpolation.py:522-524, garf/model_garf.py:283-289)
    const float flag = loss_flag != nullptr ? *loss_flag : 0.f;
    s_skip = !isfinite(flag);
    const long long eff = step - state[1];          // per-parameter `step` of torch.optim.Adam
    s_bc[0] = (float)(1.0 - pow((double)beta1, (double)eff));
    s_bc[1] = (float)sqrt(1.0 - pow((double)beta2, (double)eff));
  }
  if (threadIdx.x < NERFB200_MAX_ADAM_GROUPS && threadIdx.x < groups.n_groups) {
    const NbAdamGroup& q = groups.g[threadIdx.x];
    double e = 0.0;
    if (q.mode == NERFB200_LR_LE_NICE) {
      // SchedulerLeNice (barf/model_interpolation.py:43-48,65-67): lr0 * exp(logf * min(step, n))
      // (log_factor is 0 for the constant cases; n = -1, CameraExtrinsics' default, gives lr = stop)
      e = (double)q.log_factor * (double)(step < q.n_steps ? step : q.n_steps);
    } else if (q.mode == NERFB200_LR_EXPONENTIAL) {
      // torch ExponentialLR (garf/model_garf.py:365-428): one factor gamma per scheduler step
      e = (double)q.log_factor * (double)(step - 1);
    }
    s_lr[threadIdx.x] = (float)((double)q.lr0 * exp(e));
  }
  __syncthreads();
  if (!s_skip) {
    const float bias_c1 = s_bc[0], bias_c2_sqrt = s_bc[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
      float lr = 0.f, wd = 0.f;
      bool found = false;
#pragma unroll
      for (int q = 0; q < NERFB200_MAX_ADAM_GROUPS; ++q) {
        if (q < groups.n_groups && i >= groups.g[q].begin && i < groups.g[q].end) {
          lr = s_lr[q];
          wd = groups.g[q].weight_decay;
          found = true;
        }
      }
      if (!found) continue;
      const float w = p[i];
      float grad = g[i] * grad_scale;
      if (wd != 0.f) grad = grad + wd * w;
      const float mi = m[i] + (grad - m[i]) * (1.f - beta1);
      const float vi = v[i] * beta2 + (1.f - beta2) * grad * grad;
      m[i] = mi;
      v[i] = vi;
      const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
      p[i] = w - (lr / bias_c1) * (mi / denom);
    }
  }
  // the last block to finish advances the counters (every block has read them by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(state + 2), 1ull);
    if (ticket == (unsigned long long)gridDim.x - 1ull) {
      state[2] = 0;
      state[1] += s_skip ? 1 : 0;
      state[0] = step;
      __threadfence();
    }
  }
}

}  // namespace
}  // namespace nerfb200

extern "C" int nerfb200_adam_step_dev(float* params, const float* grads, float* exp_avg,
                                      float* exp_avg_sq, long long n, const NbAdamGroup* groups_host,
                                      int n_groups, float beta1, float beta2, float eps,
                                      float grad_scale, const float* loss_flag, long long* state,
                                      void* stream) {
  NB_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && groups_host && state && n >= 0,
               "adam_step_dev: null pointer");
  NB_CHECK_ARG(n_groups >= 1 && n_groups <= NERFB200_MAX_ADAM_GROUPS, "adam_step_dev: n_groups=%d", n_groups);
  AdamDevGroups gr;
  memset(&gr, 0, sizeof(gr));
  gr.n_groups = n_groups;
  for (int q = 0; q < n_groups; ++q) {
    gr.g[q] = groups_host[q];
    NB_CHECK_ARG(gr.g[q].mode == NERFB200_LR_LE_NICE || gr.g[q].mode == NERFB200_LR_EXPONENTIAL,
                 "adam_step_dev: group %d: unknown schedule %d", q, gr.g[q].mode);
    NB_CHECK_ARG(gr.g[q].begin >= 0 && gr.g[q].end >= gr.g[q].begin && gr.g[q].end <= n,
                 "adam_step_dev: group %d: bad range", q);
  }
  int blocks = ceil_div(n > 0 ? n : 1, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adam_dev_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, gr, beta1,
                                                            beta2, eps, grad_scale, loss_flag, state);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
