// Proposal resampling: the deterministic pdf allocator (a11) and inverse-CDF importance
// sampling (a12). One warp per ray, per-ray tables in shared memory, coalesced row I/O.
//
// a11 mirrors NerfInterpolation._sample_t_pdf_weighted (reference
// barf/model_interpolation.py:193-277): the reference's 64-iteration masked Python loop
// (~640 launches, O(B*Sc*Sf) work) becomes one launch with O(Sc^2 + Sf log Sc) work per ray.
// a12 restates nerfacc's importance_sampling (not in the reference tree; call sites
// garf/model_garf.py:210-220) — see oracle/ref_nerfacc.py for the algorithm statement.
#include "common.cuh"

namespace nerfb200 {
namespace {

constexpr int kWarps = 4;

// Sum of a row in the order the oracle defines (oracle/ref_resample.py: lane_strided_sum):
// lane l adds elements l, l+32, ... left to right, then a xor-butterfly over the lanes.
__device__ __forceinline__ float lane_strided_sum(const float* row, int n, int lane) {
  float acc = 0.f;
  for (int i = lane; i < n; i += 32) acc = __fadd_rn(acc, row[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
  return acc;
}

__global__ void __launch_bounds__(kWarps * 32)
resample_alloc_kernel(const float* __restrict__ t_coarse, const float* __restrict__ weights,
                      const float* __restrict__ delta_coarse, int B, int Sc, int Sf, float far_t,
                      float* __restrict__ t_start, float* __restrict__ t_end,
                      int32_t* __restrict__ counts_out, int32_t* __restrict__ fail_flag) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_warp = 4 * Sc + 1 + Sf;
  float* s_err = smem + (size_t)warp * per_warp;  // [Sc]   remainder e_i, later step d_i/n_i
  float* s_n = s_err + Sc;                        // [Sc]   n_i
  float* s_cum = s_n + Sc;                        // [Sc+1] exclusive cumsum of n
  float* s_tc = s_cum + Sc + 1;                   // [Sc]   coarse t
  float* s_t = s_tc + Sc;                         // [Sf]   fine t
  const float n_new = (float)(Sf - Sc);

  for (long long ray = (long long)blockIdx.x * kWarps + warp; ray < B;
       ray += (long long)gridDim.x * kWarps) {
    const float* w_row = weights + ray * Sc;
    const float wsum = lane_strided_sum(w_row, Sc, lane);
    float nsum = 0.f;
    for (int i = lane; i < Sc; i += 32) {
      const float p = __fdiv_rn(w_row[i], wsum);   // weights / weights.sum            (:215)
      const float raw = __fmul_rn(p, n_new);       // * (n_samples - n_bins)           (:216)
      const float fl = floorf(raw);                //                                   (:217)
      s_err[i] = __fsub_rn(raw, fl);               //                                   (:218)
      s_n[i] = fl;
      nsum += fl;  // integers: exact in any order
      s_tc[i] = t_coarse[ray * Sc + i];
    }
    nsum = warp_sum(nsum);
    const float excess = __fsub_rn(n_new, nsum);   // n_samples - n_bins - sum          (:224)
    const float thresh = __fsub_rn((float)Sc, excess);
    __syncwarp();
    // error_rank = argsort(argsort(err)) with ties to the lowest index              (:225)
    bool bad = false;
    for (int i = lane; i < Sc; i += 32) {
      const float e = s_err[i];
      int rank = 0;
      for (int j = 0; j < Sc; ++j) {
        const float ej = s_err[j];
        rank += (ej < e) || (ej == e && j < i);
      }
      const float add = ((float)rank >= thresh) ? 1.f : 0.f;                       // (:226)
      const float n = __fadd_rn(__fadd_rn(s_n[i], add), 1.f);                      // (:227)
      bad |= !(n >= 0.f);
      // all lanes have read s_err[*] only after the __syncwarp below
      s_cum[i + 1] = n;  // staged; turned into a cumsum next
    }
    __syncwarp();
    // inclusive cumsum of n into s_cum[1..Sc] (integers, exact), s_cum[0] = 0
    float carry = 0.f;
    for (int base = 0; base < Sc; base += 32) {
      const int i = base + lane;
      const float n = (i < Sc) ? s_cum[i + 1] : 0.f;
      const float incl = warp_inclusive_scan(n, lane);
      if (i < Sc) {
        s_n[i] = n;
        s_cum[i + 1] = carry + incl;
        s_err[i] = delta_coarse[ray * Sc + i];  // reuse: bin width
      }
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_cum[0] = 0.f;
    bad |= (carry != (float)Sf);                                                   // (:235)
    bad = __any_sync(0xffffffffu, bad);
    if (bad && lane == 0) atomicExch(fail_flag, 1);
    if (counts_out != nullptr)
      for (int i = lane; i < Sc; i += 32) counts_out[ray * Sc + i] = (int32_t)s_n[i];
    __syncwarp();
    // expand: t_k = t_c[i] + ((k - cum_i) * delta_i) / n_i   for cum_i <= k < cum_{i+1} (:262-269)
    for (int k = lane; k < Sf; k += 32) {
      const float kf = (float)k;
      int lo = 0, hi = Sc;  // largest i in [0,Sc) with cum[i] <= k
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_cum[mid] <= kf) lo = mid; else hi = mid;
      }
      const float num = __fmul_rn(__fsub_rn(kf, s_cum[lo]), s_err[lo]);
      s_t[k] = __fadd_rn(s_tc[lo], __fdiv_rn(num, s_n[lo]));
    }
    __syncwarp();
    for (int k = lane; k < Sf; k += 32) {
      t_start[ray * Sf + k] = s_t[k];
      t_end[ray * Sf + k] = (k + 1 < Sf) ? s_t[k + 1] : far_t;                     // (:127-130)
    }
    __syncwarp();
  }
}

// ---- a12: inverse-CDF importance sampling -------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
resample_icdf_kernel(const float* __restrict__ edges, const float* __restrict__ cdf,
                     const float* __restrict__ u_ray, int B, int Sc, int Sf,
                     float* __restrict__ out_edges, int32_t* __restrict__ out_idx) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_warp = 2 * (Sc + 1) + Sf;
  float* s_e = smem + (size_t)warp * per_warp;  // [Sc+1]
  float* s_c = s_e + Sc + 1;                    // [Sc+1]
  float* s_s = s_c + Sc + 1;                    // [Sf] sample centres
  for (long long ray = (long long)blockIdx.x * kWarps + warp; ray < B;
       ray += (long long)gridDim.x * kWarps) {
    for (int i = lane; i <= Sc; i += 32) {
      s_e[i] = edges[ray * (Sc + 1) + i];
      s_c[i] = cdf[ray * (Sc + 1) + i];
    }
    __syncwarp();
    const float u_floor = s_c[0], u_ceil = s_c[Sc];
    const float u_step = __fdiv_rn(__fsub_rn(u_ceil, u_floor), (float)Sf);
    const float bias = (u_ray != nullptr) ? u_ray[ray] : 0.5f;
    for (int k = lane; k < Sf; k += 32) {
      const float u = __fadd_rn(u_floor, __fmul_rn(__fadd_rn((float)k, bias), u_step));
      // p = upper_bound(cdf, u) - 1 clamped to [0, Sc-1]: cdf[p] <= u < cdf[p+1]
      int lo = 0, hi = Sc + 1;  // first index in [0,Sc] with cdf > u
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_c[mid] <= u) lo = mid + 1; else hi = mid;
      }
      int p = lo - 1;
      p = p < 0 ? 0 : (p > Sc - 1 ? Sc - 1 : p);
      const float c0 = s_c[p], c1 = s_c[p + 1], e0 = s_e[p], e1 = s_e[p + 1];
      const float dc = __fsub_rn(c1, c0);
      float s;
      if (dc < 1e-10f) {
        s = __fmul_rn(__fadd_rn(e0, e1), 0.5f);
      } else {
        const float scale = __fdiv_rn(__fsub_rn(e1, e0), dc);
        s = __fadd_rn(__fmul_rn(__fsub_rn(u, c0), scale), e0);
      }
      s_s[k] = s;
      if (out_idx != nullptr) out_idx[ray * Sf + k] = p;
    }
    __syncwarp();
    const float e_min = s_e[0], e_max = s_e[Sc];
    for (int k = lane; k <= Sf; k += 32) {
      float v;
      if (Sf == 1) {
        v = (k == 0) ? e_min : e_max;
      } else if (k == 0) {
        v = fmaxf(__fsub_rn(s_s[0], __fmul_rn(__fsub_rn(s_s[1], s_s[0]), 0.5f)), e_min);
      } else if (k == Sf) {
        v = fminf(__fadd_rn(s_s[Sf - 1], __fmul_rn(__fsub_rn(s_s[Sf - 1], s_s[Sf - 2]), 0.5f)), e_max);
      } else {
        v = __fmul_rn(__fadd_rn(s_s[k - 1], s_s[k]), 0.5f);
      }
      out_edges[ray * (Sf + 1) + k] = v;
    }
    __syncwarp();
  }
}

int warp_grid(int B) {
  int blocks = ceil_div(B, kWarps);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_resample_alloc(const float* t_coarse, const float* weights,
                                       const float* delta_coarse, int B, int Sc, int Sf,
                                       double far_t, float* t_start, float* t_end,
                                       int32_t* counts_out, int32_t* fail_flag, void* stream) {
  NB_CHECK_ARG(B >= 0 && Sc >= 1 && Sf > Sc, "resample_alloc: need B>=0, Sf>Sc>=1 (B=%d Sc=%d Sf=%d)",
               B, Sc, Sf);
  NB_CHECK_ARG(t_coarse && weights && delta_coarse && t_start && t_end && fail_flag,
               "resample_alloc: null pointer");
  if (B == 0) return NERFB200_OK;
  const size_t smem = (size_t)kWarps * (4 * Sc + 1 + Sf) * sizeof(float);
  NB_CHECK_ARG(smem <= 200 * 1024, "resample_alloc: Sc=%d Sf=%d exceed the shared-memory budget", Sc, Sf);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(resample_alloc_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  resample_alloc_kernel<<<warp_grid(B), kWarps * 32, smem, (cudaStream_t)stream>>>(
      t_coarse, weights, delta_coarse, B, Sc, Sf, (float)far_t, t_start, t_end, counts_out,
      fail_flag);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_resample_icdf(const float* edges, const float* cdf, const float* u_ray,
                                      int B, int Sc, int Sf, float* out_edges, int32_t* out_idx,
                                      void* stream) {
  NB_CHECK_ARG(B >= 0 && Sc >= 1 && Sf >= 1, "resample_icdf: bad shape B=%d Sc=%d Sf=%d", B, Sc, Sf);
  NB_CHECK_ARG(edges && cdf && out_edges, "resample_icdf: null pointer");
  if (B == 0) return NERFB200_OK;
  const size_t smem = (size_t)kWarps * (2 * (Sc + 1) + Sf) * sizeof(float);
  NB_CHECK_ARG(smem <= 200 * 1024, "resample_icdf: Sc=%d Sf=%d exceed the shared-memory budget", Sc, Sf);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(resample_icdf_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  resample_icdf_kernel<<<warp_grid(B), kWarps * 32, smem, (cudaStream_t)stream>>>(
      edges, cdf, u_ray, B, Sc, Sf, out_edges, out_idx);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
