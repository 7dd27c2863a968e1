// The rest of the nerfacc PropNet chain GarfModel.forward calls (reference garf/model_garf.py:
// 210-230,257; algorithm restated in oracle/ref_garf.py and oracle/ref_nerfacc.py — parity unpinned,
// nerfacc is not part of the reference tree): the "lindisp" map of normalised sample edges to ray
// distances, dense transmittance -> cdf with its backward, and the proposal loss (Mip-NeRF 360
// eq. 13) with its gradient w.r.t. the proposal cdf. All warp-per-ray, coalesced rows, HBM-bound;
// they replace ~40 eager ATen launches (cumsum, exp, cat, two searchsorted, two gather, clip, ...).
#include "common.cuh"

namespace nerfb200 {
namespace {

constexpr int kWarps = 8;

// t = 1 / (s / far + (1 - s) / near) for every edge; interval starts / ends / widths / mid-points.
__global__ void __launch_bounds__(256)
lindisp_kernel(const float* __restrict__ s_edges, float near, float far, int B, int E,
               float* __restrict__ t_edges, float* __restrict__ t0, float* __restrict__ t1,
               float* __restrict__ delta, float* __restrict__ t_mid) {
  const int S = E - 1;
  const long long total = (long long)B * E;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / E;
    const int k = (int)(i - r * E);
    const float s = s_edges[i];
    const float t = 1.0f / (s / far + (1.0f - s) / near);
    if (t_edges != nullptr) t_edges[i] = t;
    if (k < S) {
      const float sn = s_edges[i + 1];
      const float tn = 1.0f / (sn / far + (1.0f - sn) / near);
      const long long o = r * S + k;
      if (t0 != nullptr) t0[o] = t;
      if (t1 != nullptr) t1[o] = tn;
      if (delta != nullptr) delta[o] = tn - t;
      if (t_mid != nullptr) t_mid[o] = (t + tn) * 0.5f;
    }
  }
}

// trans_i = exp(-(cumsum(sigma delta)_i - sigma_i delta_i)); cdf = 1 - [trans, 0]
__global__ void __launch_bounds__(kWarps * 32)
trans_cdf_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ t0,
                     const float* __restrict__ t1, int B, int S, float* __restrict__ trans,
                     float* __restrict__ cdf) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (ray >= B) return;
  const float* sg = sigma + (size_t)ray * S;
  const float* a = t0 + (size_t)ray * S;
  const float* b = t1 + (size_t)ray * S;
  float carry = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int k = s0 + lane;
    const float sd = k < S ? sg[k] * (b[k] - a[k]) : 0.f;
    const float inc = warp_inclusive_scan(sd, lane) + carry;
    const float tr = __expf(-(inc - sd));
    if (k < S) {
      if (trans != nullptr) trans[(size_t)ray * S + k] = tr;
      cdf[(size_t)ray * (S + 1) + k] = 1.0f - tr;
    }
    carry = __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) cdf[(size_t)ray * (S + 1) + S] = 1.0f;
}

// d sigma_k = delta_k * sum_{k < i < S} (g_cdf_i - g_trans_i) trans_i   (cdf_S is the constant 1)
__global__ void __launch_bounds__(kWarps * 32)
trans_cdf_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ t0,
                     const float* __restrict__ t1, const float* __restrict__ g_cdf,
                     const float* __restrict__ g_trans, int B, int S, float* __restrict__ d_sigma) {
  const int lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (ray >= B) return;
  const float* sg = sigma + (size_t)ray * S;
  const float* a = t0 + (size_t)ray * S;
  const float* b = t1 + (size_t)ray * S;
  // total optical depth first, then walk the 32-sample blocks from the far end with a carried suffix sum
  float total = 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int k = s0 + lane;
    total += k < S ? sg[k] * (b[k] - a[k]) : 0.f;
  }
  total = warp_sum(total);
  float tail_depth = 0.f;     // optical depth of the blocks behind the current one
  float suffix = 0.f;         // sum of (g_cdf - g_trans) trans over the samples behind the current block
  const int n_blocks = (S + 31) / 32;
  for (int blk = n_blocks - 1; blk >= 0; --blk) {
    const int k = blk * 32 + lane;
    const float dl = k < S ? (b[k] - a[k]) : 0.f;
    const float sd = k < S ? sg[k] * dl : 0.f;
    const float blk_sum = warp_sum(sd);
    const float before = total - tail_depth - blk_sum;                 // depth in front of the block
    const float inc = warp_inclusive_scan(sd, lane);
    const float tr = __expf(-(before + inc - sd));
    float gi = 0.f;
    if (k < S) {
      gi = g_cdf != nullptr ? g_cdf[(size_t)ray * (S + 1) + k] : 0.f;
      if (g_trans != nullptr) gi -= g_trans[(size_t)ray * S + k];
      gi *= tr;
    }
    const float suf_incl = warp_inclusive_rscan(gi, lane);             // samples >= k inside the block
    if (k < S) d_sigma[(size_t)ray * S + k] = dl * (suf_incl - gi + suffix);
    suffix += __shfl_sync(0xffffffffu, suf_incl, 0);
    tail_depth += blk_sum;
  }
}

// index of the first key edge >= v (searchsorted right=False) / of the last key edge <= v
// (searchsorted right=True, minus one), both clamped to [0, n-1]
__device__ __forceinline__ int lower_bound(const float* e, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo < n - 1 ? lo : n - 1;
}
__device__ __forceinline__ int upper_bound_m1(const float* e, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= v) lo = mid + 1; else hi = mid;
  }
  const int r = lo - 1;
  return r < 0 ? 0 : (r > n - 1 ? n - 1 : r);
}

constexpr int kMaxKeyEdges = 1025;

// loss += scale * sum_i clip(w_i - w_outer_i, 0)^2 / (w_i + eps); d_cdf_k = its gradient
__global__ void __launch_bounds__(kWarps * 32)
prop_loss_kernel(const float* __restrict__ t_q, const float* __restrict__ cdf_q,
                 const float* __restrict__ t_k, const float* __restrict__ cdf_k, int B, int Sq,
                 int Sk, float eps, float scale, float* __restrict__ loss,
                 float* __restrict__ d_cdf_k) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int ray = blockIdx.x * kWarps + w;
  const int Ek = Sk + 1, Eq = Sq + 1;
  float* ek = sh + (size_t)w * 3 * Ek;      // key edges
  float* ck = ek + Ek;                      // key cdf
  float* gk = ck + Ek;                      // gradient w.r.t. the key cdf
  float part = 0.f;
  if (ray < B) {
    for (int i = lane; i < Ek; i += 32) {
      ek[i] = t_k[(size_t)ray * Ek + i];
      ck[i] = cdf_k[(size_t)ray * Ek + i];
      gk[i] = 0.f;
    }
    __syncwarp();
    const float* tq = t_q + (size_t)ray * Eq;
    const float* cq = cdf_q + (size_t)ray * Eq;
    for (int i = lane; i < Sq; i += 32) {
      const float ta = tq[i], tb = tq[i + 1];
      const float wq = cq[i + 1] - cq[i];
      const int il = upper_bound_m1(ek, Ek, ta);
      const int ir = lower_bound(ek, Ek, tb);
      const float excess = wq - (ck[ir] - ck[il]);
      if (excess > 0.f) {
        const float inv = 1.0f / (wq + eps);
        part += excess * excess * inv;
        if (d_cdf_k != nullptr) {
          const float g = 2.0f * excess * inv * scale;      // d/d w_outer = -g
          atomicAdd(&gk[ir], -g);
          atomicAdd(&gk[il], g);
        }
      }
    }
    __syncwarp();
    if (d_cdf_k != nullptr)
      for (int i = lane; i < Ek; i += 32) d_cdf_k[(size_t)ray * Ek + i] = gk[i];
  }
  part = warp_sum(part);
  if (lane == 0 && part != 0.f) atomicAdd(loss, part * scale);
}

int rows_grid(int B) { return ceil_div(B, kWarps); }

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_lindisp_intervals(const float* s_edges, float near, float far, int B, int E,
                                          float* t_edges, float* t_start, float* t_end, float* delta,
                                          float* t_mid, void* stream) {
  NB_CHECK_ARG(B >= 0 && E >= 2 && near > 0.f && far > near, "lindisp_intervals: bad arguments B=%d E=%d", B, E);
  NB_CHECK_ARG(s_edges != nullptr, "lindisp_intervals: null pointer");
  if (B == 0) return NERFB200_OK;
  int blocks = ceil_div((long long)B * E, 256);
  const int cap = sm_count() * 16;
  if (blocks > cap) blocks = cap;
  lindisp_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(s_edges, near, far, B, E, t_edges, t_start, t_end, delta, t_mid);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_trans_cdf_fwd(const float* sigma, const float* t_start, const float* t_end, int B,
                                      int S, float* out_trans, float* out_cdf, void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "trans_cdf_fwd: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(sigma && t_start && t_end && out_cdf, "trans_cdf_fwd: null pointer");
  if (B == 0) return NERFB200_OK;
  trans_cdf_fwd_kernel<<<rows_grid(B), kWarps * 32, 0, (cudaStream_t)stream>>>(sigma, t_start, t_end, B, S, out_trans, out_cdf);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_trans_cdf_bwd(const float* sigma, const float* t_start, const float* t_end,
                                      const float* g_cdf, const float* g_trans, int B, int S,
                                      float* d_sigma, void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "trans_cdf_bwd: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(sigma && t_start && t_end && d_sigma && (g_cdf || g_trans), "trans_cdf_bwd: null pointer");
  if (B == 0) return NERFB200_OK;
  trans_cdf_bwd_kernel<<<rows_grid(B), kWarps * 32, 0, (cudaStream_t)stream>>>(sigma, t_start, t_end, g_cdf, g_trans, B, S, d_sigma);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_prop_loss(const float* t_query, const float* cdf_query, const float* t_key,
                                  const float* cdf_key, int B, int Sq, int Sk, float eps, float scale,
                                  float* loss, float* d_cdf_key, void* stream) {
  NB_CHECK_ARG(B >= 0 && Sq >= 1 && Sk >= 1 && Sk + 1 <= kMaxKeyEdges, "prop_loss: bad shape B=%d Sq=%d Sk=%d", B, Sq, Sk);
  NB_CHECK_ARG(t_query && cdf_query && t_key && cdf_key && loss, "prop_loss: null pointer");
  if (B == 0) return NERFB200_OK;
  const size_t smem = (size_t)kWarps * 3 * (Sk + 1) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(prop_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kWarps * 3 * kMaxKeyEdges * sizeof(float))));
    configured = true;
  }
  prop_loss_kernel<<<rows_grid(B), kWarps * 32, smem, (cudaStream_t)stream>>>(t_query, cdf_query, t_key, cdf_key, B, Sq, Sk,
                                                                              eps, scale, loss, d_cdf_key);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
