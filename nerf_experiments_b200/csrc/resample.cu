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

constexpr int kRankBuckets = 64;   // histogram buckets of the remainder-rank selection

// Work per ray, one warp: O(Sc) for the counts, O(Sc + m^2) for the remainder ranks (m = size of
// the one histogram bucket the threshold falls into, typically 1-3), O(Sf) for the expansion.
__global__ void __launch_bounds__(kWarps * 32)
resample_alloc_kernel(const float* __restrict__ t_coarse, const float* __restrict__ weights,
                      const float* __restrict__ delta_coarse, int B, int Sc, int Sf, float far_t,
                      float* __restrict__ t_start, float* __restrict__ t_end,
                      int32_t* __restrict__ counts_out, int32_t* __restrict__ fail_flag) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int per_warp = 4 * Sc + 1 + 2 * Sf + kRankBuckets;
  float* s_err = smem + (size_t)warp * per_warp;  // [Sc]   remainder e_i, later bin width d_i
  float* s_n = s_err + Sc;                        // [Sc]   n_i
  float* s_cum = s_n + Sc;                        // [Sc+1] exclusive cumsum of n
  float* s_tc = s_cum + Sc + 1;                   // [Sc]   coarse t
  float* s_t = s_tc + Sc;                         // [Sf]   fine t
  int* s_bin = reinterpret_cast<int*>(s_t + Sf);  // [Sf]   bin of every fine sample
  int* s_hist = s_bin + Sf;                       // [kRankBuckets]
  const float n_new = (float)(Sf - Sc);
  const int chunk = (Sf + 31) / 32;               // consecutive fine samples per lane in the scan

  for (long long ray = (long long)blockIdx.x * kWarps + warp; ray < B;
       ray += (long long)gridDim.x * kWarps) {
    const float* w_row = weights + ray * Sc;
    const float wsum = lane_strided_sum(w_row, Sc, lane);
    float nsum = 0.f;
    for (int b = lane; b < kRankBuckets; b += 32) s_hist[b] = 0;
    __syncwarp();
    for (int i = lane; i < Sc; i += 32) {
      const float p = __fdiv_rn(w_row[i], wsum);   // weights / weights.sum            (:215)
      const float raw = __fmul_rn(p, n_new);       // * (n_samples - n_bins)           (:216)
      const float fl = floorf(raw);                //                                   (:217)
      const float e = __fsub_rn(raw, fl);          //                                   (:218)
      s_err[i] = e;
      s_n[i] = fl;
      nsum += fl;  // integers: exact in any order
      s_tc[i] = t_coarse[ray * Sc + i];
      // order-preserving bucket of the remainder (e in [0,1); NaN lands in bucket 0)
      int bk = (int)(e * (float)kRankBuckets);
      bk = bk < 0 ? 0 : (bk > kRankBuckets - 1 ? kRankBuckets - 1 : bk);
      atomicAdd(&s_hist[bk], 1);
    }
    nsum = warp_sum(nsum);
    const float excess_f = __fsub_rn(n_new, nsum);   // n_samples - n_bins - sum        (:224)
    __syncwarp();
    // error_rank = argsort(argsort(err)) with ties to the lowest index, and the bins whose rank
    // is >= Sc - excess get one more sample (:225-226)  <=>  the `excess` largest keys (e, index)
    // get one more. Bucket counts locate the bucket b* the threshold falls into; only its
    // members need exact comparisons.
    const int excess = (excess_f > 0.f && excess_f <= (float)Sc) ? (int)excess_f : 0;
    int bstar = -1, need = 0;
    {
      // lane l owns buckets 2l, 2l+1 (kRankBuckets == 64)
      const int h0 = s_hist[2 * lane], h1 = s_hist[2 * lane + 1];
      const int tot = h0 + h1;
      int incl = tot;   // inclusive suffix sum over lanes >= lane
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
      }
      const int above1 = incl - tot;      // elements in buckets above 2l+1
      const int above0 = above1 + h1;     // elements in buckets above 2l
      const bool hit1 = above1 < excess && excess <= above1 + h1;
      const bool hit0 = above0 < excess && excess <= above0 + h0;
      const unsigned m1 = __ballot_sync(0xffffffffu, hit1), m0 = __ballot_sync(0xffffffffu, hit0);
      if (m1 | m0) {
        const int src = __ffs(m1 ? m1 : m0) - 1;
        bstar = 2 * src + (m1 ? 1 : 0);
        need = excess - __shfl_sync(0xffffffffu, m1 ? above1 : above0, src);
      }
    }
    bool bad = false;
    for (int base = 0; base < Sc; base += 32) {
      const int i = base + lane;
      const float e = (i < Sc) ? s_err[i] : 0.f;
      int bk = (int)(e * (float)kRankBuckets);
      bk = bk < 0 ? 0 : (bk > kRankBuckets - 1 ? kRankBuckets - 1 : bk);
      const bool member = (i < Sc) && (bk == bstar);
      int greater = 0;   // members of b* with a larger key (e, index)
      // every lane walks the member list of every 32-element group (warp-uniform loops)
      for (int gb = 0; gb < Sc; gb += 32) {
        const int jj = gb + lane;
        const float ej_own = (jj < Sc) ? s_err[jj] : 0.f;
        int bj = (int)(ej_own * (float)kRankBuckets);
        bj = bj < 0 ? 0 : (bj > kRankBuckets - 1 ? kRankBuckets - 1 : bj);
        unsigned mm = __ballot_sync(0xffffffffu, (jj < Sc) && (bj == bstar));
        while (mm) {
          const int j = gb + __ffs(mm) - 1;
          mm &= mm - 1u;
          const float ej = s_err[j];
          greater += (ej > e) || (ej == e && j > i);
        }
      }
      if (i < Sc) {
        const float add = (bk > bstar && bstar >= 0) || (member && greater < need) ? 1.f : 0.f;   // (:226)
        const float n = __fadd_rn(__fadd_rn(s_n[i], add), 1.f);                                    // (:227)
        bad |= !(n >= 0.f);
        s_cum[i + 1] = n;  // staged; turned into a cumsum next
      }
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
    // bin of every fine sample: scatter the bin starts, then a running maximum
    for (int k = lane; k < Sf; k += 32) s_bin[k] = 0;
    __syncwarp();
    for (int i = lane; i < Sc; i += 32) {
      const float c = s_cum[i];
      if (c >= 0.f && c < (float)Sf && s_n[i] >= 1.f) s_bin[(int)c] = i;   // starts are distinct: n_i >= 1
    }
    __syncwarp();
    {
      const int k0 = lane * chunk;
      int run = 0;
      for (int k = k0; k < k0 + chunk && k < Sf; ++k) run = max(run, s_bin[k]);
      int excl = run;   // exclusive prefix maximum over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, excl, o);
        if (lane >= o) excl = max(excl, v);
      }
      excl = __shfl_up_sync(0xffffffffu, excl, 1);
      run = lane == 0 ? 0 : excl;
      for (int k = k0; k < k0 + chunk && k < Sf; ++k) {
        run = max(run, s_bin[k]);
        s_bin[k] = run;
      }
    }
    __syncwarp();
    // expand: t_k = t_c[i] + ((k - cum_i) * delta_i) / n_i   for cum_i <= k < cum_{i+1} (:262-269)
    for (int k = lane; k < Sf; k += 32) {
      const int lo = s_bin[k];
      const float num = __fmul_rn(__fsub_rn((float)k, s_cum[lo]), s_err[lo]);
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

// ---- a11 fast path: Sc <= 64 coarse bins, Sf <= 256 fine samples -----------------------------
// Same arithmetic as resample_alloc_kernel (bit-identical results), organised for throughput:
// lane owns coarse bins lane and lane+32 in registers and the 8 consecutive fine samples
// 8*lane .. 8*lane+7; shared memory only carries the remainder histogram, the bin-start scatter
// and one float4 table row per coarse bin. ~4x fewer instructions per ray than the generic
// kernel, 128-bit stores.
constexpr int kFastSc = 64, kFastSf = 256;
constexpr int kFastBlocksPerSm = 12;   // 48 warps per SM: the kernels are latency bound

__global__ void __launch_bounds__(kWarps * 32, kFastBlocksPerSm)
resample_alloc_fast_kernel(const float* __restrict__ t_coarse, const float* __restrict__ weights,
                           const float* __restrict__ delta_coarse, int B, int Sc, int Sf, float far_t,
                           float* __restrict__ t_start, float* __restrict__ t_end,
                           int32_t* __restrict__ counts_out, int32_t* __restrict__ fail_flag) {
  __shared__ __align__(16) float4 s_tab_all[kWarps][kFastSc];     // (t_c, cum, delta, n) per coarse bin
  __shared__ __align__(16) int s_bin_all[kWarps][kFastSf];        // bin-start scatter
  __shared__ int s_hist_all[kWarps][kRankBuckets];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float4* s_tab = s_tab_all[warp];
  int* s_bin = s_bin_all[warp];
  int* s_hist = s_hist_all[warp];
  const float n_new = (float)(Sf - Sc);
  const int i0 = lane, i1 = lane + 32;
  const bool v0 = i0 < Sc, v1 = i1 < Sc;
  const bool vec_ok = (Sf & 3) == 0;

  for (long long ray = (long long)blockIdx.x * kWarps + warp; ray < B;
       ray += (long long)gridDim.x * kWarps) {
    const long long cbase = ray * Sc;
    // all six loads of the ray are issued before the first use
    const float w0 = v0 ? weights[cbase + i0] : 0.f, w1 = v1 ? weights[cbase + i1] : 0.f;
    const float tc0 = v0 ? t_coarse[cbase + i0] : 0.f, tc1 = v1 ? t_coarse[cbase + i1] : 0.f;
    const float d0 = v0 ? delta_coarse[cbase + i0] : 0.f, d1 = v1 ? delta_coarse[cbase + i1] : 0.f;
    s_hist[lane] = 0;
    s_hist[lane + 32] = 0;
    *reinterpret_cast<int4*>(s_bin + 8 * lane) = make_int4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(s_bin + 8 * lane + 4) = make_int4(0, 0, 0, 0);
    // weights.sum in the oracle's order: lane-strided partial sums, then a xor butterfly
    float wsum = 0.f;
    if (v0) wsum = __fadd_rn(wsum, w0);
    if (v1) wsum = __fadd_rn(wsum, w1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wsum = __fadd_rn(wsum, __shfl_xor_sync(0xffffffffu, wsum, o));
    __syncwarp();
    float fl0 = 0.f, fl1 = 0.f, e0 = 0.f, e1 = 0.f;
    int b0 = -2, b1 = -2;
    if (v0) {
      const float raw = __fmul_rn(__fdiv_rn(w0, wsum), n_new);     // (:215-216)
      fl0 = floorf(raw);                                            // (:217)
      e0 = __fsub_rn(raw, fl0);                                     // (:218)
      b0 = (int)(e0 * (float)kRankBuckets);
      b0 = b0 < 0 ? 0 : (b0 > kRankBuckets - 1 ? kRankBuckets - 1 : b0);
      atomicAdd(&s_hist[b0], 1);
    }
    if (v1) {
      const float raw = __fmul_rn(__fdiv_rn(w1, wsum), n_new);
      fl1 = floorf(raw);
      e1 = __fsub_rn(raw, fl1);
      b1 = (int)(e1 * (float)kRankBuckets);
      b1 = b1 < 0 ? 0 : (b1 > kRankBuckets - 1 ? kRankBuckets - 1 : b1);
      atomicAdd(&s_hist[b1], 1);
    }
    const float nsum = warp_sum(fl0 + fl1);                         // integers: exact in any order
    const float excess_f = __fsub_rn(n_new, nsum);                  // (:224)
    const int excess = (excess_f > 0.f && excess_f <= (float)Sc) ? (int)excess_f : 0;
    __syncwarp();
    // the `excess` largest keys (e, index) get one more sample (:225-226): locate the histogram
    // bucket the threshold falls into, compare exactly only inside it
    int bstar = -1, need = 0;
    {
      const int h0 = s_hist[2 * lane], h1 = s_hist[2 * lane + 1];
      const int tot = h0 + h1;
      int incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += v;
      }
      const int above1 = incl - tot, above0 = above1 + h1;
      const unsigned m1 = __ballot_sync(0xffffffffu, above1 < excess && excess <= above1 + h1);
      const unsigned m0 = __ballot_sync(0xffffffffu, above0 < excess && excess <= above0 + h0);
      if (m1 | m0) {
        const int src = __ffs(m1 ? m1 : m0) - 1;
        bstar = 2 * src + (m1 ? 1 : 0);
        need = excess - __shfl_sync(0xffffffffu, m1 ? above1 : above0, src);
      }
    }
    int g0 = 0, g1 = 0;   // members of b* with a larger key than my elements
    {
      unsigned mm = __ballot_sync(0xffffffffu, b0 == bstar);
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1u;
        const float ej = __shfl_sync(0xffffffffu, e0, j);
        g0 += (ej > e0) || (ej == e0 && j > i0);
        g1 += (ej > e1) || (ej == e1 && j > i1);
      }
      mm = __ballot_sync(0xffffffffu, b1 == bstar);
      while (mm) {
        const int j = __ffs(mm) - 1;
        mm &= mm - 1u;
        const float ej = __shfl_sync(0xffffffffu, e1, j);
        g0 += (ej > e0) || (ej == e0 && j + 32 > i0);
        g1 += (ej > e1) || (ej == e1 && j + 32 > i1);
      }
    }
    const float add0 = ((b0 > bstar && bstar >= 0) || (b0 == bstar && g0 < need)) ? 1.f : 0.f;
    const float add1 = ((b1 > bstar && bstar >= 0) || (b1 == bstar && g1 < need)) ? 1.f : 0.f;
    const float n0 = v0 ? __fadd_rn(__fadd_rn(fl0, add0), 1.f) : 0.f;   // (:227)
    const float n1 = v1 ? __fadd_rn(__fadd_rn(fl1, add1), 1.f) : 0.f;
    bool bad = (v0 && !(n0 >= 0.f)) || (v1 && !(n1 >= 0.f));
    // exclusive cumsum of n over the bins (integers, exact)
    const float inc0 = warp_inclusive_scan(n0, lane);
    const float tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    const float inc1 = warp_inclusive_scan(n1, lane);
    const float total = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
    const float c0 = inc0 - n0, c1 = tot0 + inc1 - n1;
    bad |= (total != (float)Sf);                                        // (:235)
    bad = __any_sync(0xffffffffu, bad);
    if (bad && lane == 0) atomicExch(fail_flag, 1);
    if (counts_out != nullptr) {
      if (v0) counts_out[cbase + i0] = (int32_t)n0;
      if (v1) counts_out[cbase + i1] = (int32_t)n1;
    }
    if (v0) {
      s_tab[i0] = make_float4(tc0, c0, d0, n0);
      if (c0 >= 0.f && c0 < (float)Sf && n0 >= 1.f) s_bin[(int)c0] = i0;   // bin starts are distinct (n >= 1)
    }
    if (v1) {
      s_tab[i1] = make_float4(tc1, c1, d1, n1);
      if (c1 >= 0.f && c1 < (float)Sf && n1 >= 1.f) s_bin[(int)c1] = i1;
    }
    __syncwarp();
    // bin of my 8 consecutive fine samples: running maximum over the scattered bin starts
    int bin[8];
    {
      const int4 a = *reinterpret_cast<const int4*>(s_bin + 8 * lane);
      const int4 b = *reinterpret_cast<const int4*>(s_bin + 8 * lane + 4);
      bin[0] = a.x; bin[1] = max(bin[0], a.y); bin[2] = max(bin[1], a.z); bin[3] = max(bin[2], a.w);
      bin[4] = max(bin[3], b.x); bin[5] = max(bin[4], b.y); bin[6] = max(bin[5], b.z); bin[7] = max(bin[6], b.w);
      int excl = bin[7];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, excl, o);
        if (lane >= o) excl = max(excl, v);
      }
      excl = __shfl_up_sync(0xffffffffu, excl, 1);
      if (lane == 0) excl = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) bin[q] = max(bin[q], excl);
    }
    // expand: t_k = t_c[i] + ((k - cum_i) * delta_i) / n_i   for cum_i <= k < cum_{i+1} (:262-269)
    float t[9];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 tb = s_tab[bin[q]];
      const float num = __fmul_rn(__fsub_rn((float)(8 * lane + q), tb.y), tb.z);
      // 0 / n = +0 exactly; dividing a stand-in keeps every lane on the division's fast path
      const float quo = __fdiv_rn(num == 0.f ? 1.f : num, tb.w);
      t[q] = __fadd_rn(tb.x, num == 0.f ? 0.f : quo);
    }
    t[8] = __shfl_down_sync(0xffffffffu, t[0], 1);
    __syncwarp();   // every lane has read s_bin / s_tab: the next ray may overwrite them
    const int k0 = 8 * lane;
    float te[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) te[q] = (k0 + q + 1 < Sf) ? t[q + 1] : far_t;               // (:127-130)
    float* ps = t_start + ray * Sf + k0;
    float* pe = t_end + ray * Sf + k0;
    if (vec_ok && k0 + 8 <= Sf) {
      st_stream4(ps, make_float4(t[0], t[1], t[2], t[3]));
      st_stream4(ps + 4, make_float4(t[4], t[5], t[6], t[7]));
      st_stream4(pe, make_float4(te[0], te[1], te[2], te[3]));
      st_stream4(pe + 4, make_float4(te[4], te[5], te[6], te[7]));
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (k0 + q < Sf) { ps[q] = t[q]; pe[q] = te[q]; }
    }
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

// ---- a12 fast path: Sc <= 64 coarse intervals, Sf <= 256 samples -----------------------------
// Same results as resample_icdf_kernel for a sorted cdf (bit-identical), with a third of the
// instructions: instead of searching a bin for every sample, every BIN finds the first sample
// that falls into it — the targets u_k are an arithmetic progression, so an approximate
// division gives a candidate that two exact comparisons with the kernel's own u_k formula fix
// up — and scatters its index there; a running maximum over the samples (sequential inside a
// lane, one warp scan across lanes) then hands every sample the last bin that starts at or
// before it, which is what upper_bound(cdf, u_k) - 1 returns. The interpolation slope is
// computed once per bin. Outputs are staged in shared memory and leave as coalesced rows.
template <int PER>   // consecutive samples per lane = ceil(Sf / 32): loops unroll without predicates
__global__ void __launch_bounds__(kWarps * 32, kFastBlocksPerSm)
resample_icdf_fast_kernel(const float* __restrict__ edges, const float* __restrict__ cdf,
                          const float* __restrict__ u_ray, int B, int Sc, int Sf,
                          float* __restrict__ out_edges, int32_t* __restrict__ out_idx) {
  __shared__ float2 s_ec_all[kWarps][kFastSc + 1];      // (edge, cdf)
  __shared__ float s_sc_all[kWarps][kFastSc];           // (e1 - e0) / (c1 - c0) per bin, NaN: empty bin
  __shared__ float s_s_all[kWarps][kFastSf];            // sample centres
  __shared__ int s_p_all[kWarps][kFastSf];              // bin starts, then the bin of every sample
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float2* s_ec = s_ec_all[warp];
  float* s_sc = s_sc_all[warp];
  float* s_s = s_s_all[warp];
  int* s_p = s_p_all[warp];
  constexpr int per = PER;
  const bool full = (Sf == 32 * PER);                   // no ragged last lane: the common shapes
  for (long long ray = (long long)blockIdx.x * kWarps + warp; ray < B;
       ray += (long long)gridDim.x * kWarps) {
    const long long base = ray * (Sc + 1);
    for (int i = lane; i <= Sc; i += 32) s_ec[i] = make_float2(edges[base + i], cdf[base + i]);
#pragma unroll
    for (int j = 0; j < PER; ++j) s_p[lane + 32 * j] = 0;   // samples below cdf[1] belong to bin 0
    const float bias = (u_ray != nullptr) ? u_ray[ray] : 0.5f;
    __syncwarp();
    const float u_floor = s_ec[0].y, u_ceil = s_ec[Sc].y;
    const float u_step = __fdiv_rn(__fsub_rn(u_ceil, u_floor), (float)Sf);
    const float inv_step = u_step > 0.f ? __frcp_rn(u_step) : 0.f;
    auto u_of = [&](int k) { return __fadd_rn(u_floor, __fmul_rn(__fadd_rn((float)k, bias), u_step)); };
    for (int i = lane; i < Sc; i += 32) {
      const float2 a = s_ec[i], b = s_ec[i + 1];
      const float dc = __fsub_rn(b.y, a.y);
      // one IEEE division per bin instead of one per sample: same operands, same result
      s_sc[i] = dc < 1e-10f ? __int_as_float(0x7fc00000) : __fdiv_rn(__fsub_rn(b.x, a.x), dc);
      if (i > 0) {
        // first sample k with cdf[i] <= u_k (u_k is non-decreasing in k)
        int k = (int)ceilf(fminf(fmaxf((a.y - u_floor) * inv_step - bias, 0.f), (float)Sf));
        while (k > 0 && u_of(k - 1) >= a.y) --k;
        while (k < Sf && u_of(k) < a.y) ++k;
        if (k < Sf) atomicMax(&s_p[k], i);
      }
    }
    __syncwarp();
    // running maximum of the scattered bin starts = bin of every sample
    const int k0 = lane * per;
    int loc[PER];
    int m = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      if (full || k0 + j < Sf) m = max(m, s_p[k0 + j]);
      loc[j] = m;
    }
    int incl = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl = max(incl, v);
    }
    int before = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) before = 0;
    __syncwarp();                                        // every lane has read its s_p entries
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int k = k0 + j;
      if (full || k < Sf) {
        const int p = max(loc[j], before);
        const float u = u_of(k);
        const float2 a = s_ec[p];
        const float scale = s_sc[p];
        float sv;
        if (scale != scale) {     // NaN marks a bin without cdf mass
          sv = __fmul_rn(__fadd_rn(a.x, s_ec[p + 1].x), 0.5f);
        } else {
          sv = __fadd_rn(__fmul_rn(__fsub_rn(u, a.y), scale), a.x);
        }
        s_s[k] = sv;
        if (out_idx != nullptr) s_p[k] = p;
      }
    }
    __syncwarp();
    const float e_min = s_ec[0].x, e_max = s_ec[Sc].x;
    float* out_row = out_edges + ray * (Sf + 1);
    if (Sf == 1) {
      if (lane < 2) st_stream(out_row + lane, lane == 0 ? e_min : e_max);
    } else {
      // interior edges: midpoints of neighbouring samples; the two ends by one lane each
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int k = lane + 32 * j;
        if (k >= 1 && (full || k < Sf)) st_stream(out_row + k, __fmul_rn(__fadd_rn(s_s[k - 1], s_s[k]), 0.5f));
      }
      if (lane == 0)
        st_stream(out_row, fmaxf(__fsub_rn(s_s[0], __fmul_rn(__fsub_rn(s_s[1], s_s[0]), 0.5f)), e_min));
      if (lane == 1)
        st_stream(out_row + Sf, fminf(__fadd_rn(s_s[Sf - 1], __fmul_rn(__fsub_rn(s_s[Sf - 1], s_s[Sf - 2]), 0.5f)), e_max));
    }
    if (out_idx != nullptr) {
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int k = lane + 32 * j;
        if (full || k < Sf) out_idx[ray * Sf + k] = s_p[k];
      }
    }
    __syncwarp();
  }
}

int warp_grid(int B, int blocks_per_sm = 8) {
  int blocks = ceil_div(B, kWarps);
  const int cap = sm_count() * blocks_per_sm;
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
  if (Sc <= kFastSc && Sf <= kFastSf) {
    resample_alloc_fast_kernel<<<warp_grid(B, kFastBlocksPerSm), kWarps * 32, 0, (cudaStream_t)stream>>>(
        t_coarse, weights, delta_coarse, B, Sc, Sf, (float)far_t, t_start, t_end, counts_out, fail_flag);
    count_launch();
    NB_CHECK_LAUNCH();
    return NERFB200_OK;
  }
  const size_t smem = (size_t)kWarps * (4 * Sc + 1 + 2 * Sf + kRankBuckets) * sizeof(float);
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
  if (Sc <= kFastSc && Sf <= kFastSf) {
    const int grid = warp_grid(B, kFastBlocksPerSm);
#define NB_ICDF_CASE(P)                                                                                  \
  case P:                                                                                                \
    resample_icdf_fast_kernel<P><<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(edges, cdf, u_ray, B, Sc, \
                                                                                  Sf, out_edges, out_idx);    \
    break;
    switch ((Sf + 31) / 32) {
      NB_ICDF_CASE(1) NB_ICDF_CASE(2) NB_ICDF_CASE(3) NB_ICDF_CASE(4)
      NB_ICDF_CASE(5) NB_ICDF_CASE(6) NB_ICDF_CASE(7) NB_ICDF_CASE(8)
    }
#undef NB_ICDF_CASE
    count_launch();
    NB_CHECK_LAUNCH();
    return NERFB200_OK;
  }
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
