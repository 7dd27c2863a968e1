// Alpha compositing along rays: warp-per-ray scan kernels (HBM-bound, fp32).
//
// Mirrors NerfInterpolation._render_rays (reference barf/model_interpolation.py:316-353) and
// the nerfacc.rendering arithmetic used by GarfModel.forward (garf/model_garf.py:223-236).
//
// Layout: one warp owns one ray. With S % 4 == 0 each lane moves float4 (4 consecutive
// samples; 3 float4 for their rgb) per 128-sample block, so every request is a fully
// coalesced 512 B (sigma/delta/w) or 1536 B (rgb) line group. The exclusive cumsum of the
// optical depth is an in-register prefix + one warp shuffle scan + a carried block sum.
#include <float.h>

#include "common.cuh"

namespace nerfb200 {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxBlocks = 32;  // max 128-sample (VEC=4) or 32-sample (VEC=1) blocks per ray

__device__ __forceinline__ float optical_b(float sigma, float delta, int flavour) {
  // BARF flavour keeps the reference's two scalar multipliers, which are NOT an exact no-op
  // in fp32 (barf/model_interpolation.py:340, barf/magic.py:2).
  float t = -sigma * delta;
  if (flavour == NERFB200_COMPOSITE_BARF) {
    t = __fmul_rn(t, 3.0f);
    t = __fmul_rn(t, (float)(1.0 / 3.0));
  }
  return t;
}

template <int VEC>
struct Block {
  float b[VEC], T[VEC], eb[VEC];
};

// loads VEC consecutive floats of a row (zero beyond S)
template <int VEC>
__device__ __forceinline__ void load_vec(const float* row, int s0, int S, float (&v)[VEC]) {
  if (VEC == 4) {
    if (s0 < S) {
      float4 t = ld_stream4(row + s0);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[VEC - 1] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
  } else {
    v[0] = (s0 < S) ? ld_stream(row + s0) : 0.f;
  }
}

// same, through L1 (the row is re-read by a later pass of the same warp)
template <int VEC>
__device__ __forceinline__ void load_vec_cached(const float* row, int s0, int S, float (&v)[VEC]) {
  if (VEC == 4) {
    if (s0 < S) {
      float4 t = __ldg(reinterpret_cast<const float4*>(row + s0));
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[VEC - 1] = t.w;
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
  } else {
    v[0] = (s0 < S) ? __ldg(row + s0) : 0.f;
  }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* row, int s0, int S, const float (&v)[VEC]) {
  if (s0 >= S) return;
  if (VEC == 4) {
    st_stream4(row + s0, make_float4(v[0], v[1], v[2], v[VEC - 1]));
  } else {
    st_stream(row + s0, v[0]);
  }
}

// For one block: optical depth b, transmittance T (exclusive), exp(b).
template <int VEC>
__device__ __forceinline__ float block_transmittance(const float (&sig)[VEC],
                                                     const float (&del)[VEC], int s0, int S,
                                                     int flavour, float carry, int lane,
                                                     Block<VEC>& o) {
  float pre[VEC];
  float run = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    o.b[i] = (s0 + i < S) ? optical_b(sig[i], del[i], flavour) : 0.f;
    pre[i] = run;  // exclusive within the lane
    run += o.b[i];
  }
  float incl = warp_inclusive_scan(run, lane);
  float excl = incl - run;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    o.T[i] = expf(carry + excl + pre[i]);
    o.eb[i] = expf(o.b[i]);
  }
  return __shfl_sync(0xffffffffu, incl, 31);  // block total
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ sigma, const float* __restrict__ delta,
                     const float* __restrict__ rgb, const float* __restrict__ t_mid, int B,
                     int S, int flavour, float* __restrict__ out_rgb,
                     float* __restrict__ out_w, float* __restrict__ out_opacity,
                     float* __restrict__ out_depth) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nb = (S + 32 * VEC - 1) / (32 * VEC);
  for (long long ray = (long long)blockIdx.x * kWarpsPerBlock + warp; ray < B;
       ray += (long long)gridDim.x * kWarpsPerBlock) {
    const float* sig_row = sigma + ray * S;
    const float* del_row = delta + ray * S;
    const float* rgb_row = rgb + ray * S * 3;
    float carry = 0.f;
    float acc[3] = {0.f, 0.f, 0.f};
    float acc_o = 0.f, acc_d = 0.f;
    for (int j = 0; j < nb; ++j) {
      const int s0 = j * 32 * VEC + lane * VEC;
      float sig[VEC], del[VEC], c[3 * VEC];
      load_vec<VEC>(sig_row, s0, S, sig);
      load_vec<VEC>(del_row, s0, S, del);
      if (VEC == 4) {
        if (s0 < S) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float4 t = ld_stream4(rgb_row + (size_t)s0 * 3 + q * 4);
            c[q * 4 + 0] = t.x; c[q * 4 + 1] = t.y; c[q * 4 + 2] = t.z; c[q * 4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 3 * VEC; ++q) c[q] = 0.f;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 3; ++q) c[q] = (s0 < S) ? ld_stream(rgb_row + (size_t)s0 * 3 + q) : 0.f;
      }
      Block<VEC> blk;
      float tot = block_transmittance<VEC>(sig, del, s0, S, flavour, carry, lane, blk);
      carry += tot;
      float w[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        w[i] = (s0 + i < S) ? blk.T[i] * (1.f - blk.eb[i]) : 0.f;
        acc[0] += w[i] * c[3 * i + 0];
        acc[1] += w[i] * c[3 * i + 1];
        acc[2] += w[i] * c[3 * i + 2];
        acc_o += w[i];
      }
      if (out_depth != nullptr) {
        float tm[VEC];
        load_vec<VEC>(t_mid + ray * S, s0, S, tm);
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc_d += w[i] * tm[i];
      }
      if (out_w != nullptr) store_vec<VEC>(out_w + ray * S, s0, S, w);
    }
    acc[0] = warp_sum(acc[0]);
    acc[1] = warp_sum(acc[1]);
    acc[2] = warp_sum(acc[2]);
    acc_o = warp_sum(acc_o);
    acc_d = warp_sum(acc_d);
    if (lane == 0) {
      out_rgb[ray * 3 + 0] = acc[0];
      out_rgb[ray * 3 + 1] = acc[1];
      out_rgb[ray * 3 + 2] = acc[2];
      if (out_opacity != nullptr) out_opacity[ray] = acc_o;
      if (out_depth != nullptr) out_depth[ray] = acc_d / fmaxf(acc_o, FLT_EPSILON);
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float* __restrict__ sigma, const float* __restrict__ delta,
                     const float* __restrict__ rgb, const float* __restrict__ t_mid,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_w,
                     const float* __restrict__ g_opacity, const float* __restrict__ g_depth,
                     int B, int S, int flavour, float* __restrict__ d_sigma,
                     float* __restrict__ d_rgb) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nb = (S + 32 * VEC - 1) / (32 * VEC);
  const float kappa_a = (flavour == NERFB200_COMPOSITE_BARF) ? (float)(1.0 / 3.0) : 1.f;
  const float kappa_b = (flavour == NERFB200_COMPOSITE_BARF) ? 3.f : 1.f;
  for (long long ray = (long long)blockIdx.x * kWarpsPerBlock + warp; ray < B;
       ray += (long long)gridDim.x * kWarpsPerBlock) {
    const float* sig_row = sigma + ray * S;
    const float* del_row = delta + ray * S;
    const float* rgb_row = rgb + ray * S * 3;
    const float g0 = g_rgb[ray * 3 + 0], g1 = g_rgb[ray * 3 + 1], g2 = g_rgb[ray * 3 + 2];
    const float g_op = (g_opacity != nullptr) ? g_opacity[ray] : 0.f;
    const float g_dp = (g_depth != nullptr) ? g_depth[ray] : 0.f;

    // pass A: prefix carry of the optical depth per block (and opacity / depth numerators
    // when the depth gradient needs them). Plain loads: the lines stay in L1/L2 for pass B.
    float carries[kMaxBlocks];
    float carry = 0.f, O = 0.f, D = 0.f;
    const bool need_od = (g_depth != nullptr);
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
      const int s0 = j * 32 * VEC + lane * VEC;
      carries[j] = carry;
      if (nb == 1 && !need_od) break;
      float sig[VEC], del[VEC];
      load_vec_cached<VEC>(sig_row, s0, S, sig);
      load_vec_cached<VEC>(del_row, s0, S, del);
      Block<VEC> blk;
      float tot = block_transmittance<VEC>(sig, del, s0, S, flavour, carry, lane, blk);
      if (need_od) {
        float tm[VEC];
        load_vec_cached<VEC>(t_mid + ray * S, s0, S, tm);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          float w = (s0 + i < S) ? blk.T[i] * (1.f - blk.eb[i]) : 0.f;
          O += w;
          D += w * tm[i];
        }
      }
      carry += tot;
    }
    float inv_o = 0.f, d_corr = 0.f;
    if (need_od) {
      O = warp_sum(O);
      D = warp_sum(D);
      float Oc = fmaxf(O, FLT_EPSILON);
      inv_o = 1.f / Oc;
      d_corr = (O > FLT_EPSILON) ? D * inv_o * inv_o : 0.f;
    }

    // pass B: blocks in reverse, suffix scan of G_k w_k.
    float rcarry = 0.f;
#pragma unroll 1
    for (int j = nb - 1; j >= 0; --j) {
      const int s0 = j * 32 * VEC + lane * VEC;
      float sig[VEC], del[VEC], c[3 * VEC], h[VEC];
      load_vec<VEC>(sig_row, s0, S, sig);
      load_vec<VEC>(del_row, s0, S, del);
      if (VEC == 4) {
        if (s0 < S) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            float4 t = ld_stream4(rgb_row + (size_t)s0 * 3 + q * 4);
            c[q * 4 + 0] = t.x; c[q * 4 + 1] = t.y; c[q * 4 + 2] = t.z; c[q * 4 + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 3 * VEC; ++q) c[q] = 0.f;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 3; ++q) c[q] = (s0 < S) ? ld_stream(rgb_row + (size_t)s0 * 3 + q) : 0.f;
      }
      if (g_w != nullptr) {
        load_vec<VEC>(g_w + ray * S, s0, S, h);
      } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) h[i] = 0.f;
      }
      if (need_od || g_opacity != nullptr) {
        float tm[VEC];
        if (need_od) load_vec<VEC>(t_mid + ray * S, s0, S, tm);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          h[i] += g_op;
          if (need_od) h[i] += g_dp * (tm[i] * inv_o - d_corr);
        }
      }
      Block<VEC> blk;
      block_transmittance<VEC>(sig, del, s0, S, flavour, carries[j], lane, blk);
      float w[VEC], G[VEC], x[VEC], suf[VEC];
      float run = 0.f;
#pragma unroll
      for (int i = VEC - 1; i >= 0; --i) {
        const bool ok = (s0 + i < S);
        w[i] = ok ? blk.T[i] * (1.f - blk.eb[i]) : 0.f;
        G[i] = g0 * c[3 * i + 0] + g1 * c[3 * i + 1] + g2 * c[3 * i + 2] + h[i];
        x[i] = ok ? G[i] * w[i] : 0.f;
        suf[i] = run;  // exclusive suffix within the lane
        run += x[i];
      }
      float incl = warp_inclusive_rscan(run, lane);
      float excl = incl - run;
      float ds[VEC], dc[3 * VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        float R = rcarry + excl + suf[i];
        float db = -G[i] * blk.T[i] * blk.eb[i] + R;
        ds[i] = ((db * kappa_a) * kappa_b) * (-del[i]);
        dc[3 * i + 0] = w[i] * g0;
        dc[3 * i + 1] = w[i] * g1;
        dc[3 * i + 2] = w[i] * g2;
      }
      rcarry += __shfl_sync(0xffffffffu, incl, 0);
      store_vec<VEC>(d_sigma + ray * S, s0, S, ds);
      if (s0 < S) {
        float* out = d_rgb + ray * S * 3 + (size_t)s0 * 3;
        if (VEC == 4) {
#pragma unroll
          for (int q = 0; q < 3; ++q)
            st_stream4(out + q * 4, make_float4(dc[q * 4], dc[q * 4 + 1], dc[q * 4 + 2], dc[q * 4 + 3]));
        } else {
          out[0] = dc[0]; out[1] = dc[1]; out[2] = dc[2];
        }
      }
    }
  }
}

int grid_for(int B) {
  int blocks = ceil_div(B, kWarpsPerBlock);
  int cap = sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_composite_fwd(const float* sigma, const float* delta, const float* rgb,
                                      const float* t_mid, int B, int S, int flavour,
                                      float* out_rgb, float* out_w, float* out_opacity,
                                      float* out_depth, void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "composite_fwd: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(sigma && delta && rgb && out_rgb, "composite_fwd: null pointer");
  NB_CHECK_ARG(flavour == NERFB200_COMPOSITE_BARF || flavour == NERFB200_COMPOSITE_NERFACC,
               "composite_fwd: unknown flavour %d", flavour);
  NB_CHECK_ARG(out_depth == nullptr || t_mid != nullptr, "composite_fwd: depth needs t_mid");
  if (B == 0) return NERFB200_OK;
  const bool vec = (S % 4 == 0);
  NB_CHECK_ARG(S <= (vec ? 128 : 32) * kMaxBlocks, "composite_fwd: S=%d too large", S);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    composite_fwd_kernel<4><<<grid_for(B), kWarpsPerBlock * 32, 0, st>>>(
        sigma, delta, rgb, t_mid, B, S, flavour, out_rgb, out_w, out_opacity, out_depth);
  else
    composite_fwd_kernel<1><<<grid_for(B), kWarpsPerBlock * 32, 0, st>>>(
        sigma, delta, rgb, t_mid, B, S, flavour, out_rgb, out_w, out_opacity, out_depth);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_composite_bwd(const float* sigma, const float* delta, const float* rgb,
                                      const float* t_mid, const float* g_rgb, const float* g_w,
                                      const float* g_opacity, const float* g_depth, int B,
                                      int S, int flavour, float* d_sigma, float* d_rgb,
                                      void* stream) {
  NB_CHECK_ARG(B >= 0 && S >= 1, "composite_bwd: bad shape B=%d S=%d", B, S);
  NB_CHECK_ARG(sigma && delta && rgb && g_rgb && d_sigma && d_rgb, "composite_bwd: null pointer");
  NB_CHECK_ARG(flavour == NERFB200_COMPOSITE_BARF || flavour == NERFB200_COMPOSITE_NERFACC,
               "composite_bwd: unknown flavour %d", flavour);
  NB_CHECK_ARG(g_depth == nullptr || t_mid != nullptr, "composite_bwd: g_depth needs t_mid");
  if (B == 0) return NERFB200_OK;
  const bool vec = (S % 4 == 0);
  NB_CHECK_ARG(S <= (vec ? 128 : 32) * kMaxBlocks, "composite_bwd: S=%d too large", S);
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    composite_bwd_kernel<4><<<grid_for(B), kWarpsPerBlock * 32, 0, st>>>(
        sigma, delta, rgb, t_mid, g_rgb, g_w, g_opacity, g_depth, B, S, flavour, d_sigma, d_rgb);
  else
    composite_bwd_kernel<1><<<grid_for(B), kWarpsPerBlock * 32, 0, st>>>(
        sigma, delta, rgb, t_mid, g_rgb, g_w, g_opacity, g_depth, B, S, flavour, d_sigma, d_rgb);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
