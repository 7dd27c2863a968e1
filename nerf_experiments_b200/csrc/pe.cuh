// Positional encodings evaluated in registers by the row-threads of the fused MLP kernels.
// Mirrors reference barf/positional_encodings.py: FourierFeatures (:28-57),
// BarfPositionalEncoding (:61-148), IntegratedFourierFeatures (:151-240),
// IntegratedBarfFourierFeatures (:242-282) and IdentityPositionalEncoding (:17-25).
//
// Column order of the encoding:  [x y z] (if include_identity) | cos block | sin block, each
// block ordered coordinate-major: (x, level 0..L-1), (y, ...), (z, ...).
#pragma once
#include "mlp.h"

namespace nerfb200 {

constexpr int kMaxLevels = 16;
constexpr int kPeAnchor = 4;   // levels between exact sincos evaluations (error <= ~2e-6)

// BARF coarse-to-fine mask (positional_encodings.py:105-122) from the alpha buffer.
__device__ __forceinline__ void pe_fill_mask(const NbPeCfg& cfg, const float* alpha_ptr,
                                             float* mask /* [kMaxLevels] */) {
  const float alpha = (cfg.use_mask && alpha_ptr != nullptr) ? *alpha_ptr : (float)cfg.levels;
  const int ramp = (int)alpha;  // int(alpha): truncation
  for (int j = 0; j < kMaxLevels; ++j) {
    float m = 0.f;
    if (!cfg.use_mask) m = 1.f;
    else if (j < ramp) m = 1.f;
    else if (j == ramp) m = (1.f - cosf((alpha - (float)ramp) * 3.14159274101257324f)) * 0.5f;
    mask[j] = (j < cfg.levels) ? m : 0.f;
  }
}

struct PeSample {
  float x[3];        // query position (the encoder may shift it: integrated PE mean)
  float dir[3];      // ray direction
  float pixel_width; // integrated PE only
  float t0, t1;      // integrated PE only
};

struct PeIpe {
  float mu_diff;     // shift of the mean along the ray          (eq. 8)
  float var[3];      // per-coordinate level-0 variance           (eq. 16 or its isotropic mean)
};

// Mip-NeRF conical-frustum Gaussian (positional_encodings.py:186-226)
__device__ __forceinline__ PeIpe pe_ipe_prepare(const NbPeCfg& cfg, const PeSample& s) {
  PeIpe r;
  const float t_mu = (s.t0 + s.t1) * 0.5f;
  const float t_d = (s.t1 - s.t0) * 0.5f;
  const float td2 = t_d * t_d;
  const float den = 3.f * t_mu * t_mu + td2;
  r.mu_diff = 2.f * t_mu * td2 / den;
  const float r_dot = s.pixel_width * 2.f / 3.46410161513775459f;  // 12**0.5
  float st2 = td2 / 3.f - (4.f * td2 * td2 * (12.f * t_mu * t_mu - td2)) / (15.f * den * den);
  float sr2 = r_dot * r_dot * (t_mu * t_mu * 0.25f + 5.f * td2 / 12.f - 4.f * td2 * td2 / (15.f * den));
  if (cfg.pixel_width_sigma > 0.25f) {
    const float a = cfg.pixel_width_sigma * s.pixel_width * t_mu;
    st2 += a * a;
    sr2 += a * a;
  }
  if (cfg.distribute_variance) {
    const float v = (st2 + sr2 * 2.f) / 3.f;
    r.var[0] = r.var[1] = r.var[2] = v;
  } else {
    const float dn = s.dir[0] * s.dir[0] + s.dir[1] * s.dir[1] + s.dir[2] * s.dir[2];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d2 = s.dir[c] * s.dir[c];
      r.var[c] = st2 * d2 + sr2 * (1.f - d2 / dn);
    }
  }
  return r;
}

// Calls emit(column, value) for the columns of the encoding of one sample selected by `parts`:
// bit c (c = 0..2) = the cos / sin columns of coordinate c, bit 3 = the identity columns. The
// fused kernels split a row over four threads that way (kPeAll = everything); the value of a
// column does not depend on who computes it.
constexpr int kPeAll = 15;
template <typename Emit>
__device__ __forceinline__ void pe_encode(const NbPeCfg& cfg, const float* mask,
                                          const PeSample& s, Emit emit, int parts = kPeAll) {
  float x[3] = {s.x[0], s.x[1], s.x[2]};
  PeIpe ipe;
  if (cfg.kind == NB_PE_INTEGRATED) {
    ipe = pe_ipe_prepare(cfg, s);
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = x[c] + ipe.mu_diff * s.dir[c];
  }
  int col = 0;
  if (cfg.include_identity || cfg.kind == NB_PE_IDENTITY) {
    if (parts & 8) { emit(0, x[0]); emit(1, x[1]); emit(2, x[2]); }
    col = 3;
  }
  if (cfg.kind == NB_PE_IDENTITY) return;
  const int L = cfg.levels;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (!((parts >> c) & 1)) continue;
    float sn = 0.f, cs = 1.f;
    float lvl = 1.f;  // 4^j
    float freq = cfg.scale;
    for (int j = 0; j < L; ++j) {
      // exact evaluation every kPeAnchor levels, double-angle steps in between
      if ((j % kPeAnchor) == 0) sincosf(x[c] * freq, &sn, &cs);
      float w = mask[j];
      if (cfg.kind == NB_PE_INTEGRATED) w *= __expf(-0.5f * ipe.var[c] * lvl);
      emit(col + c * L + j, w * cs);
      emit(col + 3 * L + c * L + j, w * sn);
      // double-angle step to level j+1
      const float s2 = 2.f * sn * cs;
      cs = 1.f - 2.f * sn * sn;
      sn = s2;
      lvl *= 4.f;
      freq *= 2.f;
    }
  }
}

// Gradient of the encoding w.r.t. the query position x, given g[column] = dL/d(encoding).
// For the integrated PE the (second-order) dependence of the damping weights on the ray
// direction is not differentiated; the mean shift contributes mu_diff * dL/dx to dL/d(dir).
template <typename Load>
__device__ __forceinline__ void pe_backward(const NbPeCfg& cfg, const float* mask,
                                            const PeSample& s, Load g, float (&dx)[3],
                                            float& dir_scale) {
  float x[3] = {s.x[0], s.x[1], s.x[2]};
  PeIpe ipe;
  dir_scale = 0.f;
  if (cfg.kind == NB_PE_INTEGRATED) {
    ipe = pe_ipe_prepare(cfg, s);
    dir_scale = ipe.mu_diff;
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = x[c] + ipe.mu_diff * s.dir[c];
  }
  dx[0] = dx[1] = dx[2] = 0.f;
  int col = 0;
  if (cfg.include_identity || cfg.kind == NB_PE_IDENTITY) {
    dx[0] = g(0); dx[1] = g(1); dx[2] = g(2);
    col = 3;
  }
  if (cfg.kind == NB_PE_IDENTITY) return;
  const int L = cfg.levels;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float sn = 0.f, cs = 1.f;
    float lvl = 1.f, freq = cfg.scale;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) {
      if ((j % kPeAnchor) == 0) sincosf(x[c] * freq, &sn, &cs);
      float w = mask[j];
      if (cfg.kind == NB_PE_INTEGRATED) w *= __expf(-0.5f * ipe.var[c] * lvl);
      acc += w * freq * (cs * g(col + 3 * L + c * L + j) - sn * g(col + c * L + j));
      const float s2 = 2.f * sn * cs;
      cs = 1.f - 2.f * sn * sn;
      sn = s2;
      lvl *= 4.f;
      freq *= 2.f;
    }
    dx[c] += acc;
  }
}

// As pe_backward, for the 16-column quarter CQ of an accumulator in the canonical column order
// (NB_PE_CANON_*, mlp.h) held in registers: every column index is a compile-time constant,
// levels beyond cfg.levels are predicated off. The four quarters of a row add up to the full
// gradient; the mean-shift scale (integrated PE) is returned by every quarter.
template <int CQ>
__device__ __forceinline__ void pe_backward_canon_quarter(const NbPeCfg& cfg, const float* mask,
                                                          const PeSample& s, const uint32_t (&g)[16],
                                                          float (&dx)[3], float& dir_scale) {
  float x[3] = {s.x[0], s.x[1], s.x[2]};
  PeIpe ipe;
  dir_scale = 0.f;
  if (cfg.kind == NB_PE_INTEGRATED) {
    ipe = pe_ipe_prepare(cfg, s);
    dir_scale = ipe.mu_diff;
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = x[c] + ipe.mu_diff * s.dir[c];
  }
  dx[0] = dx[1] = dx[2] = 0.f;
  if (CQ == NB_PE_CANON_IDENTITY / 16 && (cfg.include_identity || cfg.kind == NB_PE_IDENTITY)) {
#pragma unroll
    for (int c = 0; c < 3; ++c) dx[c] = __uint_as_float(g[NB_PE_CANON_IDENTITY % 16 + c]);
  }
  if (cfg.kind == NB_PE_IDENTITY) return;
  const int L = cfg.levels;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float sn = 0.f, cs = 1.f;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < NB_PE_CANON_LEVELS; ++j) {
      constexpr int kNone = -1;
      const int col = 2 * (NB_PE_CANON_LEVELS * c + j);
      const bool mine = (col / 16 == CQ);
      // exact evaluation at the anchor level of a group this quarter needs, double-angle steps
      // inside the group (as pe_encode)
      const int anchor = (j / kPeAnchor) * kPeAnchor;
      bool group_needed = false;
#pragma unroll
      for (int jj = 0; jj < kPeAnchor; ++jj)
        if (anchor + jj < NB_PE_CANON_LEVELS && (2 * (NB_PE_CANON_LEVELS * c + anchor + jj)) / 16 == CQ) group_needed = true;
      (void)kNone;
      if (group_needed && j < L) {
        const float freq = cfg.scale * (float)(1 << j);
        if (j == anchor) sincosf(x[c] * freq, &sn, &cs);
        if (mine) {
          float w = mask[j];
          if (cfg.kind == NB_PE_INTEGRATED) w *= __expf(-0.5f * ipe.var[c] * (float)(1 << (2 * j)));
          const float g_cos = __uint_as_float(g[col % 16]);
          const float g_sin = __uint_as_float(g[col % 16 + 1]);
          acc += w * freq * (cs * g_sin - sn * g_cos);
        }
        const float s2 = 2.f * sn * cs;
        cs = 1.f - 2.f * sn * sn;
        sn = s2;
      }
    }
    dx[c] += acc;
  }
}

}  // namespace nerfb200
