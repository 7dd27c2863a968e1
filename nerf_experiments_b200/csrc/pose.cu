// Camera extrinsics: per-image so(3) rotation about the camera centre + additive translation.
// Mirrors CameraExtrinsics (reference barf/model_camera_extrinsics.py:7-85). The reference
// evaluates matrix_exp for ALL images every step and gathers; here every ray evaluates the
// closed form (Rodrigues) of its own image — 3 floats of parameters per ray from L2.
#include "common.cuh"

namespace nerfb200 {

// exp([w]x) coefficients: R = I + A*K + Bc*K^2, right Jacobian J_r = I - Bc*K + Cc*K^2.
struct So3Coef {
  float A, Bc, Cc;
};

__device__ __forceinline__ So3Coef so3_coef(float wx, float wy, float wz) {
  const float t2 = wx * wx + wy * wy + wz * wz;
  So3Coef c;
  if (t2 < 1e-4f) {  // Taylor: error < 1e-10
    c.A = 1.f - t2 * (1.f / 6.f) + t2 * t2 * (1.f / 120.f);
    c.Bc = 0.5f - t2 * (1.f / 24.f) + t2 * t2 * (1.f / 720.f);
    c.Cc = (1.f / 6.f) - t2 * (1.f / 120.f) + t2 * t2 * (1.f / 5040.f);
  } else {
    const float t = sqrtf(t2);
    float s, co;
    sincosf(t, &s, &co);
    c.A = s / t;
    c.Bc = (1.f - co) / t2;
    c.Cc = (t - s) / (t2 * t);
  }
  return c;
}

// R (row-major 3x3) = I + A*K + Bc*K^2 with K = [w]x
__device__ __forceinline__ void so3_exp(float wx, float wy, float wz, float (&R)[9]) {
  const So3Coef c = so3_coef(wx, wy, wz);
  const float xx = wx * wx, yy = wy * wy, zz = wz * wz;
  const float xy = wx * wy, xz = wx * wz, yz = wy * wz;
  R[0] = 1.f - c.Bc * (yy + zz);
  R[1] = -c.A * wz + c.Bc * xy;
  R[2] = c.A * wy + c.Bc * xz;
  R[3] = c.A * wz + c.Bc * xy;
  R[4] = 1.f - c.Bc * (xx + zz);
  R[5] = -c.A * wx + c.Bc * yz;
  R[6] = -c.A * wy + c.Bc * xz;
  R[7] = c.A * wx + c.Bc * yz;
  R[8] = 1.f - c.Bc * (xx + yy);
}

namespace {

__global__ void __launch_bounds__(256)
pose_fwd_kernel(const float* __restrict__ rotation, const float* __restrict__ translation,
                const int32_t* __restrict__ img_idx, const float* __restrict__ o,
                const float* __restrict__ d, int B, int n_images, float* __restrict__ out_o,
                float* __restrict__ out_d, float* __restrict__ out_R, float* __restrict__ out_t) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const int i = img_idx[r];
    if ((unsigned)i >= (unsigned)n_images) {
      // the reference raises IndexError on the host; a stream-ordered call cannot, so the ray is
      // poisoned instead: NaN outputs make the loss NaN (visible, and the guarded optimiser skips
      // the step) and nothing is read or written out of bounds
      const float q = __int_as_float(0x7fc00000);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        out_o[3 * r + c] = q;
        out_d[3 * r + c] = q;
        if (out_t != nullptr) out_t[3 * r + c] = q;
      }
      if (out_R != nullptr)
        for (int c = 0; c < 9; ++c) out_R[9 * (size_t)r + c] = q;
      continue;
    }
    const float wx = __ldg(rotation + 3 * i), wy = __ldg(rotation + 3 * i + 1),
                wz = __ldg(rotation + 3 * i + 2);
    const float tx = __ldg(translation + 3 * i), ty = __ldg(translation + 3 * i + 1),
                tz = __ldg(translation + 3 * i + 2);
    float R[9];
    so3_exp(wx, wy, wz, R);
    const float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];
    out_o[3 * r + 0] = o[3 * r + 0] + tx;  // translation / MAGIC_NUMBER_THE_SECOND (=1)
    out_o[3 * r + 1] = o[3 * r + 1] + ty;
    out_o[3 * r + 2] = o[3 * r + 2] + tz;
    out_d[3 * r + 0] = R[0] * dx + R[1] * dy + R[2] * dz;
    out_d[3 * r + 1] = R[3] * dx + R[4] * dy + R[5] * dz;
    out_d[3 * r + 2] = R[6] * dx + R[7] * dy + R[8] * dz;
    if (out_R != nullptr) {
#pragma unroll
      for (int q = 0; q < 9; ++q) out_R[9 * (size_t)r + q] = R[q];
    }
    if (out_t != nullptr) {
      out_t[3 * r + 0] = tx;
      out_t[3 * r + 1] = ty;
      out_t[3 * r + 2] = tz;
    }
  }
}

// dL/dw_i = J_r(w_i)^T * sum_rays ( d x (R_i^T g_d) );  dL/dt_i = sum_rays g_o.
__global__ void __launch_bounds__(256)
pose_bwd_kernel(const float* __restrict__ rotation, const int32_t* __restrict__ img_idx,
                const float* __restrict__ d, const float* __restrict__ g_o,
                const float* __restrict__ g_d, int B, int n_images, float* __restrict__ d_rotation,
                float* __restrict__ d_translation) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x) {
    const int i = img_idx[r];
    if ((unsigned)i >= (unsigned)n_images) continue;   // poisoned in the forward pass; never scatter out of bounds
    const float wx = __ldg(rotation + 3 * i), wy = __ldg(rotation + 3 * i + 1),
                wz = __ldg(rotation + 3 * i + 2);
    float R[9];
    so3_exp(wx, wy, wz, R);
    const So3Coef c = so3_coef(wx, wy, wz);
    const float gx = g_d[3 * r], gy = g_d[3 * r + 1], gz = g_d[3 * r + 2];
    // a = R^T g
    const float ax = R[0] * gx + R[3] * gy + R[6] * gz;
    const float ay = R[1] * gx + R[4] * gy + R[7] * gz;
    const float az = R[2] * gx + R[5] * gy + R[8] * gz;
    const float dx = d[3 * r], dy = d[3 * r + 1], dz = d[3 * r + 2];
    // b = d x a
    const float bx = dy * az - dz * ay;
    const float by = dz * ax - dx * az;
    const float bz = dx * ay - dy * ax;
    // J_r^T b = b + Bc (w x b) + Cc (w x (w x b))
    const float kx = wy * bz - wz * by, ky = wz * bx - wx * bz, kz = wx * by - wy * bx;
    const float lx = wy * kz - wz * ky, ly = wz * kx - wx * kz, lz = wx * ky - wy * kx;
    atomicAdd(d_rotation + 3 * i + 0, bx + c.Bc * kx + c.Cc * lx);
    atomicAdd(d_rotation + 3 * i + 1, by + c.Bc * ky + c.Cc * ly);
    atomicAdd(d_rotation + 3 * i + 2, bz + c.Bc * kz + c.Cc * lz);
    atomicAdd(d_translation + 3 * i + 0, g_o[3 * r + 0]);
    atomicAdd(d_translation + 3 * i + 1, g_o[3 * r + 1]);
    atomicAdd(d_translation + 3 * i + 2, g_o[3 * r + 2]);
  }
}

__global__ void so3_kernel(const float* __restrict__ so3, int n, float* __restrict__ out_R) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float R[9];
    so3_exp(so3[3 * i], so3[3 * i + 1], so3[3 * i + 2], R);
#pragma unroll
    for (int q = 0; q < 9; ++q) out_R[9 * (size_t)i + q] = R[q];
  }
}

int grid1d(int n) {
  int b = ceil_div(n, 256);
  const int cap = sm_count() * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : b;
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_pose_fwd(const float* rotation, const float* translation,
                                 const int32_t* img_idx, const float* o, const float* d, int B,
                                 int n_images, float* out_o, float* out_d, float* out_R,
                                 float* out_t, void* stream) {
  NB_CHECK_ARG(B >= 0 && n_images >= 1, "pose_fwd: bad shape B=%d n_images=%d", B, n_images);
  NB_CHECK_ARG(rotation && translation && img_idx && o && d && out_o && out_d, "pose_fwd: null pointer");
  if (B == 0) return NERFB200_OK;
  pose_fwd_kernel<<<grid1d(B), 256, 0, (cudaStream_t)stream>>>(rotation, translation, img_idx, o, d,
                                                               B, n_images, out_o, out_d, out_R, out_t);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_pose_bwd(const float* rotation, const int32_t* img_idx, const float* d,
                                 const float* g_o, const float* g_d, int B, int n_images,
                                 float* d_rotation, float* d_translation, void* stream) {
  NB_CHECK_ARG(B >= 0 && n_images >= 1, "pose_bwd: bad shape B=%d n_images=%d", B, n_images);
  NB_CHECK_ARG(rotation && img_idx && d && g_o && g_d && d_rotation && d_translation,
               "pose_bwd: null pointer");
  if (B == 0) return NERFB200_OK;
  pose_bwd_kernel<<<grid1d(B), 256, 0, (cudaStream_t)stream>>>(rotation, img_idx, d, g_o, g_d, B,
                                                               n_images, d_rotation, d_translation);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_so3_to_SO3(const float* so3, int n, float* out_R, void* stream) {
  NB_CHECK_ARG(n >= 0 && (n == 0 || (so3 && out_R)), "so3_to_SO3: bad arguments");
  if (n == 0) return NERFB200_OK;
  so3_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(so3, n, out_R);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
