// Pose alignment on the device: the similarity transform (R, t, c) that
// maps one set of camera origins onto another (Kabsch / Umeyama with the reference's scale
// definition), with the reference's outlier rejection (drop the points beyond the 0.9 distance
// quantile of a first fit, fit again), and the mean alignment error. Replaces
// CameraCalibrationModel.kabsch_algorithm and compute_pose_error (reference
// barf/model_camera_calibration.py:69-156, :340-346) — ~40 small torch launches, a cuSOLVER
// SVD and two host synchronisations (th.quantile + boolean-mask indexing) per call, and BARF
// calls it every training step (barf/model_barf.py:67). One block, no host sync.
#include "common.cuh"

namespace nerfb200 {
namespace {

constexpr int kKabschThreads = 256;
constexpr int kKabschMaxPoints = 2048;

__device__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kKabschThreads / 32; ++w) t += scratch[w];
  return t;
}

// Jacobi eigen-decomposition of a symmetric 3x3 matrix: A = V diag(w) V^T (thread 0)
__device__ void eig_sym3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double tq = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(tq * tq + 1.0), sn = tq * cs;
        for (int k = 0; k < 3; ++k) {   // A <- A J
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = cs * akp - sn * akq;
          A[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; ++k) {   // A <- J^T A
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = cs * apk - sn * aqk;
          A[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = A[i][i];
}

// R = argmin ||P R^T - Q||: with H = P^T Q = U S V^T,  R = V diag(1,1,det(V U^T)) U^T (thread 0)
__device__ void rotation_from_covariance(const double H[3][3], double R[3][3]) {
  double HtH[3][3], V[3][3], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += H[k][i] * H[k][j];
      HtH[i][j] = s;
    }
  eig_sym3(HtH, V, w);
  // sort the singular directions by decreasing singular value
  int order[3] = {0, 1, 2};
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (w[order[b]] > w[order[a]]) { const int t = order[a]; order[a] = order[b]; order[b] = t; }
  double Vs[3][3], U[3][3];
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < 3; ++i) Vs[i][j] = V[i][order[j]];
  // U columns: u_j = H v_j / s_j for the two dominant directions, the third by a cross product
  for (int j = 0; j < 2; ++j) {
    double u[3], n = 0.0;
    for (int i = 0; i < 3; ++i) {
      u[i] = H[i][0] * Vs[0][j] + H[i][1] * Vs[1][j] + H[i][2] * Vs[2][j];
      n += u[i] * u[i];
    }
    n = sqrt(n);
    for (int i = 0; i < 3; ++i) U[i][j] = n > 0 ? u[i] / n : (i == j ? 1.0 : 0.0);
  }
  U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
  U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
  U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  // make (v_1, v_2, v_3) right handed as well, then the sign of det(V U^T) is carried by K
  double Vc[3] = {Vs[1][0] * Vs[2][1] - Vs[2][0] * Vs[1][1], Vs[2][0] * Vs[0][1] - Vs[0][0] * Vs[2][1],
                  Vs[0][0] * Vs[1][1] - Vs[1][0] * Vs[0][1]};
  // the true third right-singular vector is +-Vc; its sign decides det(V U^T):
  // H v_3 . u_3 >= 0 for a proper SVD with u_3 := U[:,2]
  double hv[3], dot = 0.0;
  for (int i = 0; i < 3; ++i) {
    hv[i] = H[i][0] * Vc[0] + H[i][1] * Vc[1] + H[i][2] * Vc[2];
    dot += hv[i] * U[i][2];
  }
  const double sgn = dot >= 0 ? 1.0 : -1.0;   // = det(V U^T) of the SVD with non-negative singular values
  for (int i = 0; i < 3; ++i) Vs[i][2] = Vc[i] * sgn;
  // R = V K U^T with K = diag(1, 1, det(V U^T)); V, U as constructed: det(U) = 1, det(V) = sgn
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[i][j] = Vs[i][0] * U[j][0] + Vs[i][1] * U[j][1] + sgn * Vs[i][2] * U[j][2];
}

struct Fit { double R[3][3], t[3], c; };

// one weighted fit; every thread returns the same result
__device__ void fit(const float* from, const float* to, const float* weight, int n, double* scratch, Fit* out_sh) {
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += kKabschThreads) {
    const double w = weight ? (double)weight[i] : 1.0;
    for (int k = 0; k < 3; ++k) { acc[k] += w * from[i * 3 + k]; acc[3 + k] += w * to[i * 3 + k]; }
    acc[6] += w;
  }
  double s[7];
  for (int k = 0; k < 7; ++k) s[k] = block_sum(acc[k], scratch);
  const double cnt = s[6] > 0 ? s[6] : 1.0;
  const double mf[3] = {s[0] / cnt, s[1] / cnt, s[2] / cnt}, mt[3] = {s[3] / cnt, s[4] / cnt, s[5] / cnt};
  double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, nf = 0.0, nt = 0.0;
  for (int i = threadIdx.x; i < n; i += kKabschThreads) {
    const double w = weight ? (double)weight[i] : 1.0;
    double p[3], q[3];
    for (int k = 0; k < 3; ++k) { p[k] = from[i * 3 + k] - mf[k]; q[k] = to[i * 3 + k] - mt[k]; }
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) h[a * 3 + b] += w * p[a] * q[b];
    nf += w * (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    nt += w * (q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  }
  double H[3][3];
  for (int k = 0; k < 9; ++k) H[k / 3][k % 3] = block_sum(h[k], scratch);
  nf = block_sum(nf, scratch);
  nt = block_sum(nt, scratch);
  if (threadIdx.x == 0) {
    out_sh->c = sqrt(nt) / sqrt(nf);                       // (:127)
    rotation_from_covariance(H, out_sh->R);                // (:105-114)
    for (int i = 0; i < 3; ++i)                            // t = mean_to - c R mean_from  (:134)
      out_sh->t[i] = mt[i] - out_sh->c * (out_sh->R[i][0] * mf[0] + out_sh->R[i][1] * mf[1] + out_sh->R[i][2] * mf[2]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kKabschThreads)
kabsch_kernel(const float* __restrict__ from, const float* __restrict__ to, int n, int remove_outliers,
              float* __restrict__ out_R, float* __restrict__ out_t, float* __restrict__ out_c,
              float* __restrict__ out_err) {
  __shared__ double scratch[kKabschThreads / 32];
  __shared__ Fit f;
  __shared__ float dist[kKabschMaxPoints];
  __shared__ float sorted[kKabschMaxPoints];
  __shared__ float keep[kKabschMaxPoints];
  fit(from, to, nullptr, n, scratch, &f);
  auto distance_of = [&](int i) {
    double d2 = 0.0;
    for (int r = 0; r < 3; ++r) {
      const double v = (f.R[r][0] * from[i * 3] + f.R[r][1] * from[i * 3 + 1] + f.R[r][2] * from[i * 3 + 2]) * f.c + f.t[r] - to[i * 3 + r];
      d2 += v * v;
    }
    return sqrt(d2);
  };
  if (remove_outliers) {
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += kKabschThreads) {
      const float d = i < n ? (float)distance_of(i) : __int_as_float(0x7f800000);
      if (i < n) dist[i] = d;
      sorted[i] = d;
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)            // bitonic sort, ascending
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < m; i += kKabschThreads) {
          const int l = i ^ j;
          if (l > i) {
            const bool up = (i & k) == 0;
            const float a = sorted[i], b = sorted[l];
            if ((a > b) == up) { sorted[i] = b; sorted[l] = a; }
          }
        }
        __syncthreads();
      }
    // th.quantile(distances, 0.9), linear interpolation between order statistics (:146)
    const float pos = 0.9f * (float)(n - 1);
    const int lo = (int)floorf(pos), hi = min(lo + 1, n - 1);
    const float q = sorted[lo] + (pos - (float)lo) * (sorted[hi] - sorted[lo]);
    for (int i = threadIdx.x; i < n; i += kKabschThreads) keep[i] = dist[i] < q ? 1.f : 0.f;   // (:148)
    __syncthreads();
    fit(from, to, keep, n, scratch, &f);
  }
  // mean alignment error over ALL points (compute_pose_error, :343-345)
  double e = 0.0;
  for (int i = threadIdx.x; i < n; i += kKabschThreads) e += distance_of(i);
  e = block_sum(e, scratch);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) out_R[i * 3 + j] = (float)f.R[i][j];
      out_t[i] = (float)f.t[i];
    }
    out_c[0] = (float)f.c;
    if (out_err) out_err[0] = (float)(e / n);
  }
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_kabsch(const float* from, const float* to, int n, int remove_outliers,
                               float* out_R, float* out_t, float* out_c, float* out_err, void* stream) {
  NB_CHECK_ARG(n >= 1 && n <= kKabschMaxPoints, "kabsch: 1 <= n <= %d points (got %d)", kKabschMaxPoints, n);
  NB_CHECK_ARG(from && to && out_R && out_t && out_c, "kabsch: null pointer");
  kabsch_kernel<<<1, kKabschThreads, 0, (cudaStream_t)stream>>>(from, to, n, remove_outliers, out_R, out_t, out_c, out_err);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
