// Learnable per-feature activations of the GARF / SARF / Gabor radiance networks, forward and
// backward (input gradient + parameter gradients reduced over the rows) as stand-alone
// HBM-bound kernels over (N, F) row-major fp32 activations:
//   GAUSS  y = exp(-x^2 v), v = p0^2 + 1e-6                reference barf/gaussian.py:8-63
//   SARF   x' = (signbit(x)*2-1)(|x|+1e-4);
//          y = cos(f / (x'^2 + 1/f^2)) exp(-x'^2), f = p0   reference sarf/activation.py:63-65
//          (the autograd path the reference actually runs; its custom Function is commented out)
//   GABOR  y = exp(-v x^2) cos(s x), v = p0^2 + 1e-6, s = p1 reference gaborf/gabor.py:8-64
// Thread = one feature column of a chunk of rows: loads are coalesced along the feature axis,
// the parameter gradient stays in a register until one atomicAdd per thread.
#include <cuda_bf16.h>
#include "common.cuh"

namespace nerfb200 {
namespace {

constexpr int kActThreads = 128;

struct ActParams {
  int kind;
  const float* x;
  const float* p0;
  const float* p1;
  const float* g;
  long long N;
  int F;
  float* y;
  float* dx;
  float* dp0;
  float* dp1;
  float* dsum;     // optional (F): column sums of dx = bias gradient of the Linear in front
  int out_bf16;    // y / dx are written as bf16 (the next GEMM's operand type) instead of fp32
  int rows_per_block;
};

__device__ __forceinline__ float act_forward(int kind, float x, float p0, float p1) {
  if (kind == NERFB200_ACT_GAUSS) {
    const float v = p0 * p0 + 1e-6f;
    return expf(-(x * x) * v);
  } else if (kind == NERFB200_ACT_SARF) {
    const float xa = fabsf(x) + 1e-4f;
    const float u = xa * xa;
    return cosf(p0 / (u + 1.f / (p0 * p0))) * expf(-u);
  } else {
    const float v = p0 * p0 + 1e-6f;
    return expf(-v * (x * x)) * cosf(p1 * x);
  }
}

__device__ __forceinline__ void store_out(float* base, long long i, float v, int out_bf16) {
  if (out_bf16) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else st_stream(base + i, v);
}

constexpr int kActUnroll = 8;   // rows in flight per thread: the kernels are bound by memory-level parallelism

__global__ void __launch_bounds__(kActThreads) act_fwd_kernel(const ActParams p) {
  const int f = blockIdx.y * kActThreads + threadIdx.x;
  if (f >= p.F) return;
  const float p0 = p.p0[f];
  const float p1 = p.p1 ? p.p1[f] : 0.f;
  const long long r0 = (long long)blockIdx.x * p.rows_per_block;
  const long long r1 = r0 + p.rows_per_block < p.N ? r0 + p.rows_per_block : p.N;
  long long r = r0;
  for (; r + kActUnroll <= r1; r += kActUnroll) {
    float x[kActUnroll];
#pragma unroll
    for (int u = 0; u < kActUnroll; ++u) x[u] = ld_stream(p.x + (r + u) * p.F + f);
#pragma unroll
    for (int u = 0; u < kActUnroll; ++u) store_out(p.y, (r + u) * p.F + f, act_forward(p.kind, x[u], p0, p1), p.out_bf16);
  }
  for (; r < r1; ++r) {
    const long long i = r * p.F + f;
    store_out(p.y, i, act_forward(p.kind, ld_stream(p.x + i), p0, p1), p.out_bf16);
  }
}

// dx and the parameter-gradient terms of one element
__device__ __forceinline__ float act_backward(int kind, float x, float g, float p0, float p1, float& a0, float& a1) {
  float dx;
  if (kind == NERFB200_ACT_GAUSS) {
    const float v = p0 * p0 + 1e-6f;
    const float x2 = x * x;
    const float ge = g * expf(-x2 * v);
    dx = -ge * 2.f * x * v;
    a0 += -ge * x2;                       // d/dv, chained to p0 by the caller
  } else if (kind == NERFB200_ACT_SARF) {
    const float xa = fabsf(x) + 1e-4f;
    const float u = xa * xa;
    const float inv_f2 = 1.f / (p0 * p0);
    const float D = u + inv_f2;
    const float a = p0 / D;
    float sn, cs;
    sincosf(a, &sn, &cs);
    const float e = expf(-u);
    // y = cos(a) e^-u, a = f / D, D = u + f^-2, u = (|x| + eps)^2
    const float dy_du = e * (sn * p0 / (D * D) - cs);
    const float sgn = x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f);   // torch.abs backward: sign(0) = 0
    dx = g * dy_du * 2.f * xa * sgn;
    const float da_df = 1.f / D + 2.f * inv_f2 / (D * D);
    a0 += g * (-sn * e * da_df);
  } else {
    const float v = p0 * p0 + 1e-6f;
    float sn, cs;
    sincosf(p1 * x, &sn, &cs);
    const float go = -expf(-v * x * x) * g;
    dx = go * (2.f * cs * v * x + p1 * sn);
    a0 += go * x * x * cs;                 // d/dv
    a1 += go * x * sn;                     // d/ds
  }
  return dx;
}

__global__ void __launch_bounds__(kActThreads) act_bwd_kernel(const ActParams p) {
  const int f = blockIdx.y * kActThreads + threadIdx.x;
  if (f >= p.F) return;
  const float p0 = p.p0[f];
  const float p1 = p.p1 ? p.p1[f] : 0.f;
  const long long r0 = (long long)blockIdx.x * p.rows_per_block;
  const long long r1 = r0 + p.rows_per_block < p.N ? r0 + p.rows_per_block : p.N;
  float a0 = 0.f, a1 = 0.f;   // parameter-gradient partial sums of this column
  float asum = 0.f;           // column sum of dx
  long long r = r0;
  for (; r + kActUnroll <= r1; r += kActUnroll) {
    float x[kActUnroll], g[kActUnroll];
#pragma unroll
    for (int u = 0; u < kActUnroll; ++u) {
      x[u] = ld_stream(p.x + (r + u) * p.F + f);
      g[u] = ld_stream(p.g + (r + u) * p.F + f);
    }
#pragma unroll
    for (int u = 0; u < kActUnroll; ++u) {   // same summation order as the row-by-row loop
      const float dx = act_backward(p.kind, x[u], g[u], p0, p1, a0, a1);
      store_out(p.dx, (r + u) * p.F + f, dx, p.out_bf16);
      asum += dx;
    }
  }
  for (; r < r1; ++r) {
    const long long i = r * p.F + f;
    const float dx = act_backward(p.kind, ld_stream(p.x + i), ld_stream(p.g + i), p0, p1, a0, a1);
    store_out(p.dx, i, dx, p.out_bf16);
    asum += dx;
  }
  if (p.kind != NERFB200_ACT_SARF) a0 *= 2.f * p0;   // v = p0^2 + 1e-6
  if (r1 > r0) {
    atomicAdd(p.dp0 + f, a0);
    if (p.kind == NERFB200_ACT_GABOR && p.dp1) atomicAdd(p.dp1 + f, a1);
    if (p.dsum) atomicAdd(p.dsum + f, asum);
  }
}

int launch_shape(long long N, int F, dim3& grid, int& rows_per_block) {
  const int col_blocks = ceil_div(F, kActThreads);
  // enough row chunks to fill the machine a few times over, at least 32 rows each
  long long chunks = (long long)sm_count() * 8 / col_blocks;
  if (chunks < 1) chunks = 1;
  long long rpb = (N + chunks - 1) / chunks;
  if (rpb < 32) rpb = 32;
  rows_per_block = (int)rpb;
  grid = dim3((unsigned)ceil_div(N, rpb), (unsigned)col_blocks);
  return NERFB200_OK;
}

}  // namespace
}  // namespace nerfb200

namespace nerfb200 {
namespace {
// Gradient of the Gaussian widths of a layer z = W x + b, y = exp(-z^2 (s^2 + 1e-6)) WITHOUT the
// pre-activations: sum over samples of z_n dz_n = sum_c W[n,c] dW[n,c] + b_n db_n (dW = dz^T x, db = sum dz),
// and dL/ds_n = (sum_samples z_n dz_n) s_n / (s_n^2 + 1e-6) (dz = -2 v z y g). One warp per output feature.
__global__ void __launch_bounds__(256) gauss_width_grad_kernel(const NbGaussLayer* layers, int n_layers,
                                                               const float* params, float* d_params, float sign) {
  const int lane = threadIdx.x & 31;
  long long feat = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int l = 0;
  for (; l < n_layers; ++l) {
    if (feat < layers[l].out_f) break;
    feat -= layers[l].out_f;
  }
  if (l >= n_layers) return;
  const NbGaussLayer L = layers[l];
  const float* W = params + L.w_off + feat * L.in_f;
  const float* dW = d_params + L.w_off + feat * L.in_f;
  float acc = 0.f;
  for (int c = lane; c < L.in_f; c += 32) acc = fmaf(__ldg(W + c), dW[c], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    acc = fmaf(params[L.b_off + feat], d_params[L.b_off + feat], acc);
    const float sdev = params[L.g_off + feat];
    d_params[L.g_off + feat] += sign * acc * sdev / (sdev * sdev + 1e-6f);
  }
}
}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_gauss_width_grad(const NbGaussLayer* layers_dev, int n_layers, long long n_features,
                                         const float* params, float* d_params, float sign, void* stream) {
  NB_CHECK_ARG(n_layers >= 0 && n_features >= 0 && (n_layers == 0 || (layers_dev && params && d_params)),
               "gauss_width_grad: bad arguments");
  if (n_layers == 0 || n_features == 0) return NERFB200_OK;
  const int warps = 8;
  const long long blocks = (n_features + warps - 1) / warps;
  gauss_width_grad_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(layers_dev, n_layers, params,
                                                                                     d_params, sign);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}


extern "C" int nerfb200_act_fwd(int kind, const float* x, const float* p0, const float* p1,
                                long long N, int F, void* y, int out_bf16, void* stream) {
  NB_CHECK_ARG(kind >= NERFB200_ACT_GAUSS && kind <= NERFB200_ACT_GABOR, "act_fwd: unknown kind %d", kind);
  NB_CHECK_ARG(N >= 0 && F >= 1, "act_fwd: bad shape N=%lld F=%d", N, F);
  NB_CHECK_ARG(kind != NERFB200_ACT_GABOR || p1, "act_fwd: the Gabor activation needs its spread parameter");
  if (N == 0) return NERFB200_OK;
  NB_CHECK_ARG(x && p0 && y, "act_fwd: null pointer");
  ActParams p{};
  p.kind = kind; p.x = x; p.p0 = p0; p.p1 = p1; p.N = N; p.F = F; p.y = reinterpret_cast<float*>(y);
  p.out_bf16 = out_bf16;
  dim3 grid;
  launch_shape(N, F, grid, p.rows_per_block);
  act_fwd_kernel<<<grid, kActThreads, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_act_bwd(int kind, const float* x, const float* p0, const float* p1,
                                const float* g, long long N, int F, void* dx, float* dp0,
                                float* dp1, float* dsum, int out_bf16, void* stream) {
  NB_CHECK_ARG(kind >= NERFB200_ACT_GAUSS && kind <= NERFB200_ACT_GABOR, "act_bwd: unknown kind %d", kind);
  NB_CHECK_ARG(N >= 0 && F >= 1, "act_bwd: bad shape N=%lld F=%d", N, F);
  NB_CHECK_ARG(kind != NERFB200_ACT_GABOR || (p1 && dp1), "act_bwd: the Gabor activation needs p1 and dp1");
  if (N == 0) return NERFB200_OK;
  NB_CHECK_ARG(x && p0 && g && dx && dp0, "act_bwd: null pointer");
  ActParams p{};
  p.kind = kind; p.x = x; p.p0 = p0; p.p1 = p1; p.g = g; p.N = N; p.F = F; p.dx = reinterpret_cast<float*>(dx); p.dp0 = dp0; p.dp1 = dp1; p.dsum = dsum; p.out_bf16 = out_bf16;
  dim3 grid;
  launch_shape(N, F, grid, p.rows_per_block);
  act_bwd_kernel<<<grid, kActThreads, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
