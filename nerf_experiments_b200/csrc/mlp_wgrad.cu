// Weight gradients of the fused MLP on sm_100a: dW_l = dY_l^T X_l as tcgen05 MMAs whose K
// dimension is the sample index. Both operands are the bf16 slabs the forward (X) and backward
// (dY) kernels left in HBM; because a slab row is one sample, the very same bytes are valid
// MN-major UMMA operands (tc.cuh), so they stream in by cp.async.bulk with no transposition.
//
// Work item = (layer, input source, <=2 blocks of 128 output features) x a range of tiles.
// A persistent CTA accumulates its range in TMEM (2 x 256 fp32 columns) and flushes once with
// red.global.add into the flat fp32 gradient buffer. HBM-bound by construction
// (128 FLOP per byte of stash read).
//
// Bias gradients ride along: the four flush warps are idle while an item accumulates, so each
// sums the columns of one dY half slab per pipeline stage while it sits in shared memory (the
// item that carries bias_dst >= 0 for its dY slabs), fp32 in registers over the whole item.
// GARF's Gaussian-width gradients (sum over samples of z * dz) ride the same way: an item with
// "z duty" also streams the z slabs of some of its dY slabs into the stage (a stage holds up to
// nine half slabs: dY | X | Z), so the dz slabs are read from HBM once for the weight block AND
// the column sums; column-sum-only items (no weight block) take what does not fit. (The shipped GARF
// programs no longer use the z duty: for y = exp(-z^2 v) the sum z dz follows from dW and db,
// nerfb200_gauss_width_grad. It stays for activations whose parameter gradients do not reduce to those,
// and is exercised through the C ABI by tests/test_gpu_wgrad_items.py.)
#include "common.cuh"
#include "mlp.h"
#include "tc.cuh"

namespace nerfb200 {
namespace {

using namespace tc;

constexpr int kWgThreads = 192;
constexpr int kWgStages = 3;
constexpr int kHalfRows = 64;                         // samples per pipeline stage
constexpr uint32_t kHalfSlabBytes = kHalfRows * 128;  // 8 KB
constexpr int kStageHalfSlabs = 9;                    // dY (2 per 128-feature block) | X | Z, packed in that order
constexpr uint32_t kStageBytes = kStageHalfSlabs * kHalfSlabBytes;
constexpr uint32_t kWgSmemBytes = kWgStages * kStageBytes + 256;

struct WgradParams {
  const NbWgradItem* items;
  int n_items;
  int* counter;                 // dynamic work distribution (zeroed by the host wrapper)
  const uint8_t* x_stash;
  const uint8_t* dy_stash;
  const uint8_t* z_stash;       // pre-activation stash of the GARF networks (column-sum items), may be NULL
  int x_slabs_per_tile, dy_slabs_per_tile, z_slabs_per_tile;
  float* d_params;
  const float* params;          // fp32 master parameters (Gaussian widths for their chain rule)
};

__global__ void __launch_bounds__(kWgThreads, 1)
mlp_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kWgStages * kStageBytes);
  uint64_t* empty = full + kWgStages;
  uint64_t* acc_full = empty + kWgStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);
  int* next_item = reinterpret_cast<int*>(tmem_ptr + 1);   // [2] double-buffered broadcast slot
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1 + 4);   // MMA commit + the four column-sum warps
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Static round-robin over items sorted by cost (host side): item = blockIdx.x + i*gridDim.x.
  if (warp == 5) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const NbWgradItem item = p.items[it];
        const int n_x = item.mode == NB_WGRAD_COLSUM ? 0 : item.n_x_slabs;
        const int n_z = item.z_slab >= 0 ? item.n_z_slabs : 0;
        const int pos_x = 2 * ((item.n_dy_slabs + 1) / 2), pos_z = pos_x + n_x;
        const uint32_t bytes = (uint32_t)(item.n_dy_slabs + n_x + n_z) * kHalfSlabBytes;
        for (int tile = item.tile_begin; tile < item.tile_end; ++tile) {
          const uint8_t* dy = p.dy_stash + ((size_t)tile * p.dy_slabs_per_tile + item.dy_slab) * NB_SLAB_BYTES;
          const uint8_t* x = p.x_stash + ((size_t)tile * p.x_slabs_per_tile + item.x_slab) * NB_SLAB_BYTES;
          const uint8_t* z = p.z_stash + ((size_t)tile * p.z_slabs_per_tile + (n_z ? item.z_slab : 0)) * NB_SLAB_BYTES;
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full[stage], bytes);
            uint8_t* dst = smem + stage * kStageBytes;
            for (int s = 0; s < item.n_dy_slabs; ++s)
              bulk_g2s(dst + s * kHalfSlabBytes, dy + (size_t)s * NB_SLAB_BYTES + half * kHalfSlabBytes,
                       kHalfSlabBytes, &full[stage]);
            for (int s = 0; s < n_x; ++s) {
              const uint8_t* xs = (s == n_x - 1 && item.x2_slab >= 0)
                  ? p.x_stash + ((size_t)tile * p.x_slabs_per_tile + item.x2_slab) * NB_SLAB_BYTES
                  : x + (size_t)s * NB_SLAB_BYTES;
              bulk_g2s(dst + (pos_x + s) * kHalfSlabBytes, xs + half * kHalfSlabBytes, kHalfSlabBytes, &full[stage]);
            }
            for (int s = 0; s < n_z; ++s)
              bulk_g2s(dst + (pos_z + s) * kHalfSlabBytes, z + (size_t)s * NB_SLAB_BYTES + half * kHalfSlabBytes,
                       kHalfSlabBytes, &full[stage]);
            if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, e_phase = 0;
      bool flush_pending = false;
      for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
        const NbWgradItem item = p.items[it];
        if (item.mode == NB_WGRAD_COLSUM) {
          // column-sum item: no MMA — the stages only pass through the four summing warps
          for (int tile = item.tile_begin; tile < item.tile_end; ++tile) {
            for (int half = 0; half < 2; ++half) {
              mbar_wait(&full[stage], phase);
              mbar_arrive(&empty[stage]);
              if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
            }
          }
          continue;
        }
        const int n_mb = (item.n_dy_slabs + 1) / 2;
        const uint32_t pos_x = 2u * (uint32_t)n_mb;
        const uint32_t idesc = umma_idesc(128, item.n_x_slabs * 64, true, true);
        if (flush_pending) {   // the flush of the previous MMA item must have drained TMEM
          mbar_wait(acc_empty, e_phase);
          e_phase ^= 1u;
          tcgen05_fence_after();
        }
        flush_pending = true;
        uint32_t acc = 0;
        for (int tile = item.tile_begin; tile < item.tile_end; ++tile) {
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&full[stage], phase);
            tcgen05_fence_after();
            const uint32_t base = smem_u32(smem + stage * kStageBytes);
            for (int k = 0; k < kHalfRows / 16; ++k) {
              const uint64_t bdesc = umma_desc_mnmajor(base + pos_x * kHalfSlabBytes, kHalfSlabBytes, k);
              for (int mb = 0; mb < n_mb; ++mb) {
                const uint64_t adesc = umma_desc_mnmajor(base + mb * 2 * kHalfSlabBytes, kHalfSlabBytes, k);
                umma(tmem_base + (uint32_t)(mb * 256), adesc, bdesc, idesc, acc);
              }
              acc = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // flush warps: TMEM lane = output feature within the 128-block. While the item accumulates,
    // warp w sums the columns of dY slab w of every stage (bias gradient).
    uint32_t f_phase = 0, stage = 0, phase = 0;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int rg = lane >> 3, pc = lane & 7;   // 16-row group of the half slab, physical 16-byte chunk
    for (int it = blockIdx.x; it < p.n_items; it += gridDim.x) {
      const NbWgradItem item = p.items[it];
      const int n_mb = (item.n_dy_slabs + 1) / 2;
      const bool colsum = item.mode == NB_WGRAD_COLSUM;
      // z duty of this warp's dY slab: bias and Gaussian-width sums
      const bool zduty = item.z_slab >= 0 && warp >= item.z_first && warp < item.z_first + item.n_z_slabs &&
                         warp < item.n_dy_slabs;
      const bool sums = (item.bias_dst >= 0 && warp < item.n_dy_slabs) || zduty;
      const uint32_t z_pos = 2u * (uint32_t)n_mb + (colsum ? 0u : (uint32_t)item.n_x_slabs) + (uint32_t)(warp - item.z_first);
      float bacc[8][8], zacc[8][8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) { bacc[q][e] = 0.f; zacc[q][e] = 0.f; }
      for (int tile = item.tile_begin; tile < item.tile_end; ++tile) {
        for (int half = 0; half < 2; ++half) {
          mbar_wait(&full[stage], phase);
          if (sums) {
            const uint32_t lane_off = (uint32_t)(rg * 16) * 128u + (uint32_t)pc * 16u;
            const uint8_t* base = smem + stage * kStageBytes + (uint32_t)warp * kHalfSlabBytes + lane_off;
            const uint8_t* zslab = smem + stage * kStageBytes + (zduty ? z_pos : 0u) * kHalfSlabBytes + (uint32_t)(rg * 16) * 128u;
            const uint8_t* zbase_x[2] = {zslab + (uint32_t)pc * 16u, zslab + (uint32_t)(pc ^ 1) * 16u};
#pragma unroll
            for (int r8 = 0; r8 < 2; ++r8) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {   // row & 7 == q: logical chunk pc ^ q
                const uint4 v = *reinterpret_cast<const uint4*>(base + (uint32_t)(r8 * 8 + q) * 128u);
                bacc[q][0] += __uint_as_float(v.x << 16); bacc[q][1] += __uint_as_float(v.x & 0xffff0000u);
                bacc[q][2] += __uint_as_float(v.y << 16); bacc[q][3] += __uint_as_float(v.y & 0xffff0000u);
                bacc[q][4] += __uint_as_float(v.z << 16); bacc[q][5] += __uint_as_float(v.z & 0xffff0000u);
                bacc[q][6] += __uint_as_float(v.w << 16); bacc[q][7] += __uint_as_float(v.w & 0xffff0000u);
                if (zduty) {   // Gaussian width gradient: sum over samples of z * dz. The z stash keeps the halves of a
                               // 32-byte sector in natural order (garf_kernels.cuh): on odd rows the neighbouring chunk
                  const uint4 z = *reinterpret_cast<const uint4*>(zbase_x[q & 1] + (uint32_t)(r8 * 8 + q) * 128u);
                  zacc[q][0] = fmaf(__uint_as_float(z.x << 16), __uint_as_float(v.x << 16), zacc[q][0]);
                  zacc[q][1] = fmaf(__uint_as_float(z.x & 0xffff0000u), __uint_as_float(v.x & 0xffff0000u), zacc[q][1]);
                  zacc[q][2] = fmaf(__uint_as_float(z.y << 16), __uint_as_float(v.y << 16), zacc[q][2]);
                  zacc[q][3] = fmaf(__uint_as_float(z.y & 0xffff0000u), __uint_as_float(v.y & 0xffff0000u), zacc[q][3]);
                  zacc[q][4] = fmaf(__uint_as_float(z.z << 16), __uint_as_float(v.z << 16), zacc[q][4]);
                  zacc[q][5] = fmaf(__uint_as_float(z.z & 0xffff0000u), __uint_as_float(v.z & 0xffff0000u), zacc[q][5]);
                  zacc[q][6] = fmaf(__uint_as_float(z.w << 16), __uint_as_float(v.w << 16), zacc[q][6]);
                  zacc[q][7] = fmaf(__uint_as_float(z.w & 0xffff0000u), __uint_as_float(v.w & 0xffff0000u), zacc[q][7]);
                }
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (sums) {
        // lane (.., pc) holds in bacc[q] the partial sums of logical chunk pc ^ q: lane c gathers
        // bacc[q] from lane c ^ q, then the four row groups are folded
        float tot[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tot[e] = bacc[0][e];
#pragma unroll
        for (int q = 1; q < 8; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) tot[e] += __shfl_xor_sync(0xffffffffu, bacc[q][e], q);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          tot[e] += __shfl_xor_sync(0xffffffffu, tot[e], 8);
          tot[e] += __shfl_xor_sync(0xffffffffu, tot[e], 16);
        }
        const int bias_dst = zduty ? item.zbias_dst : item.bias_dst;
        if (rg == 0 && item.tile_end > item.tile_begin && bias_dst >= 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int m = warp * 64 + pc * 8 + e;
            if (m < item.m_real) atomicAdd(p.d_params + bias_dst + m, tot[e]);
          }
        }
        if (zduty && item.coef_dst >= 0) {
          float zt[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) zt[e] = zacc[0][e];
#pragma unroll
          for (int q = 1; q < 8; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) zt[e] += __shfl_xor_sync(0xffffffffu, zacc[q][e], q);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            zt[e] += __shfl_xor_sync(0xffffffffu, zt[e], 8);
            zt[e] += __shfl_xor_sync(0xffffffffu, zt[e], 16);
          }
          if (rg == 0 && item.tile_end > item.tile_begin) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int m = warp * 64 + pc * 8 + e;
              if (m < item.m_real) {
                // y = exp(-z^2 v), v = s^2 + 1e-6: dL/ds = 2 s dL/dv, dL/dv = sum z dz / (2 v)
                const float sdev = __ldg(p.params + item.coef_dst + m);
                atomicAdd(p.d_params + item.coef_dst + m, zt[e] * sdev / (sdev * sdev + 1e-6f));
              }
            }
          }
        }
      }
      if (colsum) continue;
      mbar_wait(acc_full, f_phase);
      f_phase ^= 1u;
      tcgen05_fence_after();
      if (item.tile_end > item.tile_begin) {
        for (int mb = 0; mb < n_mb; ++mb) {
          const int m = mb * 128 + threadIdx.x;
          // whole warps skip together: tcgen05.ld is warp-collective
          const bool warp_has_rows = (mb * 128 + warp * 32) < item.m_real;
          if (!warp_has_rows) continue;
          float* dst_row = p.d_params + item.dst + (long long)m * item.ld;
          for (int g = 0; g < item.n_real; g += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + (uint32_t)(mb * 256 + g), v);
            tmem_ld_wait();
            if (m < item.m_real) {
              // 16-byte vector reductions where the row segment allows it: a quarter of the
              // requests the L2 atomic units see (all SMs flush at about the same time)
              if (g + 32 <= item.n_real && ((reinterpret_cast<uintptr_t>(dst_row + g) & 15u) == 0)) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + g + i),
                               "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                               "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                               : "memory");
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (g + i < item.n_real) atomicAdd(dst_row + g + i, __uint_as_float(v[i]));
              }
            }
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(acc_empty);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_mlp_wgrad(const NbWgradItem* items_dev, int n_items, const void* x_stash,
                                  int x_slabs_per_tile, const void* dy_stash,
                                  int dy_slabs_per_tile, const void* z_stash, int z_slabs_per_tile,
                                  const float* params, float* d_params, void* stream) {
  NB_CHECK_ARG(n_items >= 0 && (n_items == 0 || (items_dev && x_stash && dy_stash && d_params)),
               "mlp_wgrad: bad arguments");
  if (n_items == 0) return NERFB200_OK;
  WgradParams p;
  p.items = items_dev;
  p.n_items = n_items;
  p.counter = nullptr;
  p.x_stash = reinterpret_cast<const uint8_t*>(x_stash);
  p.dy_stash = reinterpret_cast<const uint8_t*>(dy_stash);
  p.z_stash = reinterpret_cast<const uint8_t*>(z_stash);
  p.x_slabs_per_tile = x_slabs_per_tile;
  p.dy_slabs_per_tile = dy_slabs_per_tile;
  p.z_slabs_per_tile = z_slabs_per_tile;
  p.d_params = d_params;
  p.params = params;
  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kWgSmemBytes));
    configured = true;
  }
  const int grid = n_items < sm_count() ? n_items : sm_count();
  mlp_wgrad_kernel<<<grid, kWgThreads, kWgSmemBytes, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
