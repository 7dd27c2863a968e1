// Fused radiance-MLP backward (data gradients) on sm_100a: the mirror image of mlp_fwd.cu.
// One 128-sample tile walks the layers in reverse; dY lives in shared memory (A operand), the
// transposed weight images stream through the TMA ring (B operand), dX accumulates in TMEM and
// the epilogue applies the ReLU mask bits saved by the forward pass. Every dY tile is also
// written (bf16 slabs) to HBM for the weight-gradient kernel (mlp_wgrad.cu). Gradients w.r.t.
// the positional encodings accumulate in two dedicated TMEM blocks across layers and are
// pushed through the encoding at the end to give dL/d(ray origin), dL/d(ray direction) — the
// path the reference's autograd takes to the camera-pose parameters
// (barf/model_camera_extrinsics.py:77-85). Bias gradients (column sums of the dY slabs) are
// left to the weight-gradient kernel, which streams the same slabs anyway.
#include "common.cuh"
#include "mlp.h"
#include "mlp_kernels.cuh"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {
namespace {

using namespace tc;

struct MlpBwdParams {
  NbProgram prog;
  const uint8_t* wpack;      // transposed weight images
  NbMlpInputs in;
  int N;
  NbPeCfg pe_pos, pe_dir;
  const float* alpha_pos;
  const float* alpha_dir;
  const float* sigma;        // forward outputs
  const float* rgb;
  const float* g_sigma;      // upstream gradients (NULL = zero)
  const float* g_rgb;
  const uint32_t* masks;
  int fwd_mask_words_per_tile;
  uint8_t* dy_stash;
  int head_sigma_col3;       // 1: delayed density (sigma is column 3 of the output layer)
  int want_input_grads;
  int pos_grad_cols, dir_grad_cols;   // TMEM columns holding d(encoding), 0 = none
  float* d_ray_o;            // (B,3) += (rays mode)
  float* d_ray_d;            // (B,3) +=
  float* d_pos;              // (N,3)    (samples mode)
  float* d_dir;              // (N,3)
};

constexpr int kBwdThreads = kFwdThreads;   // same roles as the forward kernel

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// where a backward production goes in the per-tile dY stash
struct BwdStashDst {
  const MlpBwdParams& p;
  __device__ __forceinline__ int operator()(int ph, int s) const {
    if (ph < 0) return 0;                                   // head gradient
    const NbOp& op = p.prog.ops[ph];
    if (op.stash_slab < 0) return -1;
    return s == kBwdAuxSlab ? op.stash_slab + op.out_chunks : op.stash_slab + s;
  }
};

// One 16-column group of a data-gradient epilogue: ReLU mask and bf16 pack ...
__device__ __forceinline__ void grad_math16(const uint32_t (&v)[16], uint32_t bits, uint32_t (&packed)[8]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const float a = ((bits >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
    const float b = ((bits >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
    packed[i >> 1] = pack_bf16(a, b);
  }
}
// ... and the two swizzled 16-byte stores of (row, column quarter cq): the thread's two chunk
// addresses inside slab 0 are computed once per kernel, the slab index is an immediate offset
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void store_packed16(const uint32_t (&packed)[8], uint32_t sts0, uint32_t sts1, int j) {
  sts128(sts0 + (uint32_t)j * NB_SLAB_BYTES, packed[0], packed[1], packed[2], packed[3]);
  sts128(sts1 + (uint32_t)j * NB_SLAB_BYTES, packed[4], packed[5], packed[6], packed[7]);
}

__global__ void __launch_bounds__(kBwdThreads, 1)
mlp_bwd_kernel(const __grid_constant__ MlpBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem sm(smem_raw, p.prog.n_slabs, p.prog.n_stages);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;

  TileSchedule sched;
  sched.is_bwd = 1;
  sched.start_mask = 1u;          // head gradient -> slab 0
  sched.reencode_op = -1;
  sched.reencode_mask = 0u;

  if (threadIdx.x == 0) {
    sm.init_barriers();
    pe_fill_mask(p.pe_pos, p.alpha_pos, sm.mask_pos);
    pe_fill_mask(p.pe_dir, p.alpha_dir, sm.mask_dir);
  }
  if (warp == kMmaWarp) tmem_alloc(sm.tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarp) {
    if (lane == 0) weight_producer_loop(p.prog, p.wpack, sm, n_tiles);
  } else if (warp == kMmaWarp) {
    mma_issuer_loop(p.prog, sched, sm, tmem_base, n_tiles);
  } else if (warp == kStashWarp) {
    if (lane == 0)
      stash_loop(p.prog, sched, sm, n_tiles, p.dy_stash, p.prog.stash_slabs_per_tile, BwdStashDst{p});
  } else {
    const int row = threadIdx.x & kTileRowMask;
    const int cq = threadIdx.x >> 7;                    // which 16-column quarter of every slab
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sts0 = smem_u32(sm.slab(0)) + (uint32_t)row * 128u + ((uint32_t)((2 * cq) ^ (row & 7)) << 4);
    uint32_t sts1 = smem_u32(sm.slab(0)) + (uint32_t)row * 128u + ((uint32_t)((2 * cq + 1) ^ (row & 7)) << 4);
    // opaque to the compiler: otherwise it rematerialises both addresses per slab
    asm volatile("" : "+r"(sts0), "+r"(sts1));
    DrainBits drain;
    NB_TRACE_INIT();
    uint32_t g_op = 0, tile_phase = 0;
    const uint16_t* const masks16 = reinterpret_cast<const uint16_t*>(p.masks);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      // sign bits of this row: one 32-bit word per 32-column group, this thread owns half of it
      const uint16_t* tile_masks = masks16 + (((size_t)tile * p.fwd_mask_words_per_tile + (cq >> 1)) * NB_TILE_ROWS + row) * 2 + (cq & 1);

      if (tile != (int)blockIdx.x) {   // the MMA warp has consumed every publication of the previous tile
        mbar_wait(sm.tile_done, tile_phase);
        tile_phase ^= 1u;
      }
      // All MMAs of the previous tile are complete (its last accumulator was read); slab 0 may be
      // rewritten once its stash copy has drained.
      NB_TRACE(392, threadIdx.x == 0);
      drain.acquire_mask(sm.slab_drained, sched.start_mask, lane);
      NB_TRACE(393, threadIdx.x == 0);

      // ---- head: gradients w.r.t. the pre-activations of the output layer (quarter 0) ----
      float d_sigma_pre = 0.f;
      if (cq <= 1) {
        const float sg = p.sigma[n];
        if (valid && p.g_sigma != nullptr)
          d_sigma_pre = p.g_sigma[n] * (sg > 8.f ? 1.f : -expm1f(-sg));   // sigmoid(x) = 1 - exp(-softplus(x)), without cancellation for tiny densities
      }
      if (cq == 0) {
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid && p.g_rgb != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float y = p.rgb[n * 3 + c];
            d4[c] = p.g_rgb[n * 3 + c] * y * (1.f - y);
          }
        }
        if (p.head_sigma_col3) d4[3] = d_sigma_pre;
        uint8_t* slab = sm.slab(0);
        // columns 0..15 hold the head gradient (only k16 = 1 is consumed by the MMA); the rest of
        // the row is cleared because wgrad consumes the stashed slab as a full 64-column operand
        *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16(d4[0], d4[1]), pack_bf16(d4[2], d4[3]), 0u, 0u);
#pragma unroll
        for (int q = 1; q < 8; ++q)
          *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
              make_uint4(0u, 0u, 0u, 0u);
      }
      NB_TRACE(394, threadIdx.x == 0);
      signal_slabs(sm.slab_ready, sched.start_mask, lane);
      NB_TRACE(395, threadIdx.x == 0);
      drain.produced(sched.start_mask);

      // this thread's share of the gradients w.r.t. the query position / direction of the row
      float dx[3] = {0.f, 0.f, 0.f}, dd[3] = {0.f, 0.f, 0.f};

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const bool stores = bwd_stores(op.epi);
        const bool masked = (op.epi == NB_BEPI_MASK || op.epi == NB_BEPI_MASK_SIGMA);
        const bool with_sigma = bwd_with_sigma(op.epi);
        const uint32_t buf = g_op & 1u;
        const uint32_t acc = tmem_lane + buf * kAccCols;
        const int oc = op.out_chunks;
        // ReLU sign bits of the producing layer: fetched before the accumulator is waited for
        uint32_t mbits[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mbits[j] = (masked && j < oc) ? (uint32_t)__ldg(tile_masks + (size_t)(op.mask_word + 2 * j) * NB_TILE_ROWS * 2)
                                        : 0xffffu;
        warp_mbar_wait(&sm.acc_full[buf], (g_op >> 1) & 1u, lane);
        tcgen05_fence_after();
        NB_TRACE(oi * 4 + 2, threadIdx.x == 0);
        if (stores) {
          const uint32_t will_stash = op.stash_slab >= 0 ? 1u : 0u;
          // the TMEM load of slab j+1 is in flight during the math of slab j
          uint32_t va[16], vb[16], packed[8];
          const uint32_t acc_q = acc + (uint32_t)(16 * cq);
          tmem_ld16(acc_q, va);
          auto finish = [&](int j) {
            store_packed16(packed, sts0, sts1, j);
            signal_slab(sm.slab_ready, j, lane);
            if (will_stash) { drain.pending |= 1u << j; drain.last = j; }
          };
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            if (j < oc) {
              drain.acquire(sm.slab_drained, j, lane);   // the stash copy of the slab's previous contents has
              tmem_ld_wait16(va);                        // drained (checked while the TMEM load is in flight)
              if (j + 1 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 1)), vb);
              grad_math16(va, mbits[j], packed);
              finish(j);
            }
            if (j + 1 < oc) {
              drain.acquire(sm.slab_drained, j + 1, lane);
              tmem_ld_wait16(vb);
              if (j + 2 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 2)), va);
              grad_math16(vb, mbits[j + 1], packed);
              finish(j + 1);
            }
          }
          if (with_sigma) {
            // d(sigma_pre) becomes column 0 of the aux slab (an extra 16-wide K step)
            drain.acquire(sm.slab_drained, kBwdAuxSlab, lane);
            if (cq == 1) {
              uint8_t* slab = sm.slab(kBwdAuxSlab);
              *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
                  make_uint4(pack_bf16(d_sigma_pre, 0.f), 0u, 0u, 0u);
#pragma unroll
              for (int q = 1; q < 8; ++q)
                *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
                    make_uint4(0u, 0u, 0u, 0u);
            }
            signal_slab(sm.slab_ready, kBwdAuxSlab, lane);
            drain.produced(will_stash << kBwdAuxSlab);
          }
        } else if (op.epi == NB_BEPI_PEGRAD_POS || op.epi == NB_BEPI_PEGRAD_DIR) {
          // d(encoding) of this layer, canonical column order -> d(position) / d(direction);
          // every quarter pushes its 16 columns through the encoder
          uint32_t g[16];
          tmem_ld16(acc + (uint32_t)(16 * cq), g);
          tmem_ld_wait16(g);
          PeSample ps;
          load_sample(p.in, n, ps);
          const bool is_pos = (op.epi == NB_BEPI_PEGRAD_POS);
          if (!is_pos) { ps.x[0] = ps.dir[0]; ps.x[1] = ps.dir[1]; ps.x[2] = ps.dir[2]; }
          const NbPeCfg& cfg = is_pos ? p.pe_pos : p.pe_dir;
          const float* msk = is_pos ? sm.mask_pos : sm.mask_dir;
          float g3[3], dsc;
          switch (cq) {
            case 0: pe_backward_canon_quarter<0>(cfg, msk, ps, g, g3, dsc); break;
            case 1: pe_backward_canon_quarter<1>(cfg, msk, ps, g, g3, dsc); break;
            case 2: pe_backward_canon_quarter<2>(cfg, msk, ps, g, g3, dsc); break;
            default: pe_backward_canon_quarter<3>(cfg, msk, ps, g, g3, dsc); break;
          }
          if (is_pos) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              dx[c] += g3[c];
              dd[c] += dsc * g3[c];
            }
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) dd[c] += g3[c];
          }
        }
        // this accumulator buffer has been read out
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tmem_free[buf]);
        NB_TRACE(oi * 4 + 3, threadIdx.x == 0);
        ++g_op;
      }

      // ---- d(position), d(direction) of the samples -> rays (the four quarters add up) ----
      NB_TRACE(396, threadIdx.x == 0);
      if (p.want_input_grads) {
        if (p.in.pos != nullptr) {
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              atomicAdd(p.d_pos + n * 3 + c, dx[c]);
              atomicAdd(p.d_dir + n * 3 + c, dd[c]);
            }
          }
        } else {
          // x = o + t_q d  =>  dL/do += dx, dL/dd += t_q dx (+ the direction-encoding part)
          const float t0 = p.in.t_start ? __ldg(p.in.t_start + n) : 0.f;
          const float t1 = p.in.t_end ? __ldg(p.in.t_end + n) : t0;
          const float tq = p.in.t_mode == 0 ? t0 : (t0 + t1) * 0.5f;
          float go[3], gd[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            go[c] = valid ? dx[c] : 0.f;
            gd[c] = valid ? (tq * dx[c] + dd[c]) : 0.f;
          }
          const long long ray = n / p.in.S;
          const long long ray0 = __shfl_sync(0xffffffffu, ray, 0);
          const bool uniform = __all_sync(0xffffffffu, ray == ray0);
          if (uniform) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              go[c] = warp_sum(go[c]);
              gd[c] = warp_sum(gd[c]);
            }
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                atomicAdd(p.d_ray_o + ray * 3 + c, go[c]);
                atomicAdd(p.d_ray_d + ray * 3 + c, gd[c]);
              }
            }
          } else if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              atomicAdd(p.d_ray_o + ray * 3 + c, go[c]);
              atomicAdd(p.d_ray_d + ray * 3 + c, gd[c]);
            }
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_debug_trace_bwd(long long* trace_dev) {
  NB_CHECK_CUDA(cudaMemcpyToSymbol(g_trace, &trace_dev, sizeof(trace_dev)));
  return NERFB200_OK;
}

extern "C" int nerfb200_mlp_bwd(const void* program_host, const void* wpack_t,
                                const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                                const NbPeCfg* pe_dir_host, const float* alpha_pos,
                                const float* alpha_dir, const float* sigma, const float* rgb,
                                const float* g_sigma, const float* g_rgb, const uint32_t* masks,
                                int fwd_mask_words_per_tile, void* dy_stash, int head_sigma_col3,
                                int pos_grad_cols, int dir_grad_cols, float* d_ray_o,
                                float* d_ray_d, float* d_pos, float* d_dir, void* stream) {
  NB_CHECK_ARG(program_host && wpack_t && in_host && pe_pos_host && pe_dir_host, "mlp_bwd: null pointer");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  NB_CHECK_ARG(prog->n_ops >= 1 && prog->n_ops <= NB_MAX_OPS, "mlp_bwd: bad program (n_ops=%d)", prog->n_ops);
  NB_CHECK_ARG(sigma && rgb && masks && dy_stash, "mlp_bwd: null buffer");
  NB_CHECK_ARG((pos_grad_cols == 0 || pos_grad_cols == NB_PE_CANON_COLS) &&
               (dir_grad_cols == 0 || dir_grad_cols == NB_PE_CANON_COLS),
               "mlp_bwd: encoding gradient blocks must be 0 or %d (canonical) columns", NB_PE_CANON_COLS);
  NB_CHECK_ARG(pe_pos_host->levels <= NB_PE_CANON_LEVELS && pe_dir_host->levels <= NB_PE_CANON_LEVELS,
               "mlp_bwd: at most %d encoding levels", NB_PE_CANON_LEVELS);
  const bool want = (pos_grad_cols + dir_grad_cols) > 0;
  if (want) {
    if (in_host->pos != nullptr) NB_CHECK_ARG(d_pos && d_dir, "mlp_bwd: d_pos/d_dir required");
    else NB_CHECK_ARG(d_ray_o && d_ray_d, "mlp_bwd: d_ray_o/d_ray_d required");
  }
  if (in_host->N == 0) return NERFB200_OK;
  int rc = validate_program(*prog);
  if (rc != NERFB200_OK) return rc;

  MlpBwdParams p;
  p.prog = *prog;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack_t);
  p.in = *in_host;
  p.N = (int)in_host->N;
  p.pe_pos = *pe_pos_host;
  p.pe_dir = *pe_dir_host;
  p.alpha_pos = alpha_pos;
  p.alpha_dir = alpha_dir;
  p.sigma = sigma;
  p.rgb = rgb;
  p.g_sigma = g_sigma;
  p.g_rgb = g_rgb;
  p.masks = masks;
  p.fwd_mask_words_per_tile = fwd_mask_words_per_tile;
  p.dy_stash = reinterpret_cast<uint8_t*>(dy_stash);
  p.head_sigma_col3 = head_sigma_col3;
  p.want_input_grads = want ? 1 : 0;
  p.pos_grad_cols = pos_grad_cols;
  p.dir_grad_cols = dir_grad_cols;
  p.d_ray_o = d_ray_o;
  p.d_ray_d = d_ray_d;
  p.d_pos = d_pos;
  p.d_dir = d_dir;

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)MlpSmem::bytes(5, 4)));  // the larger of the two layouts
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  mlp_bwd_kernel<<<grid, kBwdThreads, MlpSmem::bytes(prog->n_slabs, prog->n_stages), (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
