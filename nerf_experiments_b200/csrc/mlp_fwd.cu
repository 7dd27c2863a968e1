// Fused radiance-MLP forward on sm_100a: positions + positional encoding in registers, every
// Linear layer as tcgen05 bf16 MMAs with the 128-sample activation tile resident in shared
// memory / TMEM, weights streamed from L2 by the TMA engine (cp.async.bulk), bias/activation
// epilogues from TMEM. Replaces NerfInterpolation._compute_positions + NerfModel.forward
// (reference barf/model_interpolation.py:288-312, barf/model_interpolation_architecture.py:96-141)
// — 12 cuBLAS GEMMs + ~40 elementwise launches in the reference — by one persistent launch.
//
// CTA = 6 warps: warps 0-3 own one tile row each (TMEM lane = row), warp 4 issues MMAs,
// warp 5 streams weight images through a 3-stage ring.
#include "common.cuh"
#include "mlp.h"
#include "mlp_kernels.cuh"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {
namespace {

using namespace tc;

// packs relu(a+b) (or a+b) pairs to bf16x2 and collects the ReLU sign bits
__device__ __forceinline__ uint32_t act_pack(float a0, float a1, bool relu, uint32_t& bits, int i) {
  if (relu) {
    // bit = 1 <=> value > 0 (sign bit clear and non-zero)
    bits |= ((a0 > 0.f) ? 1u : 0u) << i;
    bits |= ((a1 > 0.f) ? 1u : 0u) << (i + 1);
    a0 = fmaxf(a0, 0.f);
    a1 = fmaxf(a1, 0.f);
  }
  return pack_bf16(a0, a1);
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fwd_kernel(const __grid_constant__ MlpFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem sm(smem_raw, p.prog.n_slabs, p.prog.n_stages);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NB_MAX_RING_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(sm.a_ready, kRowThreads);
    mbar_init(sm.acc_full, 1);
    fence_barrier_init();
    pe_fill_mask(p.pe_pos, p.alpha_pos, sm.mask_pos);
    pe_fill_mask(p.pe_dir, p.alpha_dir, sm.mask_dir);
  }
  for (int i = threadIdx.x; i < p.n_bias_floats; i += blockDim.x) sm.floats[i] = p.bias[i];
  if (warp == kMmaWarp) tmem_alloc(sm.tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarp) {
    if (lane == 0) weight_producer_loop(p.prog, p.wpack, sm, n_tiles);
  } else if (warp == kMmaWarp) {
    mma_issuer_loop(p.prog, sm, tmem_base, n_tiles);
  } else {
    // ---------------- row threads: PE prologue + epilogues ----------------
    const int row = threadIdx.x & (kHalfThreads - 1);   // tile row
    const int half = threadIdx.x >> 7;                  // which half of the columns
    const bool leader = (row == 0);                     // owns this half's stash copies
    const int bar_id = 1 + half;                        // named barrier of this half
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_phase = 0;
    const bool training = (p.stash != nullptr);
    StashQueue sq;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      uint8_t* tile_stash = training ? p.stash + (size_t)tile * p.prog.stash_slabs_per_tile * NB_SLAB_BYTES : nullptr;

      // the previous tile's stash copies must have finished reading the encoding slabs
      if (training) {
        if (leader) sq.wait_all();
        named_bar_sync(bar_id, kHalfThreads);
      }
      // ---- positions and positional encodings: half 0 -> position slab; half 1 -> direction slab
      //      when it has a slab of its own (otherwise it shares slab 4 and is written later)
      {
        PeSample ps;
        load_sample(p.in, n, ps);
        if (half == 0) {
          encode_to_slab(p.pe_pos, sm.mask_pos, ps, sm, row);
        } else if (p.pe_dir.encode_before_op == 0) {
          PeSample pd = ps;  // the direction encoder sees the direction as its "position"
          pd.x[0] = ps.dir[0]; pd.x[1] = ps.dir[1]; pd.x[2] = ps.dir[2];
          encode_to_slab(p.pe_dir, sm.mask_dir, pd, sm, row);
        }
      }
      fence_proxy_async();
      mbar_arrive(sm.a_ready);
      if (training) {
        named_bar_sync(bar_id, kHalfThreads);
        if (leader) {
          const NbPeCfg& cfg = half == 0 ? p.pe_pos : p.pe_dir;
          if (cfg.slab >= 0 && cfg.stash_slab >= 0 && (half == 0 || cfg.encode_before_op == 0)) {
            bulk_s2g(tile_stash + (size_t)cfg.stash_slab * NB_SLAB_BYTES, sm.slab(cfg.slab), NB_SLAB_BYTES);
            bulk_commit();
          }
          sq.begin_batch();   // that copy reads slab 4/5 only: act slabs may be rewritten at once
        }
      }

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const bool last = (oi == p.prog.n_ops - 1);
        const float* bias = sm.floats + op.bias_off;
        const bool stores_act = (op.epi == NB_EPI_RELU || op.epi == NB_EPI_LINEAR ||
                                 op.epi == NB_EPI_LINEAR_SIGMA || op.epi == NB_EPI_RELU_SIGMA);
        // slabs of this half: [c_begin, c_end)
        const int c_mid = (op.out_chunks + 1) >> 1;
        const int c_begin = half == 0 ? 0 : c_mid;
        const int c_end = half == 0 ? c_mid : op.out_chunks;
        if (p.pe_dir.slab >= 0 && p.pe_dir.encode_before_op == oi && oi > 0 && half == 0) {
          // While this op's MMAs run: the direction encoding takes over the slab the position
          // encoding no longer needs (its last reader was the previous op).
          if (training) {
            if (leader) sq.wait_all();
            named_bar_sync(bar_id, kHalfThreads);
          }
          PeSample pd;
          load_sample(p.in, n, pd);
          pd.x[0] = pd.dir[0]; pd.x[1] = pd.dir[1]; pd.x[2] = pd.dir[2];
          encode_to_slab(p.pe_dir, sm.mask_dir, pd, sm, row);
          if (training && p.pe_dir.stash_slab >= 0) {
            fence_proxy_async();
            named_bar_sync(bar_id, kHalfThreads);
            if (leader) {
              bulk_s2g(tile_stash + (size_t)p.pe_dir.stash_slab * NB_SLAB_BYTES, sm.slab(p.pe_dir.slab), NB_SLAB_BYTES);
              bulk_commit();
              sq.begin_batch();
            }
          }
        }
        mbar_wait(sm.acc_full, acc_phase);
        acc_phase ^= 1u;
        tcgen05_fence_after();
        NB_TRACE(oi * 4 + 2, threadIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x));
        if (stores_act) {
          const bool relu = (op.epi == NB_EPI_RELU || op.epi == NB_EPI_RELU_SIGMA);
          uint32_t* mask_out = (training && op.mask_word >= 0)
              ? p.masks + ((size_t)tile * p.prog.mask_words_per_tile + op.mask_word) * NB_TILE_ROWS + row
              : nullptr;
          for (int g = 2 * c_begin; g < 2 * c_end; ++g) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + (uint32_t)(g * 32), v);
            if (training && (g & 1) == 0) {
              // slab g/2 is about to be rewritten: its stash copy must have drained
              if (leader) sq.wait_slab((g >> 1) - c_begin);
              named_bar_sync(bar_id, kHalfThreads);
            }
            tmem_ld_wait();
            uint32_t bits = 0;
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 bq = *reinterpret_cast<const float4*>(bias + g * 32 + i);
              packed[i >> 1] = act_pack(__uint_as_float(v[i]) + bq.x, __uint_as_float(v[i + 1]) + bq.y, relu, bits, i);
              packed[(i >> 1) + 1] = act_pack(__uint_as_float(v[i + 2]) + bq.z, __uint_as_float(v[i + 3]) + bq.w, relu, bits, i + 2);
            }
            uint8_t* slab = sm.slab(g >> 1);
            const int chunk0 = (g & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t off = (uint32_t)row * 128u + ((uint32_t)((chunk0 + q) ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(slab + off) =
                  make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
            if (mask_out != nullptr && relu) mask_out[(size_t)g * NB_TILE_ROWS] = bits;
          }
          if ((op.epi == NB_EPI_LINEAR_SIGMA || op.epi == NB_EPI_RELU_SIGMA) && half == 1) {
            uint32_t v[16];
            tmem_ld16(tmem_lane + (uint32_t)op.blocks[1].tmem_col, v);
            tmem_ld_wait();
            const float pre = __uint_as_float(v[0]) + bias[op.blocks[0].n];
            if (valid) p.out_sigma[n] = softplus8(pre + p.sigma_bias);
          }
        } else if (half == 0) {
          // NB_EPI_RGB / NB_EPI_RGB_SIGMA: first 16 accumulator columns hold the outputs
          uint32_t v[16];
          tmem_ld16(tmem_lane, v);
          tmem_ld_wait();
          if (valid) {
            p.out_rgb[n * 3 + 0] = sigmoidf(__uint_as_float(v[0]) + bias[0]);
            p.out_rgb[n * 3 + 1] = sigmoidf(__uint_as_float(v[1]) + bias[1]);
            p.out_rgb[n * 3 + 2] = sigmoidf(__uint_as_float(v[2]) + bias[2]);
            if (op.epi == NB_EPI_RGB_SIGMA)
              p.out_sigma[n] = softplus8(__uint_as_float(v[3]) + bias[3] + p.sigma_bias);
          }
        }
        tcgen05_fence_before();
        NB_TRACE(oi * 4 + 3, threadIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x));
        if (!last) {
          fence_proxy_async();
          mbar_arrive(sm.a_ready);
        }
        if (training && stores_act && op.stash_slab >= 0) {
          if (last) fence_proxy_async();
          named_bar_sync(bar_id, kHalfThreads);
          if (leader) {
            sq.begin_batch();
            for (int c = c_begin; c < c_end; ++c)
              sq.push(tile_stash + (size_t)(op.stash_slab + c) * NB_SLAB_BYTES, sm.slab(c), NB_SLAB_BYTES);
          }
        }
      }
    }
    if (training && leader) bulk_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_debug_trace_fwd(long long* trace_dev) {
  NB_CHECK_CUDA(cudaMemcpyToSymbol(g_trace, &trace_dev, sizeof(trace_dev)));
  return NERFB200_OK;
}

extern "C" int nerfb200_mlp_fwd(const void* program_host, const void* wpack, const float* bias,
                                const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                                const NbPeCfg* pe_dir_host, const float* alpha_pos,
                                const float* alpha_dir, float sigma_bias, float* out_sigma,
                                float* out_rgb, void* stash, uint32_t* masks, int n_bias_floats,
                                void* stream) {
  NB_CHECK_ARG(program_host && wpack && bias && in_host && pe_pos_host && pe_dir_host,
               "mlp_fwd: null pointer");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  NB_CHECK_ARG(prog->n_ops >= 1 && prog->n_ops <= NB_MAX_OPS, "mlp_fwd: bad program (n_ops=%d)", prog->n_ops);
  NB_CHECK_ARG(in_host->N >= 0 && in_host->S >= 1, "mlp_fwd: bad shape N=%lld S=%d", (long long)in_host->N, in_host->S);
  NB_CHECK_ARG(out_sigma && out_rgb, "mlp_fwd: null output");
  NB_CHECK_ARG((stash == nullptr) == (masks == nullptr), "mlp_fwd: stash and masks go together");
  NB_CHECK_ARG(n_bias_floats >= 1 && n_bias_floats <= (int)MlpSmem::kMaxBiasFloats,
               "mlp_fwd: %d packed bias slots (max %d)", n_bias_floats, (int)MlpSmem::kMaxBiasFloats);
  if (in_host->N == 0) return NERFB200_OK;
  int rc = validate_program(*prog);
  if (rc != NERFB200_OK) return rc;

  MlpFwdParams p;
  p.prog = *prog;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack);
  p.bias = bias;
  p.in = *in_host;
  p.N = (int)in_host->N;
  p.pe_pos = *pe_pos_host;
  p.pe_dir = *pe_dir_host;
  p.alpha_pos = alpha_pos;
  p.alpha_dir = alpha_dir;
  p.sigma_bias = sigma_bias;
  p.out_sigma = out_sigma;
  p.out_rgb = out_rgb;
  p.stash = reinterpret_cast<uint8_t*>(stash);
  p.masks = masks;
  p.n_bias_floats = n_bias_floats;

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)MlpSmem::bytes(5, 4)));  // the larger of the two layouts
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  mlp_fwd_kernel<<<grid, kMlpThreads, MlpSmem::bytes(prog->n_slabs, prog->n_stages), (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
