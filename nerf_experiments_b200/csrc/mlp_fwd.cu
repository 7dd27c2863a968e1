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

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fwd_kernel(const __grid_constant__ MlpFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem sm(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NB_RING_STAGES; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    mbar_init(sm.a_ready, kRowThreads);
    mbar_init(sm.acc_full, 1);
    fence_barrier_init();
    pe_fill_mask(p.pe_pos, p.alpha_pos, sm.mask_pos);
    pe_fill_mask(p.pe_dir, p.alpha_dir, sm.mask_dir);
  }
  if (warp == kMmaWarp) tmem_alloc(sm.tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarp) {
    // ---------------- weight producer ----------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int oi = 0; oi < p.prog.n_ops; ++oi) {
          const NbOp& op = p.prog.ops[oi];
          for (int c = 0; c < op.n_chunks; ++c) {
            mbar_wait(&sm.empty[stage], phase ^ 1u);
            const uint32_t bytes = (uint32_t)op.w_rows[c] * 128u;
            mbar_arrive_expect_tx(&sm.full[stage], bytes);
            bulk_g2s(sm.ring(stage), p.wpack + (size_t)op.w_off[c] * 1024u, bytes, &sm.full[stage]);
            if (++stage == NB_RING_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int oi = 0; oi < p.prog.n_ops; ++oi) {
          const NbOp& op = p.prog.ops[oi];
          mbar_wait(sm.a_ready, a_phase);
          a_phase ^= 1u;
          tcgen05_fence_after();
          for (int c = 0; c < op.n_chunks; ++c) {
            mbar_wait(&sm.full[stage], phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(sm.slab(op.a_src[c]));
            const uint32_t b_addr = smem_u32(sm.ring(stage));
            for (int k = 0; k < op.k16[c]; ++k) {
              const uint64_t adesc = umma_desc_kmajor(a_addr, 0, k);
              for (int b = 0; b < op.n_blocks; ++b) {
                const NbBlock& blk = op.blocks[b];
                const uint64_t bdesc = umma_desc_kmajor(b_addr, blk.row0, k);
                const uint32_t acc = (blk.accum_in || c > 0 || k > 0) ? 1u : 0u;
                umma(tmem_base + (uint32_t)blk.tmem_col, adesc, bdesc,
                     umma_idesc(NB_TILE_ROWS, blk.n, false, false), acc);
              }
            }
            umma_commit(&sm.empty[stage]);
            if (++stage == NB_RING_STAGES) { stage = 0; phase ^= 1u; }
          }
          umma_commit(sm.acc_full);
        }
      }
    }
  } else {
    // ---------------- row threads: PE prologue + epilogues ----------------
    const int row = threadIdx.x;  // 0..127
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t acc_phase = 0;
    const bool training = (p.stash != nullptr);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      uint8_t* tile_stash = training ? p.stash + (size_t)tile * p.prog.stash_slabs_per_tile * NB_SLAB_BYTES : nullptr;

      // previous tile's stash copies must have finished reading the slabs
      if (training) {
        if (threadIdx.x == 0) bulk_wait_read<0>();
        named_bar_sync(1, kRowThreads);
      }
      // ---- positions and positional encodings -> slabs 4 / 5
      PeSample ps;
      load_sample(p.in, n, ps);
      encode_to_slab(p.pe_pos, sm.mask_pos, ps, sm, row);
      {
        PeSample pd = ps;  // the direction encoder sees the direction as its "position"
        pd.x[0] = ps.dir[0]; pd.x[1] = ps.dir[1]; pd.x[2] = ps.dir[2];
        encode_to_slab(p.pe_dir, sm.mask_dir, pd, sm, row);
      }
      fence_proxy_async();
      mbar_arrive(sm.a_ready);
      if (training) {
        named_bar_sync(1, kRowThreads);
        if (threadIdx.x == 0) {
          if (p.pe_pos.slab >= 0 && p.pe_pos.stash_slab >= 0)
            bulk_s2g(tile_stash + (size_t)p.pe_pos.stash_slab * NB_SLAB_BYTES, sm.slab(p.pe_pos.slab), NB_SLAB_BYTES);
          if (p.pe_dir.slab >= 0 && p.pe_dir.stash_slab >= 0)
            bulk_s2g(tile_stash + (size_t)p.pe_dir.stash_slab * NB_SLAB_BYTES, sm.slab(p.pe_dir.slab), NB_SLAB_BYTES);
          bulk_commit();
        }
      }

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const bool last = (oi == p.prog.n_ops - 1);
        mbar_wait(sm.acc_full, acc_phase);
        acc_phase ^= 1u;
        tcgen05_fence_after();
        const float* bias = p.bias + op.bias_off;
        const bool stores_act = (op.epi == NB_EPI_RELU || op.epi == NB_EPI_LINEAR ||
                                 op.epi == NB_EPI_LINEAR_SIGMA || op.epi == NB_EPI_RELU_SIGMA);
        if (stores_act) {
          if (training) {  // an earlier stash copy may still be reading the act slabs
            if (threadIdx.x == 0) bulk_wait_read<0>();
            named_bar_sync(1, kRowThreads);
          }
          const bool relu = (op.epi == NB_EPI_RELU || op.epi == NB_EPI_RELU_SIGMA);
          const int groups = op.out_chunks * 2;  // 32-column groups
          uint32_t* mask_out = (training && op.mask_word >= 0)
              ? p.masks + ((size_t)tile * p.prog.mask_words_per_tile + op.mask_word) * NB_TILE_ROWS + row
              : nullptr;
          for (int g = 0; g < groups; ++g) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + (uint32_t)(g * 32), v);
            tmem_ld_wait();
            uint32_t bits = 0;
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a = __uint_as_float(v[i]) + __ldg(bias + g * 32 + i);
              float b = __uint_as_float(v[i + 1]) + __ldg(bias + g * 32 + i + 1);
              if (relu) {
                bits |= (a > 0.f ? 1u : 0u) << i;
                bits |= (b > 0.f ? 1u : 0u) << (i + 1);
                a = fmaxf(a, 0.f);
                b = fmaxf(b, 0.f);
              }
              packed[i >> 1] = pack_bf16(a, b);
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
          if (op.epi == NB_EPI_LINEAR_SIGMA || op.epi == NB_EPI_RELU_SIGMA) {
            uint32_t v[16];
            tmem_ld16(tmem_lane + (uint32_t)op.blocks[1].tmem_col, v);
            tmem_ld_wait();
            const float pre = __uint_as_float(v[0]) + __ldg(bias + op.blocks[0].n);
            if (valid) p.out_sigma[n] = softplus8(pre + p.sigma_bias);
          }
        } else {
          // NB_EPI_RGB / NB_EPI_RGB_SIGMA: first 16 accumulator columns hold the outputs
          uint32_t v[16];
          tmem_ld16(tmem_lane, v);
          tmem_ld_wait();
          if (valid) {
            p.out_rgb[n * 3 + 0] = sigmoidf(__uint_as_float(v[0]) + __ldg(bias + 0));
            p.out_rgb[n * 3 + 1] = sigmoidf(__uint_as_float(v[1]) + __ldg(bias + 1));
            p.out_rgb[n * 3 + 2] = sigmoidf(__uint_as_float(v[2]) + __ldg(bias + 2));
            if (op.epi == NB_EPI_RGB_SIGMA)
              p.out_sigma[n] = softplus8(__uint_as_float(v[3]) + __ldg(bias + 3) + p.sigma_bias);
          }
        }
        tcgen05_fence_before();
        if (!last) {
          fence_proxy_async();
          mbar_arrive(sm.a_ready);
        }
        if (training && stores_act && op.stash_slab >= 0) {
          if (last) fence_proxy_async();
          named_bar_sync(1, kRowThreads);
          if (threadIdx.x == 0) {
            bulk_s2g(tile_stash + (size_t)op.stash_slab * NB_SLAB_BYTES, sm.slab(0),
                     (uint32_t)op.out_chunks * NB_SLAB_BYTES);
            bulk_commit();
          }
        }
      }
    }
    if (training && threadIdx.x == 0) bulk_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_mlp_fwd(const void* program_host, const void* wpack, const float* bias,
                                const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                                const NbPeCfg* pe_dir_host, const float* alpha_pos,
                                const float* alpha_dir, float sigma_bias, float* out_sigma,
                                float* out_rgb, void* stash, uint32_t* masks, void* stream) {
  NB_CHECK_ARG(program_host && wpack && bias && in_host && pe_pos_host && pe_dir_host,
               "mlp_fwd: null pointer");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  NB_CHECK_ARG(prog->n_ops >= 1 && prog->n_ops <= NB_MAX_OPS, "mlp_fwd: bad program (n_ops=%d)", prog->n_ops);
  NB_CHECK_ARG(in_host->N >= 0 && in_host->S >= 1, "mlp_fwd: bad shape N=%lld S=%d", (long long)in_host->N, in_host->S);
  NB_CHECK_ARG(out_sigma && out_rgb, "mlp_fwd: null output");
  NB_CHECK_ARG((stash == nullptr) == (masks == nullptr), "mlp_fwd: stash and masks go together");
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

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)MlpSmem::kBytes));
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  mlp_fwd_kernel<<<grid, kMlpThreads, MlpSmem::kBytes, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
