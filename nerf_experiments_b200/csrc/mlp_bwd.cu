// Fused radiance-MLP backward (data gradients) on sm_100a: the mirror image of mlp_fwd.cu.
// One 128-sample tile walks the layers in reverse; dY lives in shared memory (A operand), the
// transposed weight images stream through the TMA ring (B operand), dX accumulates in TMEM and
// the epilogue applies the ReLU mask bits saved by the forward pass. Every dY tile is also
// written (bf16 slabs) to HBM for the weight-gradient kernel (mlp_wgrad.cu). Gradients w.r.t.
// the positional encodings accumulate in two dedicated TMEM blocks across layers and are
// pushed through the encoding at the end to give dL/d(ray origin), dL/d(ray direction) — the
// path the reference's autograd takes to the camera-pose parameters
// (barf/model_camera_extrinsics.py:77-85).
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
  int head_bias_off;         // packed bias slot of the output layer
  int n_bias_floats;         // packed bias slots in use
  const int32_t* bias_map;   // packed bias slot -> float index in d_params (-1: padding)
  float* d_params;           // flat fp32 gradient buffer (+=)
};

constexpr uint32_t kTmemPosCol = 256;
constexpr uint32_t kTmemDirCol = 320;

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_bwd_kernel(const __grid_constant__ MlpBwdParams p) {
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
  for (int i = threadIdx.x; i < p.n_bias_floats; i += blockDim.x) sm.floats[i] = 0.f;
  if (warp == kMmaWarp) tmem_alloc(sm.tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarp) {
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
    const int row = threadIdx.x;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      uint8_t* tile_stash = p.dy_stash + (size_t)tile * p.prog.stash_slabs_per_tile * NB_SLAB_BYTES;
      const uint32_t* tile_masks = p.masks + (size_t)tile * p.fwd_mask_words_per_tile * NB_TILE_ROWS + row;

      // earlier stash copies must be done reading the slabs before they are rewritten
      if (threadIdx.x == 0) bulk_wait_read<0>();
      named_bar_sync(1, kRowThreads);

      // ---- head: gradients w.r.t. the pre-activations of the output layer ----
      const float sg = p.sigma[n];
      float d_sigma_pre = 0.f;
      if (valid && p.g_sigma != nullptr)
        d_sigma_pre = p.g_sigma[n] * (sg > 8.f ? 1.f : (1.f - __expf(-sg)));
      {
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        if (valid && p.g_rgb != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float y = p.rgb[n * 3 + c];
            d4[c] = p.g_rgb[n * 3 + c] * y * (1.f - y);
          }
        }
        if (p.head_sigma_col3) d4[3] = d_sigma_pre;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float t = warp_sum(d4[c]);
          if (lane == c) atomicAdd(&sm.floats[p.head_bias_off + c], t);
        }
        uint8_t* slab = sm.slab(0);
        // columns 0..15 (two 16-byte pieces) of the row; only k16 = 1 is consumed
        *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16(d4[0], d4[1]), pack_bf16(d4[2], d4[3]), 0u, 0u);
        *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(1 ^ (row & 7)) << 4)) =
            make_uint4(0u, 0u, 0u, 0u);
        // the stash slab of the head is consumed as a full 64-column slab by wgrad: clear the rest
#pragma unroll
        for (int q = 2; q < 8; ++q)
          *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
              make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(sm.a_ready);
      named_bar_sync(1, kRowThreads);
      if (threadIdx.x == 0) {
        bulk_s2g(tile_stash, sm.slab(0), NB_SLAB_BYTES);   // head dY = stash slab 0
        bulk_commit();
      }

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const bool last = (oi == p.prog.n_ops - 1);
        mbar_wait(sm.acc_full, acc_phase);
        acc_phase ^= 1u;
        tcgen05_fence_after();
        if (op.epi != NB_BEPI_NONE) {
          if (threadIdx.x == 0) bulk_wait_read<0>();
          named_bar_sync(1, kRowThreads);
          const bool masked = (op.epi == NB_BEPI_MASK || op.epi == NB_BEPI_MASK_SIGMA);
          const int groups = op.out_chunks * 2;
          for (int g = 0; g < groups; ++g) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + (uint32_t)(g * 32), v);
            uint32_t bits = 0xffffffffu;
            if (masked) bits = tile_masks[(size_t)(op.mask_word + g) * NB_TILE_ROWS];
            tmem_ld_wait();
            uint32_t packed[16];
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              f[i] = ((bits >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              f[i + 1] = ((bits >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
              packed[i >> 1] = pack_bf16(f[i], f[i + 1]);
            }
            // bias gradient of the producing layer: column sums over the tile rows
            if (op.bias_off >= 0) {
              const float cs = warp_colsum32(f, lane);
              atomicAdd(&sm.floats[op.bias_off + g * 32 + lane], cs);
            }
            uint8_t* slab = sm.slab(g >> 1);
            const int chunk0 = (g & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t off = (uint32_t)row * 128u + ((uint32_t)((chunk0 + q) ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(slab + off) =
                  make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
          }
          if (op.epi == NB_BEPI_PLAIN_SIGMA || op.epi == NB_BEPI_MASK_SIGMA) {
            // d(sigma_pre) becomes column 0 of the aux slab (an extra 16-wide K step)
            if (op.bias_off >= 0) {
              const float t = warp_sum(d_sigma_pre);
              if (lane == 0) atomicAdd(&sm.floats[op.bias_off + op.out_chunks * 64], t);
            }
            uint8_t* slab = sm.slab(4);
            *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16(d_sigma_pre, 0.f), 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(1 ^ (row & 7)) << 4)) =
                make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 2; q < 8; ++q)
              *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
                  make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tcgen05_fence_before();
        if (!last) {
          fence_proxy_async();
          mbar_arrive(sm.a_ready);
        }
        if (op.epi != NB_BEPI_NONE && op.stash_slab >= 0) {
          if (last) fence_proxy_async();
          named_bar_sync(1, kRowThreads);
          if (threadIdx.x == 0) {
            bulk_s2g(tile_stash + (size_t)op.stash_slab * NB_SLAB_BYTES, sm.slab(0),
                     (uint32_t)op.out_chunks * NB_SLAB_BYTES);
            if (op.epi == NB_BEPI_PLAIN_SIGMA || op.epi == NB_BEPI_MASK_SIGMA)
              bulk_s2g(tile_stash + (size_t)(op.stash_slab + op.out_chunks) * NB_SLAB_BYTES, sm.slab(4),
                       NB_SLAB_BYTES);
            bulk_commit();
          }
        }
      }

      // ---- gradients w.r.t. the encodings -> positions / directions -> rays ----
      if (p.want_input_grads) {
        // all MMAs of the tile are complete (acc_full of the last op); the slabs are free
        if (threadIdx.x == 0) bulk_wait_read<0>();
        named_bar_sync(1, kRowThreads);
        float* scratch = reinterpret_cast<float*>(sm.slab(0));   // [128 cols][128 rows] fp32
        auto stage_block = [&](uint32_t tmem_col, int cols, int first) {
          for (int g = 0; g < cols; g += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_lane + tmem_col + (uint32_t)g, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) scratch[(first + g + i) * NB_TILE_ROWS + row] = __uint_as_float(v[i]);
          }
        };
        if (p.pos_grad_cols > 0) stage_block(kTmemPosCol, p.pos_grad_cols, 0);
        if (p.dir_grad_cols > 0) stage_block(kTmemDirCol, p.dir_grad_cols, 64);
        tcgen05_fence_before();
        PeSample ps;
        load_sample(p.in, n, ps);
        float dx[3] = {0.f, 0.f, 0.f}, dd[3] = {0.f, 0.f, 0.f};
        if (p.pos_grad_cols > 0) {
          float dsc;
          pe_backward(p.pe_pos, sm.mask_pos, ps, [&](int col) { return scratch[col * NB_TILE_ROWS + row]; }, dx, dsc);
#pragma unroll
          for (int c = 0; c < 3; ++c) dd[c] += dsc * dx[c];
        }
        if (p.dir_grad_cols > 0) {
          PeSample pd = ps;
          pd.x[0] = ps.dir[0]; pd.x[1] = ps.dir[1]; pd.x[2] = ps.dir[2];
          float g3[3], dsc;
          pe_backward(p.pe_dir, sm.mask_dir, pd, [&](int col) { return scratch[(64 + col) * NB_TILE_ROWS + row]; }, g3, dsc);
#pragma unroll
          for (int c = 0; c < 3; ++c) dd[c] += g3[c];
        }
        if (p.in.pos != nullptr) {
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              p.d_pos[n * 3 + c] = dx[c];
              p.d_dir[n * 3 + c] = dd[c];
            }
          }
        } else {
          // x = o + t_q d  =>  dL/do += dx, dL/dd += t_q dx (+ the direction-encoding part)
          const float tq = p.in.t_mode == 0 ? ps.t0 : (ps.t0 + ps.t1) * 0.5f;
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
        named_bar_sync(1, kRowThreads);   // scratch reads done before the next tile's head write
      }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
  // bias gradients of this CTA's tiles -> flat gradient buffer
  for (int i = threadIdx.x; i < p.n_bias_floats; i += blockDim.x) {
    const int dst = p.bias_map[i];
    if (dst >= 0) atomicAdd(p.d_params + dst, sm.floats[i]);
  }
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_mlp_bwd(const void* program_host, const void* wpack_t,
                                const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                                const NbPeCfg* pe_dir_host, const float* alpha_pos,
                                const float* alpha_dir, const float* sigma, const float* rgb,
                                const float* g_sigma, const float* g_rgb, const uint32_t* masks,
                                int fwd_mask_words_per_tile, void* dy_stash, int head_sigma_col3,
                                int pos_grad_cols, int dir_grad_cols, float* d_ray_o,
                                float* d_ray_d, float* d_pos, float* d_dir, int head_bias_off,
                                int n_bias_floats, const int32_t* bias_map, float* d_params,
                                void* stream) {
  NB_CHECK_ARG(program_host && wpack_t && in_host && pe_pos_host && pe_dir_host, "mlp_bwd: null pointer");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  NB_CHECK_ARG(prog->n_ops >= 1 && prog->n_ops <= NB_MAX_OPS, "mlp_bwd: bad program (n_ops=%d)", prog->n_ops);
  NB_CHECK_ARG(sigma && rgb && masks && dy_stash && bias_map && d_params, "mlp_bwd: null buffer");
  NB_CHECK_ARG(n_bias_floats >= 0 && n_bias_floats <= (int)MlpSmem::kMaxBiasFloats && head_bias_off >= 0 &&
               head_bias_off + 4 <= n_bias_floats, "mlp_bwd: bad bias layout (%d slots)", n_bias_floats);
  NB_CHECK_ARG(pos_grad_cols >= 0 && pos_grad_cols <= 64 && pos_grad_cols % 32 == 0 &&
               dir_grad_cols >= 0 && dir_grad_cols <= 64 && dir_grad_cols % 32 == 0,
               "mlp_bwd: encoding gradient blocks must be 0, 32 or 64 columns");
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
  p.head_bias_off = head_bias_off;
  p.n_bias_floats = n_bias_floats;
  p.bias_map = bias_map;
  p.d_params = d_params;

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)MlpSmem::kBytes));
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  mlp_bwd_kernel<<<grid, kMlpThreads, MlpSmem::kBytes, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
