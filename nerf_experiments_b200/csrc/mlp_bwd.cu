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
constexpr int kHelperWarp0 = 10;         // warps 10 and 11: bias-gradient column sums
constexpr int kBwdThreads = 384;

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__global__ void __launch_bounds__(kBwdThreads, 1)
mlp_bwd_kernel(const __grid_constant__ MlpBwdParams p) {
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
    mbar_init(sm.epi_done, kRowThreads);
    mbar_init(sm.helper_done, 2);
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
    if (lane == 0) weight_producer_loop(p.prog, p.wpack, sm, n_tiles);
  } else if (warp == kMmaWarp) {
    mma_issuer_loop(p.prog, sm, tmem_base, n_tiles);
  } else if (warp >= kHelperWarp0) {
    // ---------------- bias gradients: column sums of every dY tile, off the critical path ---
    // Helper warp h owns slabs 2h and 2h+1. Lane = (slab sl, row half rh, physical chunk pc) sums
    // its 16-byte chunk position over 64 rows; the swizzle (logical chunk = pc ^ (row & 7)) is
    // undone with xor-shuffles, so every column ends with exactly one owner lane: no atomics.
    const int h = warp - kHelperWarp0;
    const int sl = lane >> 4, rh = (lane >> 3) & 1, pc = lane & 7;
    const int s_idx = 2 * h + sl;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        if (op.epi == NB_BEPI_NONE) continue;
        mbar_wait(sm.epi_done, ph);
        ph ^= 1u;
        if (op.bias_off >= 0 && 2 * h < op.out_chunks) {
          float acc[8][8];
#pragma unroll
          for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
          const bool live = s_idx < op.out_chunks;
          const uint8_t* base = sm.slab(live ? s_idx : 0) + (uint32_t)(rh * 64) * 128u + (uint32_t)pc * 16u;
#pragma unroll 1
          for (int r8 = 0; r8 < 8; ++r8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {   // row & 7 == q
              const uint4 v = *reinterpret_cast<const uint4*>(base + (uint32_t)(r8 * 8 + q) * 128u);
              acc[q][0] += bf16_lo(v.x); acc[q][1] += bf16_hi(v.x);
              acc[q][2] += bf16_lo(v.y); acc[q][3] += bf16_hi(v.y);
              acc[q][4] += bf16_lo(v.z); acc[q][5] += bf16_hi(v.z);
              acc[q][6] += bf16_lo(v.w); acc[q][7] += bf16_hi(v.w);
            }
          }
          // lane (.., pc) holds in acc[q] the partial sums of logical chunk pc ^ q: lane c gathers
          // acc[q] from lane c ^ q
          float tot[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) tot[e] = acc[0][e];
#pragma unroll
          for (int q = 1; q < 8; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) tot[e] += __shfl_xor_sync(0xffffffffu, acc[q][e], q);
#pragma unroll
          for (int e = 0; e < 8; ++e) tot[e] += __shfl_xor_sync(0xffffffffu, tot[e], 8);   // other row half
          if (live && rh == 0) {
            float* dst = sm.floats + op.bias_off + s_idx * 64 + pc * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[e] += tot[e];
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.helper_done);
      }
    }
  } else {
    const int row = threadIdx.x & (kHalfThreads - 1);
    const int half = threadIdx.x >> 7;
    const bool leader = (row == 0);
    const int bar_id = 1 + half;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc_phase = 0, helper_phase = 0;
    bool helper_pending = false;   // a helper pass may still be reading the act slabs
    StashQueue sq;
    auto wait_helper = [&]() {
      if (helper_pending) {
        mbar_wait(sm.helper_done, helper_phase);
        helper_phase ^= 1u;
        helper_pending = false;
      }
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      uint8_t* tile_stash = p.dy_stash + (size_t)tile * p.prog.stash_slabs_per_tile * NB_SLAB_BYTES;
      const uint32_t* tile_masks = p.masks + (size_t)tile * p.fwd_mask_words_per_tile * NB_TILE_ROWS + row;

      // earlier stash copies / helper passes must be done reading the slabs before the rewrite
      if (leader) sq.wait_all();
      named_bar_sync(3, kRowThreads);
      wait_helper();

      // ---- head: gradients w.r.t. the pre-activations of the output layer (half 0) ----
      const float sg = p.sigma[n];
      float d_sigma_pre = 0.f;
      if (valid && p.g_sigma != nullptr)
        d_sigma_pre = p.g_sigma[n] * (sg > 8.f ? 1.f : (1.f - __expf(-sg)));
      if (half == 0) {
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
        // columns 0..15 hold the head gradient (only k16 = 1 is consumed by the MMA); the rest of
        // the row is cleared because wgrad consumes the stashed slab as a full 64-column operand
        *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16(d4[0], d4[1]), pack_bf16(d4[2], d4[3]), 0u, 0u);
#pragma unroll
        for (int q = 1; q < 8; ++q)
          *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
              make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(sm.a_ready);
      if (half == 0) {
        named_bar_sync(bar_id, kHalfThreads);
        if (leader) {
          sq.begin_batch();
          sq.push(tile_stash, sm.slab(0), NB_SLAB_BYTES);   // head dY = stash slab 0
        }
      }

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const bool last = (oi == p.prog.n_ops - 1);
        const bool stores = (op.epi != NB_BEPI_NONE);
        const bool masked = (op.epi == NB_BEPI_MASK || op.epi == NB_BEPI_MASK_SIGMA);
        const bool with_sigma = (op.epi == NB_BEPI_PLAIN_SIGMA || op.epi == NB_BEPI_MASK_SIGMA);
        const int c_mid = (op.out_chunks + 1) >> 1;
        const int c_begin = half == 0 ? 0 : c_mid;
        const int c_end = half == 0 ? c_mid : op.out_chunks;
        // ReLU sign bits of the producing layer: fetched before the accumulator is waited for
        uint32_t mbits[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = 2 * c_begin + j;
          mbits[j] = (masked && g < 2 * c_end) ? __ldg(tile_masks + (size_t)(op.mask_word + g) * NB_TILE_ROWS)
                                               : 0xffffffffu;
        }
        mbar_wait(sm.acc_full, acc_phase);
        acc_phase ^= 1u;
        tcgen05_fence_after();
        NB_TRACE(oi * 4 + 2, threadIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x));
        if (stores) {
          wait_helper();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int g = 2 * c_begin + j;
            if (g < 2 * c_end) {
              uint32_t v[32];
              tmem_ld32(tmem_lane + (uint32_t)(g * 32), v);
              if ((g & 1) == 0) {
                if (leader) sq.wait_slab((g >> 1) - c_begin);
                named_bar_sync(bar_id, kHalfThreads);
              }
              const uint32_t bits = mbits[j];
              tmem_ld_wait();
              uint32_t packed[16];
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float a = ((bits >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
                const float b = ((bits >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
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
            }
          }
          if (with_sigma && half == 1) {
            // d(sigma_pre) becomes column 0 of the aux slab (an extra 16-wide K step)
            if (op.bias_off >= 0) {
              const float t = warp_sum(d_sigma_pre);
              if (lane == 0) atomicAdd(&sm.floats[op.bias_off + op.out_chunks * 64], t);
            }
            uint8_t* slab = sm.slab(4);
            *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(0 ^ (row & 7)) << 4)) =
                make_uint4(pack_bf16(d_sigma_pre, 0.f), 0u, 0u, 0u);
#pragma unroll
            for (int q = 1; q < 8; ++q)
              *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
                  make_uint4(0u, 0u, 0u, 0u);
          }
        }
        tcgen05_fence_before();
        NB_TRACE(oi * 4 + 3, threadIdx.x == 0 && tile == (int)(blockIdx.x + gridDim.x));
        if (stores) {
          mbar_arrive(sm.epi_done);     // release: the helper warps may read the slabs
          helper_pending = true;
        }
        if (!last) {
          fence_proxy_async();
          mbar_arrive(sm.a_ready);
        }
        if (stores && op.stash_slab >= 0) {
          if (last) fence_proxy_async();
          named_bar_sync(bar_id, kHalfThreads);
          if (leader) {
            sq.begin_batch();
            for (int c = c_begin; c < c_end; ++c)
              sq.push(tile_stash + (size_t)(op.stash_slab + c) * NB_SLAB_BYTES, sm.slab(c), NB_SLAB_BYTES);
            if (with_sigma && half == 1)
              sq.push(tile_stash + (size_t)(op.stash_slab + op.out_chunks) * NB_SLAB_BYTES, sm.slab(4), NB_SLAB_BYTES);
          }
        }
      }

      // ---- gradients w.r.t. the encodings -> positions / directions -> rays ----
      if (p.want_input_grads) {
        // all MMAs of the tile are complete (acc_full of the last op); the slabs are free once the
        // stash copies (both halves') and the helper pass have drained
        if (leader) sq.wait_all();
        named_bar_sync(3, kRowThreads);
        wait_helper();
      }
      if (p.want_input_grads && half == 0) {
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
      }
    }
    wait_helper();
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
