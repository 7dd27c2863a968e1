// Fused radiance-MLP forward on sm_100a: positions + positional encoding in registers, every
// Linear layer as tcgen05 bf16 MMAs with the 128-sample activation tile resident in shared
// memory / TMEM, weights streamed from L2 by the TMA engine (cp.async.bulk), bias/activation
// epilogues from TMEM. Replaces NerfInterpolation._compute_positions + NerfModel.forward
// (reference barf/model_interpolation.py:288-312, barf/model_interpolation_architecture.py:96-141)
// — 12 cuBLAS GEMMs + ~40 elementwise launches in the reference — by one persistent launch.
//
// CTA = 19 warps: warps 0-15 row threads (TMEM lane = tile row, four 16-column quarters of every
// slab), warp 16 issues MMAs, warp 17 streams weight images through the ring, warp 18 copies
// activation slabs to the HBM stash (training). Consecutive layers alternate between two TMEM
// accumulator buffers and hand activations over slab by slab, so the tensor pipe works on
// layer l+1 while the row threads are still in the epilogue of layer l (see mlp_kernels.cuh).
// nerfb200_mlp_fwd2 at the end of the file is the experimental two-tiles-in-flight variant.
#include <type_traits>
#include "common.cuh"
#include "mlp.h"
#include "mlp_kernels.cuh"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {
namespace {

using namespace tc;

// where a forward production goes in the per-tile stash (-1: nowhere)
struct FwdStashDst {
  const MlpFwdParams& p;
  const TileSchedule& sc;
  __device__ __forceinline__ int operator()(int ph, int s) const {
    if (ph < 0) {
      if (p.pe_pos.slab == s) return p.pe_pos.stash_slab;
      return p.pe_dir.stash_slab;
    }
    if (ph == sc.reencode_op && ((sc.reencode_mask >> s) & 1u)) return p.pe_dir.stash_slab;
    const NbOp& op = p.prog.ops[ph];
    return op.stash_slab >= 0 ? op.stash_slab + s : -1;
  }
};

// One 16-column group of an activation epilogue: bias, ReLU + sign bits, bf16 pack.
// Returns the sign half-word: bit i set <=> pre-activation i is not negative. The epilogue is
// bound by the number of instructions the four row warps of a scheduler issue per slab, so
// ReLU rides in the convert (cvt.rn.relu.bf16x2.f32) and each sign bit costs one funnel shift.
template <bool kRelu>
__device__ __forceinline__ uint32_t act_math16(const uint32_t (&v)[16], const float4 (&bias16)[4],
                                               uint32_t (&packed)[8]) {
  // two independent sign chains (one long funnel-shift chain is latency bound): chain c collects
  // elements 8c..8c+7, the first element ends up in the highest of its 8 bits
  uint32_t neg[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 bq = bias16[i >> 2];
    const float a0 = __uint_as_float(v[i]) + bq.x, a1 = __uint_as_float(v[i + 1]) + bq.y;
    const float a2 = __uint_as_float(v[i + 2]) + bq.z, a3 = __uint_as_float(v[i + 3]) + bq.w;
    if (kRelu) {
      uint32_t& n = neg[i >> 3];
      n = __funnelshift_l(__float_as_uint(a0), n, 1);
      n = __funnelshift_l(__float_as_uint(a1), n, 1);
      n = __funnelshift_l(__float_as_uint(a2), n, 1);
      n = __funnelshift_l(__float_as_uint(a3), n, 1);
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(packed[i >> 1]) : "f"(a1), "f"(a0));
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(packed[(i >> 1) + 1]) : "f"(a3), "f"(a2));
    } else {
      packed[i >> 1] = pack_bf16(a0, a1);
      packed[(i >> 1) + 1] = pack_bf16(a2, a3);
    }
  }
  if (!kRelu) return 0xffffu;
  // element i sits at bit 15 - i of the merged half-word; reverse to bit i
  const uint32_t merged = ((neg[0] & 0xffu) << 8) | (neg[1] & 0xffu);
  return (~(__brev(merged) >> 16)) & 0xffffu;
}
// the two swizzled 16-byte stores of (row, column quarter cq): the thread's two chunk addresses
// inside slab 0 are computed once per kernel, the slab index is an immediate offset
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void store_packed16(const uint32_t (&packed)[8], uint32_t sts0, uint32_t sts1, int j) {
  sts128(sts0 + (uint32_t)j * NB_SLAB_BYTES, packed[0], packed[1], packed[2], packed[3]);
  sts128(sts1 + (uint32_t)j * NB_SLAB_BYTES, packed[4], packed[5], packed[6], packed[7]);
}

__global__ void __launch_bounds__(kFwdThreads, 1)
mlp_fwd_kernel(const __grid_constant__ MlpFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem sm(smem_raw, p.prog.n_slabs, p.prog.n_stages);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  const bool training = (p.stash != nullptr);

  TileSchedule sched;
  sched.is_bwd = 0;
  sched.start_mask = 0u;
  if (p.pe_pos.slab >= 0) sched.start_mask |= 1u << p.pe_pos.slab;
  const bool dir_at_start = (p.pe_dir.slab >= 0 && p.pe_dir.encode_before_op == 0);
  if (dir_at_start) sched.start_mask |= 1u << p.pe_dir.slab;
  sched.reencode_op = (p.pe_dir.slab >= 0 && p.pe_dir.encode_before_op > 0) ? p.pe_dir.encode_before_op : -1;
  sched.reencode_mask = sched.reencode_op >= 0 ? (1u << p.pe_dir.slab) : 0u;

  if (threadIdx.x == 0) {
    sm.init_barriers();
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
    mma_issuer_loop(p.prog, sched, sm, tmem_base, n_tiles);
  } else if (warp == kStashWarp) {
    if (training && lane == 0)
      stash_loop(p.prog, sched, sm, n_tiles, p.stash, p.prog.stash_slabs_per_tile, FwdStashDst{p, sched});
  } else {
    // ---------------- row threads: encodings + epilogues ----------------
    const int row = threadIdx.x & kTileRowMask;         // tile row
    const int cq = threadIdx.x >> 7;                    // which 16-column quarter of every slab
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sts0 = smem_u32(sm.slab(0)) + (uint32_t)row * 128u + ((uint32_t)((2 * cq) ^ (row & 7)) << 4);
    uint32_t sts1 = smem_u32(sm.slab(0)) + (uint32_t)row * 128u + ((uint32_t)((2 * cq + 1) ^ (row & 7)) << 4);
    // opaque to the compiler: otherwise it rematerialises both addresses (12 instructions) per slab
    asm volatile("" : "+r"(sts0), "+r"(sts1));
    const FwdStashDst dst_of{p, sched};
    DrainBits drain;       // stash copies that must finish before a slab is rewritten
    NB_TRACE_INIT();
    uint32_t g_op = 0, tile_phase = 0;
    auto stashed = [&](int ph, uint32_t mask) {
      uint32_t out = 0u;
      for (uint32_t m = mask; m; m &= m - 1u) {
        const int s = __ffs(m) - 1;
        if (dst_of(ph, s) >= 0) out |= 1u << s;
      }
      return out;
    };
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      if (tile != (int)blockIdx.x) {   // the MMA warp has consumed every publication of the previous tile
        mbar_wait(sm.tile_done, tile_phase);
        tile_phase ^= 1u;
      }

      // ---- positions and positional encodings: quarter 0 -> position slab; quarter 1 ->
      //      direction slab when it has one of its own (otherwise it shares slab 4 and is written
      //      later). All MMAs of the previous tile have completed (its last accumulator was read).
      {
        NB_TRACE(392, threadIdx.x == 0);
        if (training) drain.acquire_mask(sm.slab_drained, sched.start_mask, lane);
        NB_TRACE(393, threadIdx.x == 0);
        {   // all four threads of a row share the work (a coordinate each + the identity columns)
          PeSample ps;
          load_sample(p.in, n, ps);
          if (p.pe_pos.slab >= 0) encode_to_slab_split(p.pe_pos, sm.mask_pos, ps, sm.slab(p.pe_pos.slab), row, cq);
          if (dir_at_start) {   // the direction encoder sees the direction as its "position"
            ps.x[0] = ps.dir[0]; ps.x[1] = ps.dir[1]; ps.x[2] = ps.dir[2];
            encode_to_slab_split(p.pe_dir, sm.mask_dir, ps, sm.slab(p.pe_dir.slab), row, cq);
          }
        }
        NB_TRACE(394, threadIdx.x == 0);
        signal_slabs(sm.slab_ready, sched.start_mask, lane);
        NB_TRACE(395, threadIdx.x == 0);
        if (training) drain.produced(stashed(-1, sched.start_mask));
      }

      for (int oi = 0; oi < p.prog.n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const float* bias = sm.floats + op.bias_off;
        const uint32_t buf = g_op & 1u;
        const uint32_t acc = tmem_lane + buf * kAccCols;
        if (oi == sched.reencode_op) {
          // The direction encoding takes over the slab the position encoding no longer needs
          // (its last reader was op oi-1, whose accumulator these threads have already seen).
          if (training) drain.acquire_mask(sm.slab_drained, sched.reencode_mask, lane);
#ifdef NB_EXP_REENCODE_SPLIT
          {
            PeSample pd;
            load_sample(p.in, n, pd);
            pd.x[0] = pd.dir[0]; pd.x[1] = pd.dir[1]; pd.x[2] = pd.dir[2];
            encode_to_slab_split(p.pe_dir, sm.mask_dir, pd, sm.slab(p.pe_dir.slab), row, cq);
          }
#else
          // one quarter does it alone (a third of the position encoding's work, mid-tile): the split
          // version makes all sixteen warps wait at its barrier and measured slower in training
          if (cq == 1) {
            PeSample pd;
            load_sample(p.in, n, pd);
            pd.x[0] = pd.dir[0]; pd.x[1] = pd.dir[1]; pd.x[2] = pd.dir[2];
            encode_to_slab(p.pe_dir, sm.mask_dir, pd, sm, row);
          }
#endif
          signal_slabs(sm.slab_ready, sched.reencode_mask, lane);
          if (training && p.pe_dir.stash_slab >= 0) drain.produced(sched.reencode_mask);
        }
        // Everything the epilogue needs from the program (constant-bank loads with a dynamic
        // index, ~100 cycles each and dependent) is fetched and pinned in registers BEFORE the
        // wait for the accumulator, so that the first TMEM load issues right after it.
        int epi_r = op.epi, oc_r = op.out_chunks, mask_word_r = op.mask_word, stash_slab_r = op.stash_slab;
        uint32_t* mask_base = p.masks + (((size_t)tile * p.prog.mask_words_per_tile + (mask_word_r < 0 ? 0 : mask_word_r) + (cq >> 1)) * NB_TILE_ROWS + row);
        asm volatile("" : "+r"(epi_r), "+r"(oc_r), "+r"(mask_word_r), "+r"(stash_slab_r), "+l"(mask_base));
        warp_mbar_wait(&sm.acc_full[buf], (g_op >> 1) & 1u, lane);
        tcgen05_fence_after();
        NB_TRACE(oi * 4 + 2, threadIdx.x == 0);
        if (fwd_stores_act(epi_r)) {
          const bool relu = (epi_r == NB_EPI_RELU || epi_r == NB_EPI_RELU_SIGMA);
          const int oc = oc_r;
          if ((epi_r == NB_EPI_LINEAR_SIGMA || epi_r == NB_EPI_RELU_SIGMA) && cq == 3) {
            // the density column sits in the first columns of the other buffer: it must be read
            // before slab 0 is published (the next op's MMAs reuse that buffer from then on)
            uint32_t e[16];
            const uint32_t col = (uint32_t)op.blocks[1].tmem_col;
            tmem_ld16(tmem_lane + (col >= kAccCols ? (buf ^ 1u) * kAccCols + (col - kAccCols) : buf * kAccCols + col), e);
            tmem_ld_wait();
            const float pre = __uint_as_float(e[0]) + bias[op.blocks[0].n];
            if (valid) p.out_sigma[n] = softplus8(pre + p.sigma_bias);
          }
          // sign bits: one 32-bit word per (row, 32-column group), this thread owns half of it
          uint16_t* mask_out = (training && relu && mask_word_r >= 0)
              ? reinterpret_cast<uint16_t*>(mask_base) + (cq & 1) : nullptr;
          const uint32_t will_stash = (training && stash_slab_r >= 0) ? 1u : 0u;
          // the TMEM load of slab j+1 is in flight during the math of slab j
          uint32_t va[16], vb[16], packed[8];
          const float* bias_q = bias + 16 * cq;
          const uint32_t acc_q = acc + (uint32_t)(16 * cq);
          tmem_ld16(acc_q, va);
          // A slab may be rewritten once its stash copy has drained; checked slab by slab (while the
          // slab's TMEM load is in flight): waiting up front for the newest copy (the last slab of the previous op) held
          // every epilogue back by ~900 cycles.
          NB_TRACE(399, threadIdx.x == 0 && oi == 3);
          uint32_t sign_bits[4];
          auto finish = [&](int j, uint32_t bits) {
            const bool tr = (threadIdx.x == 0 || threadIdx.x == 480) && oi == 3;   // trace builds only
            const int ts = 400 + (threadIdx.x == 0 ? 0 : 24) + 6 * j;
            NB_TRACE(ts + 1, tr);
            store_packed16(packed, sts0, sts1, j);
            sign_bits[j] = bits;    // stored after the last slab has been published: a global store in
                                    // flight makes the proxy fence (MEMBAR.ALL.CTA) wait for its ack
            NB_TRACE(ts + 2, tr);
            writer_proxy_fence();
            NB_TRACE(ts + 3, tr);
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.slab_ready[j]);
            NB_TRACE(ts + 4, tr);
            if (will_stash) { drain.pending |= 1u << j; drain.last = j; }
          };
          float4 bq[4];
          auto load_bias = [&](const float* b16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) bq[i] = *reinterpret_cast<const float4*>(b16 + 4 * i);
          };
          auto run = [&](auto relu_tag) {
            constexpr bool kRelu = decltype(relu_tag)::value;
#pragma unroll
            for (int j = 0; j < 4; j += 2) {
              if (j < oc) {
                drain.acquire(sm.slab_drained, j, lane);       // both overlap the TMEM load in flight
                load_bias(bias_q + 64 * j);
                tmem_ld_wait16(va);
                NB_TRACE(400 + (threadIdx.x == 0 ? 0 : 24) + 6 * j, (threadIdx.x == 0 || threadIdx.x == 480) && oi == 3);
                if (j + 1 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 1)), vb);
                finish(j, act_math16<kRelu>(va, bq, packed));
              }
              if (j + 1 < oc) {
                drain.acquire(sm.slab_drained, j + 1, lane);
                load_bias(bias_q + 64 * (j + 1));
                tmem_ld_wait16(vb);
                NB_TRACE(400 + (threadIdx.x == 0 ? 0 : 24) + 6 * (j + 1), (threadIdx.x == 0 || threadIdx.x == 480) && oi == 3);
                if (j + 2 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 2)), va);
                finish(j + 1, act_math16<kRelu>(vb, bq, packed));
              }
            }
          };
          if (relu) run(std::true_type{}); else run(std::false_type{});
          if (mask_out != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < oc) mask_out[(size_t)(2 * j) * NB_TILE_ROWS * 2] = (uint16_t)sign_bits[j];
          }
        } else if (cq == 0) {
          // NB_EPI_RGB / NB_EPI_RGB_SIGMA: first 16 accumulator columns hold the outputs
          uint32_t v[16];
          tmem_ld16(acc, v);
          tmem_ld_wait();
          if (valid) {
            p.out_rgb[n * 3 + 0] = sigmoidf(__uint_as_float(v[0]) + bias[0]);
            p.out_rgb[n * 3 + 1] = sigmoidf(__uint_as_float(v[1]) + bias[1]);
            p.out_rgb[n * 3 + 2] = sigmoidf(__uint_as_float(v[2]) + bias[2]);
            if (epi_r == NB_EPI_RGB_SIGMA)
              p.out_sigma[n] = softplus8(__uint_as_float(v[3]) + bias[3] + p.sigma_bias);
          }
        }
        // this accumulator buffer has been read out
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tmem_free[buf]);
        NB_TRACE(oi * 4 + 3, threadIdx.x == 0);
        ++g_op;
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
}


// =================================================================================================
// Two tiles in flight (mlp_fwd2).  Reading a 128 x 256 fp32 accumulator back costs as many cycles
// (128 KB at the 64 B/clk of the TMEM read port) as the MMAs that produce it, and inside ONE tile
// the two can only overlap within a layer, not across the layer boundary. This kernel keeps two
// 128-sample tiles per CTA, one 256-column accumulator each: while the row warps run the epilogue
// of layer l of tile A, the tensor pipe runs layer l of tile B, and so on — no slab-granular
// hand-off, one proxy fence and one barrier arrival per warp, layer and tile. Ten activation
// slabs leave 48 KB for the weight ring, so the weights stream as one 8 KB image per MMA
// ([256][16] bf16, SWIZZLE_32B, packed by mlp_pack with img_rows > 0). The density column of the
// "extra" layer has no room in TMEM: the epilogue of the layer in front evaluates it as a dot
// product of the bf16 activations it has in registers with the (bf16-rounded) weight row.
// Stash / sign-bit layout in HBM is the one of mlp_fwd_kernel: the backward kernels do not care.
// =================================================================================================
constexpr int kT2Slabs = 5;
#ifndef NB_T2_STAGES
#define NB_T2_STAGES 3
#endif
#ifndef NB_T2_KSTEPS
#define NB_T2_KSTEPS 2
#endif
constexpr int kT2Stages = NB_T2_STAGES;
constexpr int kT2KPerStage = NB_T2_KSTEPS;            // weight images (K steps of one chunk) per ring stage
constexpr uint32_t kT2ImageBytes = 8192;
constexpr uint32_t kT2StageBytes = kT2KPerStage * kT2ImageBytes;

struct T2Smem {
  static constexpr uint32_t kCtrlBytes = 512;
  __host__ __device__ static constexpr uint32_t bytes() {
    return 2u * kT2Slabs * NB_SLAB_BYTES + kT2Stages * kT2StageBytes + kCtrlBytes +
           MlpSmem::kMaxBiasFloats * 4u + 2u * NB_TILE_ROWS * 4u;
  }
  uint8_t* base;
  uint8_t* ring_base;
  uint64_t* full;       // [kT2Stages]
  uint64_t* empty;      // [kT2Stages]
  uint64_t* in_ready;   // [2] row warps -> MMA / stash warp: the tile's slabs for the next op are published
  uint64_t* acc_full;   // [2] MMA warp -> row warps
  uint64_t* drained;    // [2] stash warp -> row warps: the copies of the tile's last publication have read shared memory
  uint32_t* tmem_ptr;
  float* mask_pos;
  float* mask_dir;
  float* floats;        // packed biases (+ the density weight row)
  float* dens;          // [2][128] density pre-activation partial sums
  __device__ explicit T2Smem(uint8_t* b) : base(b) {
    ring_base = b + 2u * kT2Slabs * NB_SLAB_BYTES;
    uint8_t* c = ring_base + kT2Stages * kT2StageBytes;
    full = reinterpret_cast<uint64_t*>(c);
    empty = full + kT2Stages;
    in_ready = empty + kT2Stages;
    acc_full = in_ready + 2;
    drained = acc_full + 2;
    tmem_ptr = reinterpret_cast<uint32_t*>(drained + 2);
    mask_pos = reinterpret_cast<float*>(c + 256);
    mask_dir = mask_pos + kMaxLevels;
    floats = reinterpret_cast<float*>(c + kCtrlBytes);
    dens = floats + MlpSmem::kMaxBiasFloats;
  }
  __device__ uint8_t* slab(int t, int s) const { return base + (uint32_t)(t * kT2Slabs + s) * NB_SLAB_BYTES; }
  __device__ uint8_t* ring(int s) const { return ring_base + (uint32_t)s * kT2StageBytes; }
};
static_assert(T2Smem::bytes() <= 227 * 1024, "shared memory budget of the two-tile kernel");
static_assert((2 * kT2Stages + 6) * 8 + 4 <= 256, "control block layout");

struct MlpFwd2Params {
  MlpFwdParams f;
  int density_w_off;     // float offset of the density weight row in the packed biases, -1: none
};

// slabs of tile-phase `ph` (-1: encodings) that have a place in the stash, as a bit mask, and where
__device__ __forceinline__ uint32_t t2_stash_mask(const MlpFwdParams& p, int reencode_op, int ph) {
  if (ph < 0) return (p.pe_pos.slab >= 0 && p.pe_pos.stash_slab >= 0) ? (1u << p.pe_pos.slab) : 0u;
  uint32_t m = 0u;
  const NbOp& op = p.prog.ops[ph];
  if (fwd_stores_act(op.epi) && op.stash_slab >= 0) m |= (1u << op.out_chunks) - 1u;
  if (ph == reencode_op && p.pe_dir.stash_slab >= 0) m |= 1u << p.pe_dir.slab;
  return m;
}
__device__ __forceinline__ int t2_stash_dst(const MlpFwdParams& p, int reencode_op, int ph, int s) {
  if (ph < 0) return p.pe_pos.stash_slab;
  if (ph == reencode_op && s == p.pe_dir.slab) return p.pe_dir.stash_slab;
  return p.prog.ops[ph].stash_slab + s;
}

__global__ void __launch_bounds__(kFwdThreads, 1)
mlp_fwd2_kernel(const __grid_constant__ MlpFwd2Params pp) {
  const MlpFwdParams& p = pp.f;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  T2Smem sm(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  const int n_pairs = (n_tiles + 1) / 2;
  const int n_ops = p.prog.n_ops;
  const bool training = (p.stash != nullptr);
  const int reencode_op = (p.pe_dir.slab >= 0 && p.pe_dir.encode_before_op > 0) ? p.pe_dir.encode_before_op : -1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT2Stages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&sm.in_ready[t], kRowWarps);
      mbar_init(&sm.acc_full[t], 1);
      mbar_init(&sm.drained[t], 1);
    }
    fence_barrier_init();
    pe_fill_mask(p.pe_pos, p.alpha_pos, sm.mask_pos);
    pe_fill_mask(p.pe_dir, p.alpha_dir, sm.mask_dir);
  }
  for (int i = threadIdx.x; i < p.n_bias_floats; i += blockDim.x) sm.floats[i] = p.bias[i];
  for (int i = threadIdx.x; i < 2 * NB_TILE_ROWS; i += blockDim.x) sm.dens[i] = 0.f;
  if (warp == kMmaWarp) tmem_alloc(sm.tmem_ptr, kTmemCols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarp) {
    // ---------------- weight producer: one image per MMA, in the MMA warp's order ----------------
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int oi = 0; oi < n_ops; ++oi) {
          const NbOp& op = p.prog.ops[oi];
          for (int t = 0; t < 2; ++t) {
            for (int c = 0; c < op.n_chunks; ++c) {
              if (!(op.blk_mask[c] & 1)) continue;      // density images: not used by this kernel
              const uint32_t bytes = (uint32_t)op.w_rows[c] * 32u;
              const uint8_t* src = p.wpack + (size_t)op.w_off[c] * 1024u;
              for (int k = 0; k < op.k16[c]; k += kT2KPerStage) {
                const int nk = op.k16[c] - k < kT2KPerStage ? op.k16[c] - k : kT2KPerStage;
                mbar_wait(&sm.empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&sm.full[stage], bytes * (uint32_t)nk);
                for (int kk = 0; kk < nk; ++kk)
                  bulk_g2s(sm.ring(stage) + (uint32_t)kk * kT2ImageBytes, src + (size_t)(k + kk) * bytes, bytes, &sm.full[stage]);
                if (++stage == kT2Stages) { stage = 0; phase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ---------------- MMA issuer (whole warp converged, one elected lane issues) ----------------
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t desc_hi_a = (uint32_t)(umma_desc(0u, 0u, 1024u) >> 32);
    const uint32_t desc_hi_b = (uint32_t)(umma_desc_kmajor_sw32(0u) >> 32);
    const uint32_t slab0 = smem_u32(sm.slab(0, 0)) >> 4;
    const uint32_t ring0 = smem_u32(sm.ring(0)) >> 4;
    const bool elected = elect_one();
    uint32_t stage = 0, phase = 0;
    uint32_t n_in[2] = {0u, 0u};
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int oi = 0; oi < n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const uint32_t idesc = umma_idesc(NB_TILE_ROWS, op.blocks[0].n, false, false);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // the inputs of op oi of tile t are published, and its accumulator has been read out
          mbar_wait(&sm.in_ready[t], n_in[t] & 1u);
          ++n_in[t];
          tcgen05_fence_after();
          const uint32_t d_col = tb + (uint32_t)t * kAccCols;
          uint32_t acc = 0u;
          for (int c = 0; c < op.n_chunks; ++c) {
            if (!(op.blk_mask[c] & 1)) continue;
            const uint32_t a_lo = slab0 + (uint32_t)(t * kT2Slabs + op.a_src[c]) * (NB_SLAB_BYTES >> 4);
            const int k16 = op.k16[c];
            for (int k = 0; k < k16; k += kT2KPerStage) {
              mbar_wait(&sm.full[stage], phase);
              tcgen05_fence_after();
#pragma unroll
              for (int kk = 0; kk < kT2KPerStage; ++kk) {
                if (k + kk < k16) {
                  const uint64_t adesc = ((uint64_t)desc_hi_a << 32) | (uint64_t)(a_lo + 2u * (uint32_t)(k + kk));
                  const uint64_t bdesc = ((uint64_t)desc_hi_b << 32) |
                                         (uint64_t)(ring0 + stage * (kT2StageBytes >> 4) + (uint32_t)kk * (kT2ImageBytes >> 4));
                  if (elected) umma(d_col, adesc, bdesc, idesc, acc);
                  acc = 1u;
                }
              }
              if (elected) umma_commit(&sm.empty[stage]);
              if (++stage == kT2Stages) { stage = 0; phase ^= 1u; }
            }
          }
          if (elected) umma_commit(&sm.acc_full[t]);
          __syncwarp();
        }
      }
      // The arrival after the last op's epilogue has no MMA behind it, but it must be CONSUMED, not
      // skipped: a parity wait issued one completion early returns at once (the phase "before the
      // previous one" has the parity that is being waited for).
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&sm.in_ready[t], n_in[t] & 1u);
        ++n_in[t];
      }
    }
  } else if (warp == kStashWarp) {
    // ---------------- stash copies: every publication, in publication order ----------------
    if (training && lane == 0) {
      uint32_t n_in[2] = {0u, 0u};
      for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        for (int ph = -1; ph < n_ops; ++ph) {
          const uint32_t mask = t2_stash_mask(p, reencode_op, ph);
          for (int t = 0; t < 2; ++t) {
            const int tile = 2 * pair + t;
            mbar_wait(&sm.in_ready[t], n_in[t] & 1u);
            ++n_in[t];
            if (mask == 0u) continue;
            if (tile < n_tiles) {
              uint8_t* tile_stash = p.stash + (size_t)tile * p.prog.stash_slabs_per_tile * NB_SLAB_BYTES;
              for (uint32_t m = mask; m; m &= m - 1u) {
                const int s = __ffs(m) - 1;
#ifndef NB_EXP_T2_NOCOPY
                bulk_s2g(tile_stash + (size_t)t2_stash_dst(p, reencode_op, ph, s) * NB_SLAB_BYTES, sm.slab(t, s), NB_SLAB_BYTES);
#endif
              }
              bulk_commit();
              bulk_wait_read<0>();
            }
            mbar_arrive(&sm.drained[t]);
          }
        }
      }
      bulk_wait_all<0>();
    }
  } else {
    // ---------------- row threads: encodings + epilogues ----------------
    const int row = threadIdx.x & kTileRowMask;
    const int cq = threadIdx.x >> 7;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sts_base = smem_u32(sm.slab(0, 0)) + (uint32_t)row * 128u + ((uint32_t)((2 * cq) ^ (row & 7)) << 4);
    asm volatile("" : "+r"(sts_base));
    uint32_t n_acc[2] = {0u, 0u}, n_dr[2] = {0u, 0u};
    bool dr_pending[2] = {false, false};
    // before a tile's slabs are rewritten the stash copies of its previous publication must have drained
    auto wait_drained = [&](int t) {
      if (dr_pending[t]) {
        mbar_wait(&sm.drained[t], n_dr[t] & 1u);
        ++n_dr[t];
        dr_pending[t] = false;
      }
    };
    auto publish = [&](int t, bool wrote) {
      if (wrote) fence_proxy_async();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.in_ready[t]);
    };
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {      // ---- phase -1: positions and their encoding ----
        const long long n_raw = (long long)(2 * pair + t) * NB_TILE_ROWS + row;
        const long long n = n_raw < p.N ? n_raw : (long long)p.N - 1;
        wait_drained(t);
        if (p.pe_pos.slab >= 0) {   // all four threads of a row share the work
          PeSample ps;
          load_sample(p.in, n, ps);
          encode_to_slab_split(p.pe_pos, sm.mask_pos, ps, sm.slab(t, p.pe_pos.slab), row, cq);
        }
        publish(t, true);
        dr_pending[t] = training && t2_stash_mask(p, reencode_op, -1) != 0u;
      }
      for (int oi = 0; oi < n_ops; ++oi) {
        const NbOp& op = p.prog.ops[oi];
        const int epi = op.epi, oc = op.out_chunks;
        const bool stores_act = fwd_stores_act(epi);
        const bool relu = (epi == NB_EPI_RELU || epi == NB_EPI_RELU_SIGMA);
        const bool sigma_here = (epi == NB_EPI_LINEAR_SIGMA || epi == NB_EPI_RELU_SIGMA);
        const bool density_next = pp.density_w_off >= 0 && oi + 1 < n_ops &&
                                  (p.prog.ops[oi + 1].epi == NB_EPI_LINEAR_SIGMA || p.prog.ops[oi + 1].epi == NB_EPI_RELU_SIGMA);
        const float* bias = sm.floats + op.bias_off;
        const uint32_t phase_stash = training ? t2_stash_mask(p, reencode_op, oi) : 0u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int tile = 2 * pair + t;
          const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
          const bool valid = n_raw < p.N;
          const long long n = valid ? n_raw : (long long)p.N - 1;
          wait_drained(t);
          bool wrote = false;
          if (oi == reencode_op) {   // the direction encoding takes over the encoding slab (its last reader is done)
            if (cq == 1) {
              PeSample pd;
              load_sample(p.in, n, pd);
              pd.x[0] = pd.dir[0]; pd.x[1] = pd.dir[1]; pd.x[2] = pd.dir[2];
              encode_to_slab_at(p.pe_dir, sm.mask_dir, pd, sm.slab(t, p.pe_dir.slab), row);
            }
            wrote = true;
          }
          mbar_wait(&sm.acc_full[t], n_acc[t] & 1u);
          ++n_acc[t];
          tcgen05_fence_after();
          const uint32_t acc = tmem_lane + (uint32_t)t * kAccCols;
          if (stores_act) {
            wrote = true;
            if (sigma_here && cq == 3) {
              // the density pre-activation was summed up by the epilogue of the layer in front
              float* dsum = sm.dens + t * NB_TILE_ROWS + row;
              const float pre = *dsum + bias[op.blocks[0].n];
              *dsum = 0.f;
              if (valid) p.out_sigma[n] = softplus8(pre + p.sigma_bias);
            }
            uint16_t* mask_out = (training && relu && op.mask_word >= 0 && tile < n_tiles)
                ? reinterpret_cast<uint16_t*>(p.masks + (((size_t)tile * p.prog.mask_words_per_tile + op.mask_word + (cq >> 1)) * NB_TILE_ROWS + row)) + (cq & 1)
                : nullptr;
            uint32_t va[16], vb[16], packed[8];
            const float* bias_q = bias + 16 * cq;
            const uint32_t acc_q = acc + (uint32_t)(16 * cq);
            const uint32_t sts_t = sts_base + (uint32_t)(t * kT2Slabs) * NB_SLAB_BYTES;
            const float* wd = sm.floats + (pp.density_w_off >= 0 ? pp.density_w_off : 0) + 16 * cq;
            float dpart = 0.f;
            uint32_t sign_bits[4];
            float4 bq[4];
            auto load_bias = [&](const float* b16) {
#pragma unroll
              for (int i = 0; i < 4; ++i) bq[i] = *reinterpret_cast<const float4*>(b16 + 4 * i);
            };
            auto finish = [&](int j, uint32_t bits) {
              sts128(sts_t + (uint32_t)j * NB_SLAB_BYTES, packed[0], packed[1], packed[2], packed[3]);
              sts128((sts_t ^ 16u) + (uint32_t)j * NB_SLAB_BYTES, packed[4], packed[5], packed[6], packed[7]);
              sign_bits[j] = bits;
              if (density_next) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float w0 = __bfloat162float(__float2bfloat16_rn(wd[64 * j + 2 * i]));
                  const float w1 = __bfloat162float(__float2bfloat16_rn(wd[64 * j + 2 * i + 1]));
                  dpart = fmaf(__uint_as_float(packed[i] << 16), w0, dpart);
                  dpart = fmaf(__uint_as_float(packed[i] & 0xffff0000u), w1, dpart);
                }
              }
            };
            auto run = [&](auto relu_tag) {
              constexpr bool kRelu = decltype(relu_tag)::value;
#pragma unroll
              for (int j = 0; j < 4; j += 2) {
                if (j < oc) {
                  load_bias(bias_q + 64 * j);
                  tmem_ld_wait16(va);
                  if (j + 1 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 1)), vb);
                  finish(j, act_math16<kRelu>(va, bq, packed));
                }
                if (j + 1 < oc) {
                  load_bias(bias_q + 64 * (j + 1));
                  tmem_ld_wait16(vb);
                  if (j + 2 < oc) tmem_ld16(acc_q + (uint32_t)(64 * (j + 2)), va);
                  finish(j + 1, act_math16<kRelu>(vb, bq, packed));
                }
              }
            };
            tmem_ld16(acc_q, va);
            if (relu) run(std::true_type{}); else run(std::false_type{});
            if (density_next) atomicAdd(sm.dens + t * NB_TILE_ROWS + row, dpart);
            publish(t, true);
            if (mask_out != nullptr) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < oc) mask_out[(size_t)(2 * j) * NB_TILE_ROWS * 2] = (uint16_t)sign_bits[j];
            }
          } else {
            if (cq == 0) {   // NB_EPI_RGB / NB_EPI_RGB_SIGMA: the first accumulator columns hold the outputs
              uint32_t v[16];
              tmem_ld16(acc, v);
              tmem_ld_wait();
              if (valid) {
                p.out_rgb[n * 3 + 0] = sigmoidf(__uint_as_float(v[0]) + bias[0]);
                p.out_rgb[n * 3 + 1] = sigmoidf(__uint_as_float(v[1]) + bias[1]);
                p.out_rgb[n * 3 + 2] = sigmoidf(__uint_as_float(v[2]) + bias[2]);
                if (epi == NB_EPI_RGB_SIGMA)
                  p.out_sigma[n] = softplus8(__uint_as_float(v[3]) + bias[3] + p.sigma_bias);
              }
            }
            publish(t, wrote);
          }
          dr_pending[t] = phase_stash != 0u;
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

extern "C" int nerfb200_mlp_workspace_bytes(const void* program_host, long long n_samples,
                                            long long* stash_bytes, long long* mask_bytes) {
  NB_CHECK_ARG(program_host && n_samples >= 0, "mlp_workspace_bytes: bad arguments");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  const long long n_tiles = (n_samples + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  if (stash_bytes) *stash_bytes = n_tiles * prog->stash_slabs_per_tile * (long long)NB_SLAB_BYTES;
  if (mask_bytes) *mask_bytes = n_tiles * prog->mask_words_per_tile * (long long)NB_TILE_ROWS * 4;
  return NERFB200_OK;
}

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
  mlp_fwd_kernel<<<grid, kFwdThreads, MlpSmem::bytes(prog->n_slabs, prog->n_stages), (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}

extern "C" int nerfb200_mlp_fwd2(const void* program_host, const void* wpack_k16, const float* bias,
                                 const NbMlpInputs* in_host, const NbPeCfg* pe_pos_host,
                                 const NbPeCfg* pe_dir_host, const float* alpha_pos,
                                 const float* alpha_dir, float sigma_bias, float* out_sigma,
                                 float* out_rgb, void* stash, uint32_t* masks, int n_bias_floats,
                                 int density_w_off, void* stream) {
  NB_CHECK_ARG(program_host && wpack_k16 && bias && in_host && pe_pos_host && pe_dir_host,
               "mlp_fwd2: null pointer");
  const NbProgram* prog = reinterpret_cast<const NbProgram*>(program_host);
  NB_CHECK_ARG(prog->n_ops >= 1 && prog->n_ops <= NB_MAX_OPS, "mlp_fwd2: bad program (n_ops=%d)", prog->n_ops);
  NB_CHECK_ARG(in_host->N >= 0 && in_host->S >= 1, "mlp_fwd2: bad shape N=%lld S=%d", (long long)in_host->N, in_host->S);
  NB_CHECK_ARG(out_sigma && out_rgb, "mlp_fwd2: null output");
  NB_CHECK_ARG((stash == nullptr) == (masks == nullptr), "mlp_fwd2: stash and masks go together");
  NB_CHECK_ARG(n_bias_floats >= 1 && n_bias_floats <= (int)MlpSmem::kMaxBiasFloats,
               "mlp_fwd2: %d packed bias slots (max %d)", n_bias_floats, (int)MlpSmem::kMaxBiasFloats);
  NB_CHECK_ARG(density_w_off < 0 || density_w_off + 256 <= n_bias_floats, "mlp_fwd2: density weights out of range");
  NB_CHECK_ARG(prog->n_slabs == 5, "mlp_fwd2: needs the 5-slab program shape (shared encoding slab)");
  NB_CHECK_ARG(pe_pos_host->slab < kT2Slabs && pe_dir_host->slab < kT2Slabs, "mlp_fwd2: encoding slab out of range");
  for (int i = 0; i < prog->n_ops; ++i) {
    const NbOp& op = prog->ops[i];
    const bool sigma = op.epi == NB_EPI_LINEAR_SIGMA || op.epi == NB_EPI_RELU_SIGMA;
    NB_CHECK_ARG(!sigma || density_w_off >= 0, "mlp_fwd2: op %d has a density block but no density weights", i);
    NB_CHECK_ARG(op.blocks[0].tmem_col == 0 && op.blocks[0].row0 == 0 && op.blocks[0].n >= 16 && op.blocks[0].n <= 256,
                 "mlp_fwd2: op %d: unsupported main block", i);
    for (int c = 0; c < op.n_chunks; ++c) {
      if (!(op.blk_mask[c] & 1)) continue;
      NB_CHECK_ARG(op.n_sub[c] == 1 && op.a_src[c] >= 0 && op.a_src[c] < kT2Slabs && op.k16[c] >= 1 && op.k16[c] <= 4 &&
                   op.w_rows[c] == op.blocks[0].n && op.w_rows[c] * 32 <= (int)kT2ImageBytes,
                   "mlp_fwd2: op %d chunk %d is not a plain main chunk", i, c);
    }
  }
  if (in_host->N == 0) return NERFB200_OK;

  MlpFwd2Params pp;
  MlpFwdParams& p = pp.f;
  p.prog = *prog;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack_k16);
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
  pp.density_w_off = density_w_off;

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(mlp_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T2Smem::bytes()));
    configured = true;
  }
  const int n_pairs = (ceil_div(p.N, NB_TILE_ROWS) + 1) / 2;
  const int grid = n_pairs < sm_count() ? n_pairs : sm_count();
  mlp_fwd2_kernel<<<grid, kFwdThreads, T2Smem::bytes(), (cudaStream_t)stream>>>(pp);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
