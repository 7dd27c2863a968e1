// Shared pieces of the fused MLP kernels (forward, backward-data): shared-memory carve-up,
// kernel parameter blocks, sample loading and small math helpers.
#pragma once
#include "common.cuh"
#include "mlp.h"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {

// Warp roles. Warps 0-15 are the "row threads": thread t owns tile row (t & 127) — the TMEM lane
// quarter of a warp is fixed by (warp & 3) — and the 16-column quarter (t >> 7) of every
// 64-column slab of an accumulator, so a slab of the next layer's input is complete (and its
// MMAs can start) as soon as the sixteen warps have passed it. Four row warps per scheduler
// hide the TMEM / shared-memory / barrier latencies of one another; the epilogue is issue-bound
// (one FADD, one funnel shift, half a convert and half a max per element). Warp 16 issues the
// MMAs, warp 17 streams the weight images, warp 18 copies finished slabs to the HBM stash.
constexpr int kRowWarps = 16;
constexpr int kRowThreads = 512;
constexpr int kTileRowMask = 127;
constexpr int kMmaWarp = 16;
constexpr int kProducerWarp = 17;
constexpr int kStashWarp = 18;
constexpr int kFwdThreads = 608;
// Register rebalancing between warpgroups (setmaxnreg is a warpgroup-wide instruction; it must be executed
// INSIDE the role's branch: ptxas budgets the code dominated by it). Used by the GARF kernels (640 threads =
// 5 warpgroups at 96 registers: helper warpgroup 64, row warpgroups 104: their epilogues spill at 96). Tried on
// mlp_fwd / mlp_bwd too and reverted: their row code fits 96 registers, while the MMA issuer loop of the ReLU
// program spills at 64 (forward 1.05 -> 1.17 ms).
constexpr int kHelperRegs = 64;      // 128 x 64 + 512 x 104 = 640 x 96
constexpr int kRowRegs = 104;
__device__ __forceinline__ void regs_helper() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelperRegs)); }
__device__ __forceinline__ void regs_row() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRowRegs)); }
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 256;      // two accumulator buffers: consecutive ops alternate
constexpr uint32_t kWeightCopyBytes = 8192;
constexpr int kMaxSlabs = 8;

// Dynamic shared memory of the MLP kernels: n_slabs slabs | n_stages ring stages | control |
// floats. (5 slabs + 4 stages or 6 slabs + 3 stages, chosen by the program.)
struct MlpSmem {
  static constexpr uint32_t kMaxBiasFloats = 3072;   // packed bias slots of one network
  static constexpr uint32_t kCtrlBytes = 512;
  __host__ __device__ static constexpr uint32_t bytes(int n_slabs, int n_stages) {
    return (uint32_t)n_slabs * NB_SLAB_BYTES + (uint32_t)n_stages * NB_RING_STAGE_BYTES + kCtrlBytes +
           kMaxBiasFloats * 4;
  }

  uint8_t* base;
  uint8_t* ring_base;
  uint64_t* full;          // [NB_MAX_RING_STAGES] weight image landed
  uint64_t* empty;         // [NB_MAX_RING_STAGES] weight image consumed
  uint64_t* slab_ready;    // [kMaxSlabs] row warps -> MMA / stash / helper warps: slab (re)written
  uint64_t* slab_drained;  // [kMaxSlabs] stash warp -> row warps: the stash copy has read the slab
  uint64_t* acc_full;      // [2] MMA warp -> row warps: accumulator buffer complete
  uint64_t* tmem_free;     // [2] row warps -> MMA warp: accumulator buffer read out
  uint64_t* tile_done;     // MMA warp -> row warps: every slab publication of the tile has been consumed
  uint32_t* tmem_ptr;
  float* mask_pos;         // [kMaxLevels]
  float* mask_dir;         // [kMaxLevels]
  float* floats;           // [kMaxBiasFloats] packed biases (forward)
  int n_stages;

  __device__ MlpSmem(uint8_t* b, int n_slabs, int n_stages_) : base(b), n_stages(n_stages_) {
    ring_base = b + (uint32_t)n_slabs * NB_SLAB_BYTES;
    uint8_t* c = ring_base + (uint32_t)n_stages_ * NB_RING_STAGE_BYTES;
    full = reinterpret_cast<uint64_t*>(c);
    empty = full + NB_MAX_RING_STAGES;
    slab_ready = empty + NB_MAX_RING_STAGES;
    slab_drained = slab_ready + kMaxSlabs;
    acc_full = slab_drained + kMaxSlabs;
    tmem_free = acc_full + 2;
    tile_done = tmem_free + 2;
    tmem_ptr = reinterpret_cast<uint32_t*>(tile_done + 1);
    mask_pos = reinterpret_cast<float*>(c + 256);
    mask_dir = mask_pos + kMaxLevels;
    floats = reinterpret_cast<float*>(c + kCtrlBytes);
  }
  __device__ uint8_t* slab(int i) const { return base + (uint32_t)i * NB_SLAB_BYTES; }
  __device__ uint8_t* ring(int s) const { return ring_base + (uint32_t)s * NB_RING_STAGE_BYTES; }

  // thread 0, before the CTA-wide barrier
  __device__ void init_barriers() const {
    for (int s = 0; s < NB_MAX_RING_STAGES; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kMaxSlabs; ++s) {
      tc::mbar_init(&slab_ready[s], kRowWarps);
      tc::mbar_init(&slab_drained[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&tmem_free[b], kRowWarps);
    }
    tc::mbar_init(tile_done, 1);
    tc::fence_barrier_init();
  }
};
static_assert(MlpSmem::bytes(5, 4) <= 227 * 1024 && MlpSmem::bytes(6, 3) <= 227 * 1024, "shared memory budget");
static_assert(NB_RING_STAGE_BYTES % 1024 == 0, "ring stages must keep 1024 B alignment");
static_assert((2 * NB_MAX_RING_STAGES + 2 * kMaxSlabs + 5) * 8 + 4 <= 256, "control block layout");

// ---- who writes which slab when ----------------------------------------------------------------
// Every role of the CTA replays the same production schedule of a tile: "phase -1" is the tile
// start (encodings / head gradient), phase oi is what the row threads write around op oi's
// epilogue. A production of slab s = one completion of slab_ready[s] (all sixteen row warps arrive,
// whether or not they wrote part of it).
struct TileSchedule {
  uint32_t start_mask;      // slabs written at tile start
  int reencode_op;          // forward: phase that starts by re-encoding a shared slab (-1: none)
  uint32_t reencode_mask;
  int is_bwd;
};

__device__ __forceinline__ bool fwd_stores_act(int epi) {
  return epi == NB_EPI_RELU || epi == NB_EPI_LINEAR || epi == NB_EPI_LINEAR_SIGMA || epi == NB_EPI_RELU_SIGMA;
}
__device__ __forceinline__ bool bwd_stores(int epi) {
  return epi == NB_BEPI_MASK || epi == NB_BEPI_PLAIN || epi == NB_BEPI_PLAIN_SIGMA || epi == NB_BEPI_MASK_SIGMA;
}
__device__ __forceinline__ bool bwd_with_sigma(int epi) {
  return epi == NB_BEPI_PLAIN_SIGMA || epi == NB_BEPI_MASK_SIGMA;
}
constexpr int kBwdAuxSlab = 4;   // backward: d(sigma_pre) column

// slabs the epilogue of `op` writes
__device__ __forceinline__ uint32_t op_out_mask(const NbOp& op, int is_bwd) {
  if (!is_bwd) return fwd_stores_act(op.epi) ? ((1u << op.out_chunks) - 1u) : 0u;
  if (!bwd_stores(op.epi)) return 0u;
  return ((1u << op.out_chunks) - 1u) | (bwd_with_sigma(op.epi) ? (1u << kBwdAuxSlab) : 0u);
}
__device__ __forceinline__ uint32_t phase_mask(const TileSchedule& sc, const NbProgram& prog, int oi) {
  return (oi == sc.reencode_op ? sc.reencode_mask : 0u) | op_out_mask(prog.ops[oi], sc.is_bwd);
}

// Consumer-side bookkeeping of slab productions: 4-bit pending count and phase parity per slab.
// A waiter must never fall two phases behind an mbarrier, so every role consumes (acquires)
// every production of the slabs it tracks before the next one can be made.
struct SlabTracker {
  uint32_t pending = 0;
  uint32_t parity = 0;
  __device__ __forceinline__ void produced(uint32_t mask) {
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1u;
      pending += 1u << (4 * s);
    }
  }
  __device__ __forceinline__ void acquire(uint64_t* bars, int s) {
    while ((pending >> (4 * s)) & 15u) {
      tc::mbar_wait(&bars[s], (parity >> s) & 1u);
      parity ^= 1u << s;
      pending -= 1u << (4 * s);
    }
  }
  // Warp-collective form: lane 0 polls, the warp follows through __syncwarp (32 lanes polling the
  // same mbarrier serialise in the shared-memory pipe: ~500 cycles per wait with 8 warps).
  __device__ __forceinline__ void acquire_warp(uint64_t* bars, int s, int lane) {
    if ((pending >> (4 * s)) & 15u) {
      if (lane == 0) {
        uint32_t pd = pending, pr = parity;
        while ((pd >> (4 * s)) & 15u) {
          tc::mbar_wait(&bars[s], (pr >> s) & 1u);
          pr ^= 1u << s;
          pd -= 1u << (4 * s);
        }
      }
      while ((pending >> (4 * s)) & 15u) {   // every lane replays the bookkeeping
        parity ^= 1u << s;
        pending -= 1u << (4 * s);
      }
      __syncwarp();
    }
  }
  __device__ __forceinline__ void acquire_mask_warp(uint64_t* bars, uint32_t mask, int lane) {
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1u;
      acquire_warp(bars, s, lane);
    }
  }
  __device__ __forceinline__ void acquire_all_warp(uint64_t* bars, int lane) {
    if (pending == 0u) return;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s) acquire_warp(bars, s, lane);
  }
  __device__ __forceinline__ void acquire_mask(uint64_t* bars, uint32_t mask) {
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1u;
      acquire(bars, s);
    }
  }
  __device__ __forceinline__ void acquire_all(uint64_t* bars) {
    if (pending == 0u) return;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s) acquire(bars, s);
  }
};

// Row-thread view of the stash copies: at most one copy per slab is ever outstanding, so one
// pending bit and one parity bit per slab suffice (no loops in the epilogue's hot path).
struct DrainBits {
  uint32_t pending = 0, parity = 0;
  int last = -1;   // most recently published slab with a stash copy
  __device__ __forceinline__ void produced(uint32_t mask) {
    pending |= mask;
    if (mask) last = 31 - __clz(mask);
  }
  // warp-collective: lane 0 polls slab_drained[s] if a copy of slab s is outstanding
  __device__ __forceinline__ void acquire(uint64_t* bars, int s, int lane) {
    if ((pending >> s) & 1u) {
      tc::mbar_wait(&bars[s], (parity >> s) & 1u);
      parity ^= 1u << s;
      pending &= ~(1u << s);
    }
  }
  __device__ __forceinline__ void acquire_mask(uint64_t* bars, uint32_t mask, int lane) {
    mask &= pending;
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1u;
      acquire(bars, s, lane);
    }
  }
  // The stash copies drain in the order the slabs were published, so when the most recently
  // published slab is among those about to be rewritten, waiting for it covers the others: one
  // barrier round trip per op instead of four.
  __device__ __forceinline__ void acquire_ordered(uint64_t* bars, uint32_t mask, int lane) {
    mask &= pending;
    if (mask == 0u) return;
    if (last >= 0 && ((mask >> last) & 1u)) {
      tc::mbar_wait(&bars[last], (parity >> last) & 1u);
      parity ^= mask;
      pending &= ~mask;
    } else {
      acquire_mask(bars, mask, lane);
    }
  }
};

// publish one slab: proxy fence, then one arrival per warp
__device__ __forceinline__ void signal_slab(uint64_t* slab_ready, int s, int lane) {
  tc::writer_proxy_fence();
  __syncwarp();
  if (lane == 0) tc::mbar_arrive(&slab_ready[s]);
}

// lane 0 polls, the warp follows
__device__ __forceinline__ void warp_mbar_wait(uint64_t* bar, uint32_t parity, int lane) {
  (void)lane;
  tc::mbar_wait(bar, parity);   // one warp-wide TRYWAIT (+ hardware sleep): cheaper than a divergent leader poll
}

// Row warps: publish the slabs in `mask` (generic-proxy writes -> async proxy, one arrival per warp).
__device__ __forceinline__ void signal_slabs(const uint64_t* slab_ready, uint32_t mask, int lane) {
  tc::writer_proxy_fence();
  __syncwarp();
  if (lane == 0) {
    while (mask) {
      const int s = __ffs(mask) - 1;
      mask &= mask - 1u;
      tc::mbar_arrive(const_cast<uint64_t*>(&slab_ready[s]));
    }
  }
}

struct MlpFwdParams {
  NbProgram prog;
  const uint8_t* wpack;
  const float* bias;
  NbMlpInputs in;
  int N;
  NbPeCfg pe_pos, pe_dir;
  const float* alpha_pos;
  const float* alpha_dir;
  float sigma_bias;
  float* out_sigma;
  float* out_rgb;
  uint8_t* stash;
  uint32_t* masks;
  int n_bias_floats;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// torch.nn.Softplus(beta=1, threshold=8) (barf/model_interpolation_architecture.py:89)
__device__ __forceinline__ float softplus8(float x) { return x > 8.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }

// Query position / direction / frustum parameters of sample n.
__device__ __forceinline__ void load_sample(const NbMlpInputs& in, long long n, PeSample& s) {
  const long long ray = n / in.S;
  s.t0 = in.t_start ? __ldg(in.t_start + n) : 0.f;
  s.t1 = in.t_end ? __ldg(in.t_end + n) : s.t0;
  s.pixel_width = in.pixel_width ? __ldg(in.pixel_width + (in.pixel_width_per_sample ? n : ray)) : 0.f;
  if (in.pos != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.x[c] = __ldg(in.pos + n * 3 + c);
      s.dir[c] = __ldg(in.dir + n * 3 + c);
    }
  } else {
    const float tq = in.t_mode == 0 ? s.t0 : (s.t0 + s.t1) * 0.5f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.dir[c] = __ldg(in.ray_d + ray * 3 + c);
      s.x[c] = __fadd_rn(__ldg(in.ray_o + ray * 3 + c), __fmul_rn(tq, s.dir[c]));
    }
  }
}

// Writes the encoding of one sample as bf16 into row `row` of a slab (zero padded).
__device__ __forceinline__ void encode_to_slab_at(const NbPeCfg& cfg, const float* mask,
                                                  const PeSample& s, uint8_t* slab, int row) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
        make_uint4(0u, 0u, 0u, 0u);
  pe_encode(cfg, mask, s, [&](int col, float v) {
    *reinterpret_cast<__nv_bfloat16*>(slab + tc::slab_offset((uint32_t)row, (uint32_t)col)) =
        __float2bfloat16_rn(v);
  });
}

// The same, split over the four threads of a row (cq = column quarter = threadIdx.x >> 7): every
// thread clears its 16-column quarter, a named barrier over the row threads orders the clears
// before the element stores, then quarter c < 3 encodes coordinate c and quarter 3 the identity
// columns. One thread per row used to do all of it: ~15 k cycles per tile on a single warp per
// scheduler (nine sincosf, sixty 2-byte stores and nothing to hide their latency behind).
// Must be called by ALL row threads (the barrier counts kRowThreads).
__device__ __forceinline__ void encode_to_slab_split(const NbPeCfg& cfg, const float* mask, const PeSample& s,
                                                     uint8_t* slab, int row, int cq) {
#pragma unroll
  for (int q = 0; q < 2; ++q)
    *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)((2 * cq + q) ^ (row & 7)) << 4)) =
        make_uint4(0u, 0u, 0u, 0u);
  named_bar_sync(1, kRowThreads);
  pe_encode(cfg, mask, s, [&](int col, float v) {
    *reinterpret_cast<__nv_bfloat16*>(slab + tc::slab_offset((uint32_t)row, (uint32_t)col)) =
        __float2bfloat16_rn(v);
  }, cq < 3 ? (1 << cq) : 8);
}

// Writes the encoding of one sample as bf16 into row `row` of the encoder's slab (zero padded).
__device__ __forceinline__ void encode_to_slab(const NbPeCfg& cfg, const float* mask,
                                               const PeSample& s, const MlpSmem& sm, int row) {
  if (cfg.slab < 0) return;
  uint8_t* slab = sm.slab(cfg.slab);
  // zero the row first (pad columns must be exact zeros: they meet zero weights)
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<uint4*>(slab + (uint32_t)row * 128u + ((uint32_t)(q ^ (row & 7)) << 4)) =
        make_uint4(0u, 0u, 0u, 0u);
  pe_encode(cfg, mask, s, [&](int col, float v) {
    *reinterpret_cast<__nv_bfloat16*>(slab + tc::slab_offset((uint32_t)row, (uint32_t)col)) =
        __float2bfloat16_rn(v);
  });
}

// Optional cycle trace (debug builds of the profiling scripts): when non-null, block 0 records
// clock64() stamps of its second tile, 4 per op: [buffer free seen, MMAs issued, acc_full seen,
// epilogue done].
static __device__ long long* g_trace = nullptr;
#ifdef NB_ENABLE_TRACE
// every role loads the pointer once (`NB_TRACE_INIT()`), block 0 traces its second tile
#define NB_TRACE_INIT()                                                                         \
  long long* const nb_trace_ptr = (blockIdx.x == 0) ? *(long long* volatile*)&g_trace : nullptr; \
  const int nb_trace_tile = (int)(blockIdx.x + gridDim.x)
#define NB_TRACE(slot, cond)                                                          \
  do {                                                                                \
    if (nb_trace_ptr != nullptr && tile == nb_trace_tile && (cond)) nb_trace_ptr[(slot)] = clock64(); \
  } while (0)
#else
// production builds carry no trace code: the stamps cost ~15 % of the epilogue's instructions.
// scripts/trace_mlp.py runs against a build made with EXTRA=-DNB_ENABLE_TRACE.
#define NB_TRACE_INIT() do { } while (0)
#define NB_TRACE(slot, cond) do { } while (0)
#endif

// ---- warp-specialised loops shared by the forward and backward-data kernels ------------------
// Weight producer (one thread): streams every weight image of the program, tile after tile,
// through the ring with cp.async.bulk; completion lands on full[stage].
__device__ __forceinline__ void weight_producer_loop(const NbProgram& prog, const uint8_t* wpack,
                                                     const MlpSmem& sm, int n_tiles) {
  uint32_t stage = 0, phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int oi = 0; oi < prog.n_ops; ++oi) {
      const NbOp& op = prog.ops[oi];
      const int n_chunks = op.n_chunks;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t bytes = (uint32_t)op.w_rows[c] * 128u * (uint32_t)op.n_sub[c];
        const uint8_t* src = wpack + (size_t)op.w_off[c] * 1024u;
        tc::mbar_wait(&sm.empty[stage], phase ^ 1u);
        tc::mbar_arrive_expect_tx(&sm.full[stage], bytes);
        // several smaller copies per image keep more requests in flight in the TMA engine
        for (uint32_t off = 0; off < bytes; off += kWeightCopyBytes) {
          const uint32_t nb = (bytes - off) < kWeightCopyBytes ? (bytes - off) : kWeightCopyBytes;
          tc::bulk_g2s(sm.ring(stage) + off, src + off, nb, &sm.full[stage]);
        }
        if (++stage == (uint32_t)sm.n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  }
}

// MMA issuer. Run by the WHOLE warp, converged, with warp-uniform values only: ptxas then keeps
// descriptors, counters and addresses in uniform registers and emits one predicated UTCHMMA per
// MMA. (Issuing from inside an `if (lane == 0)` region made every tcgen05.mma a ~300-cycle
// R2UR + elect waterfall loop.) Only the elected lane executes tcgen05.mma / tcgen05.commit.
// The descriptor high words are constant; per MMA only the 14-bit start-address field changes.
//
// Pipelining: op g accumulates into TMEM buffer (g & 1), so its MMAs run while the row threads
// still read op g-1's accumulator; each K chunk is issued as soon as the slab it reads has been
// published (slab_ready), i.e. the MMAs of a layer trail the epilogue of the previous layer by
// one slab instead of waiting for all of it.
__device__ __forceinline__ void mma_issuer_loop(const NbProgram& prog, const TileSchedule& sched,
                                                const MlpSmem& sm, uint32_t tmem_base_in, int n_tiles) {
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);   // provably uniform
  const uint32_t desc_hi = (uint32_t)(tc::umma_desc(0u, 0u, 1024u) >> 32);
  const uint32_t slab0 = tc::smem_u32(sm.slab(0)) >> 4;
  const uint32_t ring0 = tc::smem_u32(sm.ring(0)) >> 4;
  const bool elected = tc::elect_one();
  uint32_t stage = 0, phase = 0, g_op = 0;
  SlabTracker trk;
  NB_TRACE_INIT();
  uint32_t carry = 0;   // productions of the phase before the next op
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int oi = 0; oi < prog.n_ops; ++oi) {
      const NbOp& op = prog.ops[oi];
      const int n_chunks = op.n_chunks, n_blocks = op.n_blocks;
      trk.produced(carry | (oi == 0 ? sched.start_mask : 0u));
      carry = phase_mask(sched, prog, oi);
      const uint32_t buf = g_op & 1u;
      uint32_t idesc[NB_MAX_BLOCKS], tcol[NB_MAX_BLOCKS], brow[NB_MAX_BLOCKS], acc_in[NB_MAX_BLOCKS];
#pragma unroll
      for (int b = 0; b < NB_MAX_BLOCKS; ++b) {
        const NbBlock blk = op.blocks[b < n_blocks ? b : 0];
        idesc[b] = tc::umma_idesc(NB_TILE_ROWS, blk.n, false, false);
        const uint32_t col = (uint32_t)blk.tmem_col;
        tcol[b] = tmem_base + (col >= kAccCols ? (buf ^ 1u) * kAccCols + (col - kAccCols) : buf * kAccCols + col);
        brow[b] = (uint32_t)blk.row0 * 8u;               // row0 * 128 B >> 4
        acc_in[b] = 0u;
      }
      // the accumulator buffer must have been read out by the epilogue of op g-2
      if (g_op >= 2u) tc::mbar_wait(&sm.tmem_free[buf], ((g_op >> 1) - 1u) & 1u);
      tc::tcgen05_fence_after();
      NB_TRACE(oi * 4 + 0, elected);
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t a_lo = slab0 + (uint32_t)op.a_src[c] * (NB_SLAB_BYTES >> 4);
        uint32_t b_lo = ring0 + stage * (NB_RING_STAGE_BYTES >> 4);
        const int k16 = op.k16[c];
        const int bmask = op.blk_mask[c];
        const int n_sub = op.n_sub[c];
        const uint32_t sub_stride = (uint32_t)op.w_rows[c] * 8u;   // image rows * 128 B >> 4
        for (int sub = 0; sub < n_sub; ++sub) trk.acquire(sm.slab_ready, op.a_src[c] + sub);
        tc::consumer_proxy_fence();
        NB_TRACE(128 + oi * 8 + c, elected && oi < 16 && c < 8);
        tc::mbar_wait(&sm.full[stage], phase);
        NB_TRACE(256 + oi * 8 + c, elected && oi < 16 && c < 8);
        tc::tcgen05_fence_after();
        if (n_blocks == 1) {
          // the common shape (one N block): nothing but the two descriptor low words changes per MMA
          const uint32_t tc0 = tcol[0], id0 = idesc[0], br0 = brow[0];
          for (int sub = 0; sub < n_sub; ++sub) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < k16) {
                const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + br0 + 2u * k);
                if (elected) tc::umma(tc0, adesc, bdesc, id0, acc_in[0] | (uint32_t)(k > 0));
              }
            }
            acc_in[0] = 1u;
            a_lo += (NB_SLAB_BYTES >> 4);
            b_lo += sub_stride;
          }
        } else
        for (int sub = 0; sub < n_sub; ++sub) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < k16) {
              const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * k);
#pragma unroll
              for (int b = 0; b < NB_MAX_BLOCKS; ++b) {
                if (b < n_blocks && ((bmask >> b) & 1)) {
                  const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + brow[b] + 2u * k);
                  if (elected) tc::umma(tcol[b], adesc, bdesc, idesc[b], acc_in[b] | (uint32_t)(k > 0));
                  if (k == k16 - 1) acc_in[b] = 1u;   // later MMAs accumulate onto this block
                }
              }
            }
          }
          a_lo += (NB_SLAB_BYTES >> 4);
          b_lo += sub_stride;
        }
        if (elected) tc::umma_commit(&sm.empty[stage]);
        NB_TRACE(448 + c, elected && oi == 4);
        if (++stage == (uint32_t)sm.n_stages) { stage = 0; phase ^= 1u; }
      }
      // consume the productions this op did not read, so that no barrier runs two phases ahead
      trk.acquire_all(sm.slab_ready);
      if (elected) tc::umma_commit(&sm.acc_full[buf]);
      NB_TRACE(oi * 4 + 1, elected);
      __syncwarp();
      ++g_op;
    }
    // A waiter must never fall two phases behind an mbarrier: the slabs the last op published
    // (possibly read by no MMA) are consumed here, and only then may the row warps publish the
    // next tile's first slabs (tile_done).
    trk.produced(carry);
    carry = 0u;
    trk.acquire_all(sm.slab_ready);
    if (elected) tc::mbar_arrive(sm.tile_done);
    __syncwarp();
  }
}

// Stash copier (one thread): every published slab that has a place in the HBM stash goes out as
// one cp.async.bulk; when the copies of a phase have finished reading shared memory the slabs
// are handed back to the row warps (slab_drained). dst_of(phase, slab) -> stash slab or -1.
template <typename DstFn>
__device__ __forceinline__ void stash_loop(const NbProgram& prog, const TileSchedule& sched,
                                           const MlpSmem& sm, int n_tiles, uint8_t* stash,
                                           int slabs_per_tile, DstFn dst_of) {
  SlabTracker trk;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    uint8_t* tile_stash = stash + (size_t)tile * slabs_per_tile * NB_SLAB_BYTES;
    for (int ph = -1; ph < prog.n_ops; ++ph) {
      const uint32_t mask = ph < 0 ? sched.start_mask : phase_mask(sched, prog, ph);
      if (mask == 0u) continue;
      const uint32_t first = (ph >= 0 && ph == sched.reencode_op) ? sched.reencode_mask : 0u;
      int prev = -1;   // slab of the newest copy in flight
      trk.produced(mask);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        uint32_t m = pass == 0 ? first : (mask & ~first);
        while (m) {
          const int s = __ffs(m) - 1;
          m &= m - 1u;
          trk.acquire(sm.slab_ready, s);
          tc::consumer_proxy_fence();
          const int d = dst_of(ph, s);
          if (d >= 0) {
            tc::bulk_s2g(tile_stash + (size_t)d * NB_SLAB_BYTES, sm.slab(s), NB_SLAB_BYTES);
            tc::bulk_commit();
            if (prev >= 0) {               // the copy before this one has read its slab
              tc::bulk_wait_read<1>();
              tc::mbar_arrive(&sm.slab_drained[prev]);
            }
            prev = s;
          }
        }
      }
      if (prev >= 0) {
        tc::bulk_wait_read<0>();
        tc::mbar_arrive(&sm.slab_drained[prev]);
      }
    }
  }
  tc::bulk_wait_all<0>();
}

// Column sums over the 32 lanes of a warp: lane i holds v[0..31] (one matrix row); on return
// lane i holds sum over lanes of v[i]. 31 shuffles (recursive halving), v is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = upper ? v[i] : v[i + s];
      const float keep = upper ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// Range checks of a program before it reaches a kernel.
inline int validate_program(const NbProgram& prog) {
  NB_CHECK_ARG((prog.n_slabs == 5 && prog.n_stages == 4) || (prog.n_slabs == 6 && prog.n_stages == 3),
               "program: unsupported shared-memory shape (%d slabs, %d stages)", prog.n_slabs, prog.n_stages);
  for (int i = 0; i < prog.n_ops; ++i) {
    const NbOp& op = prog.ops[i];
    NB_CHECK_ARG(op.n_chunks >= 1 && op.n_chunks <= NB_MAX_CHUNKS, "program op %d: n_chunks=%d", i, op.n_chunks);
    NB_CHECK_ARG(op.n_blocks >= 1 && op.n_blocks <= NB_MAX_BLOCKS, "program op %d: n_blocks=%d", i, op.n_blocks);
    NB_CHECK_ARG(op.out_chunks >= 0 && op.out_chunks <= 4, "program op %d: out_chunks=%d", i, op.out_chunks);
    for (int c = 0; c < op.n_chunks; ++c) {
      NB_CHECK_ARG(op.a_src[c] >= 0 && op.a_src[c] < prog.n_slabs, "program op %d: a_src=%d", i, op.a_src[c]);
      NB_CHECK_ARG(op.k16[c] >= 1 && op.k16[c] <= 4, "program op %d: k16=%d", i, op.k16[c]);
      NB_CHECK_ARG(op.w_rows[c] >= 8 && op.w_rows[c] * 128 <= NB_RING_STAGE_BYTES && op.w_rows[c] % 8 == 0,
                   "program op %d: w_rows=%d", i, op.w_rows[c]);
      NB_CHECK_ARG(op.w_off[c] >= 0, "program op %d: w_off=%d", i, op.w_off[c]);
      NB_CHECK_ARG(op.blk_mask[c] > 0 && op.blk_mask[c] < (1 << op.n_blocks), "program op %d: blk_mask=%d", i, op.blk_mask[c]);
      NB_CHECK_ARG(op.n_sub[c] >= 1 && op.a_src[c] + op.n_sub[c] <= prog.n_slabs &&
                   op.w_rows[c] * 128 * op.n_sub[c] <= NB_RING_STAGE_BYTES, "program op %d: n_sub=%d", i, op.n_sub[c]);
    }
    for (int b = 0; b < op.n_blocks; ++b) {
      const NbBlock& k = op.blocks[b];
      NB_CHECK_ARG(k.n >= 16 && k.n <= 256 && k.n % 16 == 0, "program op %d: block n=%d", i, k.n);
      NB_CHECK_ARG(k.tmem_col >= 0 && (k.tmem_col % (int)kAccCols) + k.n <= (int)kAccCols && k.tmem_col < 2 * (int)kAccCols,
                   "program op %d: tmem_col=%d", i, k.tmem_col);
      // a block in the other accumulator buffer lands on columns the previous epilogue reads
      // first: safe only if every MMA of the op is issued after slab 0 has been published
      if (k.tmem_col >= (int)kAccCols)
        NB_CHECK_ARG((i == 0 || op.a_src[0] == 0) && k.tmem_col - (int)kAccCols + k.n <= 32,
                     "program op %d: side block needs slab 0 as first K chunk", i);
      NB_CHECK_ARG(k.row0 >= 0 && k.row0 % 8 == 0, "program op %d: row0=%d", i, k.row0);
      for (int c = 0; c < op.n_chunks; ++c)
        if ((op.blk_mask[c] >> b) & 1)
          NB_CHECK_ARG(k.row0 + k.n <= op.w_rows[c], "program op %d: block rows exceed image", i);
    }
  }
  return NERFB200_OK;
}

}  // namespace nerfb200
