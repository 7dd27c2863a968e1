// Shared pieces of the fused GARF field kernels (garf_fwd.cu, garf_bwd.cu): shared-memory
// carve-up, the weight producer, the MMA issuer and the stash copier of an NgProgram
// (include/nerfb200_garf.h). The hand-off protocol is op-granular (an op starts when the step in front
// of it is done) except behind the Gaussian steps, whose slabs are handed over one by one (slab_ready,
// NgOp.early) so that the MMAs of op k run under the epilogue of step k.
//
// Barriers (all in shared memory; two of each op-granular kind, used alternately by the global op index g):
//   in_ready[g & 1]  16 row warps -> MMA warp, stash warp: "step g is done": the inputs of op g are
//                    in shared memory and no row thread reads the accumulator columns it overwrites
//   acc_full[g & 1]  MMA warp (tcgen05.commit) -> row warps: op g and every op before it completed
//   drained[g & 1]   stash warp -> row warps: the HBM copies of the slabs published by step g have
//                    read shared memory
//   slab_ready[s]    16 row warps -> MMA warp, stash warp, only in steps whose op is `early`
//                    (NgOp.early: a Gaussian step that publishes slabs, followed by an op that writes
//                    none of the accumulator columns the step reads): slab s of the step is written.
//                    Chunk c of the op is issued as soon as ITS slab is there, so the MMAs of op k run
//                    under the epilogue of step k; in_ready[g & 1] of such a step is still arrived at
//                    (and consumed by both roles) when the whole step is done. The row warps do not
//                    fence these publications: the consumer executes the generic -> async proxy fence
//                    after it has acquired the barrier phase (tc.cuh, consumer_proxy_fence).
// A waiter never falls two phases behind a barrier: step k consumes the completions of every op up
// to k - 1 - wait_lag (wait_lag <= 1) before it does anything, and op k + 1 cannot be issued before
// step k + 1 has arrived; the MMA warp and the stash warp consume every slab_ready completion of an
// early step before they leave that step.
#pragma once
#include "common.cuh"
#include "garf.h"
#include "mlp.h"
#include "mlp_kernels.cuh"   // load_sample, PeSample, softplus8, sigmoidf, named_bar_sync
#include "tc.cuh"

namespace nerfb200 {
namespace garf {

using namespace tc;

constexpr int kRowWarpsG = 16;
constexpr int kRowThreadsG = 512;
constexpr int kMmaWarpG = 16;
constexpr int kProducerWarpG = 17;
constexpr int kStashWarpG = 18;
// 20 warps: 16 row warps + MMA issuer, weight producer, stash copier and one idle warp that completes the
// fifth warpgroup — setmaxnreg is a warpgroup-wide instruction. The kernels launch with 96 registers per
// thread (640 x 96 = 60 K); the helper warpgroup then gives registers back (setmaxnreg.dec) and the four row
// warpgroups take them (setmaxnreg.inc): the epilogues are the register-hungry part (spills at 96).
constexpr int kThreadsG = 640;
// (regs_helper / regs_row: mlp_kernels.cuh)
constexpr uint32_t kTmemColsG = 512;

struct GarfSmem {
  static constexpr uint32_t kCtrlBytes = 512;
  static constexpr uint32_t kXyzBytes = 2u * NB_TILE_ROWS * 16u;   // float4 position + float4 direction per row
  static constexpr uint32_t kStepTableBytes = (uint32_t)((sizeof(NgStep) * (NG_MAX_OPS + 1) + 15) / 16 * 16);
  static constexpr uint32_t kTableBytes = kStepTableBytes + (uint32_t)((sizeof(NgOp) * NG_MAX_OPS + 15) / 16 * 16);
  static constexpr uint32_t kMaxProgramFloats = (uint32_t)NG_MAX_FLOATS - kTableBytes / 4u;   // what a program may pack
  __host__ __device__ static constexpr uint32_t bytes() {
    return (uint32_t)NG_N_SLABS * NB_SLAB_BYTES + (uint32_t)NG_N_STAGES * NB_RING_STAGE_BYTES + kCtrlBytes +
           kXyzBytes + (uint32_t)NG_MAX_FLOATS * 4u;
  }
  uint8_t* base;
  uint8_t* ring_base;
  uint64_t* full;       // [NG_N_STAGES]
  uint64_t* empty;      // [NG_N_STAGES]
  uint64_t* in_ready;   // [2]
  uint64_t* acc_full;   // [2]
  uint64_t* drained;    // [2]
  uint64_t* slab_ready; // [NG_N_SLABS]
  uint32_t* tmem_ptr;
  float4* pos;          // [128] query position of every tile row
  float4* dir;          // [128] ray direction of every tile row
  float* floats;        // [NG_MAX_FLOATS]: packed fp32 values, and behind them (at the end of the region)
  NgStep* steps;        // [NG_MAX_OPS + 1] the program's step and op tables: kernel parameters live in the
  NgOp* ops;            // [NG_MAX_OPS]     constant bank, where a dynamically indexed field costs a dependent
                        //                  ~100-cycle LDC each (10 % of the stall samples of the first version)

  __device__ explicit GarfSmem(uint8_t* b) : base(b) {
    ring_base = b + (uint32_t)NG_N_SLABS * NB_SLAB_BYTES;
    uint8_t* c = ring_base + (uint32_t)NG_N_STAGES * NB_RING_STAGE_BYTES;
    full = reinterpret_cast<uint64_t*>(c);
    empty = full + NG_N_STAGES;
    in_ready = empty + NG_N_STAGES;
    acc_full = in_ready + 2;
    drained = acc_full + 2;
    slab_ready = drained + 2;
    tmem_ptr = reinterpret_cast<uint32_t*>(slab_ready + NG_N_SLABS);
    pos = reinterpret_cast<float4*>(c + kCtrlBytes);
    dir = pos + NB_TILE_ROWS;
    floats = reinterpret_cast<float*>(c + kCtrlBytes + kXyzBytes);
    uint8_t* tables = c + kCtrlBytes + kXyzBytes + (uint32_t)NG_MAX_FLOATS * 4u - kTableBytes;
    steps = reinterpret_cast<NgStep*>(tables);
    ops = reinterpret_cast<NgOp*>(tables + kStepTableBytes);
  }
  // copies the tables of `prog` (all threads, before the CTA-wide barrier)
  __device__ void load_tables(const NgProgram& prog) const {
    const uint32_t* src_s = reinterpret_cast<const uint32_t*>(prog.steps);
    const uint32_t* src_o = reinterpret_cast<const uint32_t*>(prog.ops);
    uint32_t* dst_s = reinterpret_cast<uint32_t*>(steps);
    uint32_t* dst_o = reinterpret_cast<uint32_t*>(ops);
    for (uint32_t i = threadIdx.x; i < sizeof(NgStep) * (NG_MAX_OPS + 1) / 4; i += blockDim.x) dst_s[i] = src_s[i];
    for (uint32_t i = threadIdx.x; i < sizeof(NgOp) * NG_MAX_OPS / 4; i += blockDim.x) dst_o[i] = src_o[i];
  }
  __device__ uint8_t* slab(int i) const { return base + (uint32_t)i * NB_SLAB_BYTES; }
  __device__ uint8_t* ring(int s) const { return ring_base + (uint32_t)s * NB_RING_STAGE_BYTES; }
  __device__ void init_barriers() const {
    for (int s = 0; s < NG_N_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&in_ready[b], kRowWarpsG);
      mbar_init(&acc_full[b], 1);
      mbar_init(&drained[b], 1);
    }
    for (int i = 0; i < NG_N_SLABS; ++i) mbar_init(&slab_ready[i], kRowWarpsG);
    fence_barrier_init();
  }
};
static_assert(GarfSmem::bytes() <= 227 * 1024, "shared memory budget of the GARF kernels");
static_assert((2 * NG_N_STAGES + 6 + NG_N_SLABS) * 8 + 4 <= GarfSmem::kCtrlBytes, "control block layout");

// phase parity of the g-th use of a pair of alternating barriers
__device__ __forceinline__ uint32_t pair_parity(uint32_t g) { return (g >> 1) & 1u; }

// Row-warp side bookkeeping: how many completions of acc_full / drained have been consumed.
struct RowSync {
  uint32_t acc_seen = 0;     // ops whose completion has been observed
  uint32_t drain_seen = 0;   // rounds whose stash copies have been observed drained
  // wait until op `through` (global index, may be -1) and everything before it has completed
  __device__ __forceinline__ void acc_through(const GarfSmem& sm, int through) {
    while ((int)acc_seen <= through) {
      mbar_wait(&sm.acc_full[acc_seen & 1u], pair_parity(acc_seen));
      ++acc_seen;
    }
  }
  __device__ __forceinline__ void drain_through(const GarfSmem& sm, int through) {
    while ((int)drain_seen <= through) {
      mbar_wait(&sm.drained[drain_seen & 1u], pair_parity(drain_seen));
      ++drain_seen;
    }
  }
};

// publish step g: shared-memory writes -> async proxy, accumulator reads ordered before the next MMAs
__device__ __forceinline__ void publish_step(const GarfSmem& sm, uint32_t g, bool wrote_smem, int lane) {
  if (wrote_smem) writer_proxy_fence();
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&sm.in_ready[g & 1u]);
}

// early steps: one slab of the step is written (no proxy fence here: the consumers fence)
__device__ __forceinline__ void publish_slab(const GarfSmem& sm, int slab, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(&sm.slab_ready[slab]);
}

// ---- weight producer (one thread) -------------------------------------------------------------
__device__ __forceinline__ void producer_loop(const NgProgram& prog, const uint8_t* wpack, const GarfSmem& sm,
                                              int n_tiles) {
  uint32_t stage = 0, phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int k = 0; k < prog.n_ops; ++k) {
      const NgOp& op = sm.ops[k];
      const uint32_t bytes = (uint32_t)op.w_rows * 128u;
      for (int c = 0; c < op.n_chunks; ++c) {
        const uint8_t* src = wpack + (size_t)op.w_off[c] * 1024u;
        mbar_wait(&sm.empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&sm.full[stage], bytes);
        for (uint32_t off = 0; off < bytes; off += kWeightCopyBytes) {
          const uint32_t nb = (bytes - off) < kWeightCopyBytes ? (bytes - off) : kWeightCopyBytes;
          bulk_g2s(sm.ring(stage) + off, src + off, nb, &sm.full[stage]);
        }
        if (++stage == (uint32_t)NG_N_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  }
}

// ---- MMA issuer (whole warp converged, one elected lane issues; see mlp_kernels.cuh) -----------
__device__ __forceinline__ void mma_loop(const NgProgram& prog, const GarfSmem& sm, uint32_t tmem_base_in,
                                         int n_tiles) {
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);
  const uint32_t desc_hi = (uint32_t)(umma_desc(0u, 0u, 1024u) >> 32);
  const uint32_t slab0 = smem_u32(sm.slab(0)) >> 4;
  const uint32_t ring0 = smem_u32(sm.ring(0)) >> 4;
  const bool elected = elect_one();
  uint32_t stage = 0, phase = 0, g = 0;
  uint32_t slab_par = 0;     // phase parity of every slab_ready barrier (early steps)
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int k = 0; k < prog.n_ops; ++k, ++g) {
      const NgOp& op = sm.ops[k];
      const int n_chunks = op.n_chunks, n_blocks = op.n_blocks;
      const bool early = op.early != 0;
      // slabs step k publishes one by one (early ops only)
      uint32_t pending = early ? (((1u << sm.steps[k].n_slabs) - 1u) << sm.steps[k].out_slab) : 0u;
      uint32_t idesc[2], tcol[2], brow[2];
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const NgBlock blk = op.blocks[b < n_blocks ? b : 0];
        idesc[b] = umma_idesc(NB_TILE_ROWS, blk.n, false, false);
        tcol[b] = tmem_base + (uint32_t)blk.tmem_col;
        brow[b] = (uint32_t)blk.row0 * 8u;               // row0 * 128 B >> 4
      }
      uint32_t acc = op.accumulate ? 1u : 0u;
      if (!early) {
        mbar_wait(&sm.in_ready[g & 1u], pair_parity(g));
        consumer_proxy_fence();
        tcgen05_fence_after();
      }
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t a_lo = slab0 + (uint32_t)op.a_slab[c] * (NB_SLAB_BYTES >> 4);
        const uint32_t b_lo = ring0 + stage * (NB_RING_STAGE_BYTES >> 4);
        const int k16 = op.k16[c];
        if ((pending >> op.a_slab[c]) & 1u) {      // this chunk's slab is being written by step k
          const uint32_t s = (uint32_t)op.a_slab[c];
          mbar_wait(&sm.slab_ready[s], (slab_par >> s) & 1u);
          slab_par ^= 1u << s;
          pending &= ~(1u << s);
          fence_proxy_async();
        }
        mbar_wait(&sm.full[stage], phase);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (kk < k16) {
            const uint64_t adesc = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2u * kk);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              if (b < n_blocks) {
                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + brow[b] + 2u * kk);
                if (elected) umma(tcol[b], adesc, bdesc, idesc[b], acc);
              }
            }
            acc = 1u;
          }
        }
        if (elected) umma_commit(&sm.empty[stage]);
        if (++stage == (uint32_t)NG_N_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (early) {   // consume what is left of step k: slabs no chunk read, and the step's own completion
        while (pending) {
          const uint32_t s = (uint32_t)__ffs(pending) - 1u;
          mbar_wait(&sm.slab_ready[s], (slab_par >> s) & 1u);
          slab_par ^= 1u << s;
          pending &= pending - 1u;
        }
        mbar_wait(&sm.in_ready[g & 1u], pair_parity(g));
        tcgen05_fence_after();
      }
      if (elected) umma_commit(&sm.acc_full[g & 1u]);
      __syncwarp();
    }
  }
}

// ---- stash copier (one thread): slabs published by step k -> the per-tile HBM stash ------------
__device__ __forceinline__ void stash_copier_loop(const NgProgram& prog, const GarfSmem& sm, int n_tiles,
                                                  uint8_t* stash, int slabs_per_tile) {
  uint32_t g = 0, slab_par = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    uint8_t* tile_stash = stash + (size_t)tile * (size_t)slabs_per_tile * NB_SLAB_BYTES;
    for (int k = 0; k < prog.n_ops; ++k, ++g) {
      const NgStep& st = sm.steps[k];
      if (sm.ops[k].early) {
        // the slabs of an early step arrive one by one: every copy leaves as soon as its slab is written
        for (int j = 0; j < st.n_slabs; ++j) {
          const uint32_t s = (uint32_t)(st.out_slab + j);
          mbar_wait(&sm.slab_ready[s], (slab_par >> s) & 1u);
          slab_par ^= 1u << s;
          fence_proxy_async();
#ifndef NG_EXP_NO_Y
          if (st.y_stash >= 0)
            bulk_s2g(tile_stash + (size_t)(st.y_stash + j) * NB_SLAB_BYTES, sm.slab(st.out_slab + j), NB_SLAB_BYTES);
#endif
        }
        mbar_wait(&sm.in_ready[g & 1u], pair_parity(g));
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(&sm.drained[g & 1u]);
        continue;
      }
      mbar_wait(&sm.in_ready[g & 1u], pair_parity(g));
      consumer_proxy_fence();
      if (st.y_stash >= 0 && st.out_slab >= 0 && !(st.flags & NG_F_DIRECT)) {
        const int extra = (st.kind == NG_BSTEP_PLAIN && (st.flags & NG_F_SIGMA)) ? 1 : 0;
        const int n = st.n_slabs + extra;
#ifndef NG_EXP_NO_Y
        for (int j = 0; j < n; ++j)
          bulk_s2g(tile_stash + (size_t)(st.y_stash + j) * NB_SLAB_BYTES, sm.slab(st.out_slab + j), NB_SLAB_BYTES);
#endif
        bulk_commit();
        bulk_wait_read<0>();
      }
      mbar_arrive(&sm.drained[g & 1u]);
    }
  }
  bulk_wait_all<0>();
}

// Range checks of a program before it reaches a kernel.
inline int validate_garf_program(const NgProgram& prog, bool backward) {
  NB_CHECK_ARG(prog.n_ops >= 1 && prog.n_ops <= NG_MAX_OPS, "garf program: n_ops=%d", prog.n_ops);
  NB_CHECK_ARG(prog.n_floats >= 0 && prog.n_floats <= (int)GarfSmem::kMaxProgramFloats,
               "garf program: %d packed floats (max %d)", prog.n_floats, (int)GarfSmem::kMaxProgramFloats);
  for (int k = 0; k < prog.n_ops; ++k) {
    const NgOp& op = prog.ops[k];
    NB_CHECK_ARG(op.n_chunks >= 0 && op.n_chunks <= NG_MAX_CHUNKS, "garf op %d: n_chunks=%d", k, op.n_chunks);
    if (op.n_chunks == 0) continue;
    NB_CHECK_ARG(op.n_blocks >= 1 && op.n_blocks <= 2, "garf op %d: n_blocks=%d", k, op.n_blocks);
    NB_CHECK_ARG(op.w_rows >= 16 && op.w_rows <= 256 && op.w_rows % 8 == 0, "garf op %d: w_rows=%d", k, op.w_rows);
    for (int c = 0; c < op.n_chunks; ++c) {
      NB_CHECK_ARG(op.a_slab[c] >= 0 && op.a_slab[c] < NG_N_SLABS, "garf op %d: a_slab=%d", k, op.a_slab[c]);
      NB_CHECK_ARG(op.k16[c] >= 1 && op.k16[c] <= 4, "garf op %d: k16=%d", k, op.k16[c]);
      NB_CHECK_ARG(op.w_off[c] >= 0, "garf op %d: w_off=%d", k, op.w_off[c]);
    }
    for (int b = 0; b < op.n_blocks; ++b) {
      const NgBlock& blk = op.blocks[b];
      NB_CHECK_ARG(blk.n >= 16 && blk.n <= 256 && blk.n % 16 == 0, "garf op %d: block n=%d", k, blk.n);
      NB_CHECK_ARG(blk.tmem_col >= 0 && blk.tmem_col + blk.n <= (int)kTmemColsG, "garf op %d: tmem_col=%d", k, blk.tmem_col);
      NB_CHECK_ARG(blk.row0 >= 0 && blk.row0 % 8 == 0 && blk.row0 + blk.n <= op.w_rows, "garf op %d: block rows", k);
    }
  }
  for (int k = 0; k <= prog.n_ops; ++k) {
    const NgStep& st = prog.steps[k];
    NB_CHECK_ARG(st.wait_lag == 0 || st.wait_lag == 1, "garf step %d: wait_lag=%d", k, st.wait_lag);
    NB_CHECK_ARG(st.n_slabs >= 0 && st.n_slabs <= 4, "garf step %d: n_slabs=%d", k, st.n_slabs);
    NB_CHECK_ARG(st.out_slab < 0 || st.out_slab + st.n_slabs <= NG_N_SLABS, "garf step %d: out_slab=%d", k, st.out_slab);
    NB_CHECK_ARG(st.src_col >= 0 && st.src_col + 64 * st.n_slabs <= (int)kTmemColsG + 48, "garf step %d: src_col=%d", k, st.src_col);
    const bool fwd_kind = st.kind >= NG_STEP_NONE && st.kind <= NG_STEP_SIGMA;
    const bool bwd_kind = st.kind == NG_STEP_NONE || (st.kind >= NG_BSTEP_HEAD && st.kind <= NG_BSTEP_PLAIN);
    NB_CHECK_ARG(backward ? bwd_kind : fwd_kind, "garf step %d: kind %d does not belong to this pass", k, st.kind);
    if (k == prog.n_ops) NB_CHECK_ARG(st.out_slab < 0 || (st.flags & NG_F_DIRECT), "garf: the last step cannot publish slabs");
  }
  return NERFB200_OK;
}

// bf16 pair helpers
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts128g(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
// A row thread's two 16-byte chunks of a slab row — (2cq) ^ (row & 7) and its neighbour — are the two
// halves of ONE 32-byte sector: a single 256-bit access instead of two scattered 128-bit ones (with a
// lane per row every access is its own L1 transaction, and the row warps' global accesses are bound
// by the number of those). On odd rows the halves are swapped. `sec` = the sector's address.
__device__ __forceinline__ void stg256_row(void* sec, bool odd, const uint32_t (&v)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(sec),
               "r"(odd ? v[4] : v[0]), "r"(odd ? v[5] : v[1]), "r"(odd ? v[6] : v[2]), "r"(odd ? v[7] : v[3]),
               "r"(odd ? v[0] : v[4]), "r"(odd ? v[1] : v[5]), "r"(odd ? v[2] : v[6]), "r"(odd ? v[3] : v[7])
               : "memory");
}
__device__ __forceinline__ void stg256(void* sec, const uint32_t (&v)[8]) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(sec),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// Z stash layout (private to garf_fwd, garf_bwd and the column sums of mlp_wgrad; never an MMA operand): the
// slab layout with the two 16-byte halves of every 32-byte sector in NATURAL order on all rows, so the row
// thread that owns the sector stores / loads its 16 columns as they are (the slab layout proper swaps
// the halves on odd rows: eight SELs per access). Byte offset of element (row, col) inside a z slab:
__host__ __device__ __forceinline__ uint32_t zstash_offset(uint32_t row, uint32_t col) {
  return row * 128u + ((((col >> 4) ^ ((row & 7u) >> 1)) & 3u) << 5) + ((col & 15u) << 1);
}
// raw sector (halves in memory order); `unswap_row` restores the logical order
struct Sector32 { uint32_t w[8]; };
__device__ __forceinline__ Sector32 ldg256_stream(const void* sec) {
  Sector32 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
               : "l"(sec));
  return r;
}
__device__ __forceinline__ void unswap_row(const Sector32& s, bool odd, uint32_t (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = odd ? s.w[4 + i] : s.w[i];
    v[4 + i] = odd ? s.w[i] : s.w[4 + i];
  }
}
__device__ __forceinline__ uint4 ldg128_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace garf
}  // namespace nerfb200
