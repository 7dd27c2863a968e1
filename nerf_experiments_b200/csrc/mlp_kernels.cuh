// Shared pieces of the fused MLP kernels (forward, backward-data): shared-memory carve-up,
// kernel parameter blocks, sample loading and small math helpers.
#pragma once
#include "common.cuh"
#include "mlp.h"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {

// Warp roles. Warps 0-7 are the "row threads": thread t owns tile row (t & 127) — the TMEM lane
// quarter of a warp is fixed by (warp & 3) — and column half (t >> 7) of every accumulator, so
// two warps per SM sub-partition share the epilogue of each row.
constexpr int kRowThreads = 256;
constexpr int kHalfThreads = 128;
constexpr int kMmaWarp = 8;
constexpr int kProducerWarp = 9;
constexpr int kMlpThreads = 320;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kWeightCopyBytes = 8192;

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
  uint64_t* full;      // [NB_MAX_RING_STAGES]
  uint64_t* empty;     // [NB_MAX_RING_STAGES]
  uint64_t* a_ready;
  uint64_t* acc_full;
  uint64_t* epi_done;     // backward: row threads -> bias-gradient helper warps
  uint64_t* helper_done;  // backward: helper warps -> row threads
  uint32_t* tmem_ptr;
  float* mask_pos;     // [kMaxLevels]
  float* mask_dir;     // [kMaxLevels]
  float* floats;       // [kMaxBiasFloats] biases (forward) / bias-gradient accumulators (backward)
  int n_stages;

  __device__ MlpSmem(uint8_t* b, int n_slabs, int n_stages_) : base(b), n_stages(n_stages_) {
    ring_base = b + (uint32_t)n_slabs * NB_SLAB_BYTES;
    uint8_t* c = ring_base + (uint32_t)n_stages_ * NB_RING_STAGE_BYTES;
    full = reinterpret_cast<uint64_t*>(c);
    empty = full + NB_MAX_RING_STAGES;
    a_ready = empty + NB_MAX_RING_STAGES;
    acc_full = a_ready + 1;
    epi_done = acc_full + 1;
    helper_done = epi_done + 1;
    tmem_ptr = reinterpret_cast<uint32_t*>(helper_done + 1);
    mask_pos = reinterpret_cast<float*>(c + 128);
    mask_dir = mask_pos + kMaxLevels;
    floats = reinterpret_cast<float*>(c + kCtrlBytes);
  }
  __device__ uint8_t* slab(int i) const { return base + (uint32_t)i * NB_SLAB_BYTES; }
  __device__ uint8_t* ring(int s) const { return ring_base + (uint32_t)s * NB_RING_STAGE_BYTES; }
};
static_assert(MlpSmem::bytes(5, 4) <= 227 * 1024 && MlpSmem::bytes(6, 3) <= 227 * 1024, "shared memory budget");
static_assert(NB_RING_STAGE_BYTES % 1024 == 0, "ring stages must keep 1024 B alignment");

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

// Thread-0 bookkeeping of the shared->global stash copies in flight. Every slab goes out as
// its own bulk group, in slab order, so that the next epilogue can start rewriting slab j as
// soon as the copies that read slabs <= j have drained, while the later slabs still stream out.
struct StashQueue {
  int last_batch = 0;
  __device__ __forceinline__ void begin_batch() { last_batch = 0; }
  __device__ __forceinline__ void push(void* gdst, const void* ssrc, uint32_t bytes) {
    tc::bulk_s2g(gdst, ssrc, bytes);
    tc::bulk_commit();
    ++last_batch;
  }
  // returns once every copy reading act slab j (or an older batch) has finished reading
  __device__ __forceinline__ void wait_slab(int j) const {
    int allow = last_batch - 1 - j;
    switch (allow) {
      case 4: tc::bulk_wait_read<4>(); break;
      case 3: tc::bulk_wait_read<3>(); break;
      case 2: tc::bulk_wait_read<2>(); break;
      case 1: tc::bulk_wait_read<1>(); break;
      default: tc::bulk_wait_read<0>(); break;
    }
  }
  __device__ __forceinline__ void wait_all() const { tc::bulk_wait_read<0>(); }
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
// clock64() stamps of its second tile, 4 per op: [a_ready seen, MMAs issued, acc_full seen,
// epilogue done].
static __device__ long long* g_trace = nullptr;
#define NB_TRACE(slot, cond)                                                   \
  do {                                                                         \
    if (g_trace != nullptr && blockIdx.x == 0 && (cond)) g_trace[(slot)] = clock64(); \
  } while (0)

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
        NB_TRACE(64 + oi * 8 + c, tile == (int)(blockIdx.x + gridDim.x) && oi < 4);
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
__device__ __forceinline__ void mma_issuer_loop(const NbProgram& prog, const MlpSmem& sm,
                                                uint32_t tmem_base_in, int n_tiles) {
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_in, 0);   // provably uniform
  const uint32_t desc_hi = (uint32_t)(tc::umma_desc(0u, 0u, 1024u) >> 32);
  const uint32_t slab0 = tc::smem_u32(sm.slab(0)) >> 4;
  const uint32_t ring0 = tc::smem_u32(sm.ring(0)) >> 4;
  const bool elected = tc::elect_one();
  uint32_t stage = 0, phase = 0, a_phase = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int oi = 0; oi < prog.n_ops; ++oi) {
      const NbOp& op = prog.ops[oi];
      const int n_chunks = op.n_chunks, n_blocks = op.n_blocks;
      uint32_t idesc[NB_MAX_BLOCKS], tcol[NB_MAX_BLOCKS], brow[NB_MAX_BLOCKS], acc_in[NB_MAX_BLOCKS];
#pragma unroll
      for (int b = 0; b < NB_MAX_BLOCKS; ++b) {
        const NbBlock blk = op.blocks[b < n_blocks ? b : 0];
        idesc[b] = tc::umma_idesc(NB_TILE_ROWS, blk.n, false, false);
        tcol[b] = tmem_base + (uint32_t)blk.tmem_col;
        brow[b] = (uint32_t)blk.row0 * 8u;               // row0 * 128 B >> 4
        acc_in[b] = blk.accum_in ? 1u : 0u;
      }
      tc::mbar_wait(sm.a_ready, a_phase);
      a_phase ^= 1u;
      tc::tcgen05_fence_after();
      NB_TRACE(oi * 4 + 0, elected && tile == (int)(blockIdx.x + gridDim.x));
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t a_lo = slab0 + (uint32_t)op.a_src[c] * (NB_SLAB_BYTES >> 4);
        uint32_t b_lo = ring0 + stage * (NB_RING_STAGE_BYTES >> 4);
        const int k16 = op.k16[c];
        const int bmask = op.blk_mask[c];
        const int n_sub = op.n_sub[c];
        const uint32_t sub_stride = (uint32_t)op.w_rows[c] * 8u;   // image rows * 128 B >> 4
        tc::mbar_wait(&sm.full[stage], phase);
        tc::tcgen05_fence_after();
        NB_TRACE(128 + oi * 8 + c, elected && tile == (int)(blockIdx.x + gridDim.x) && oi < 4);
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
        if (++stage == (uint32_t)sm.n_stages) { stage = 0; phase ^= 1u; }
      }
      if (elected) tc::umma_commit(sm.acc_full);
      NB_TRACE(oi * 4 + 1, elected && tile == (int)(blockIdx.x + gridDim.x));
      __syncwarp();
    }
  }
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
      NB_CHECK_ARG(k.tmem_col >= 0 && k.tmem_col + k.n <= (int)kTmemCols, "program op %d: tmem_col=%d", i, k.tmem_col);
      NB_CHECK_ARG(k.row0 >= 0 && k.row0 % 8 == 0, "program op %d: row0=%d", i, k.row0);
      for (int c = 0; c < op.n_chunks; ++c)
        if ((op.blk_mask[c] >> b) & 1)
          NB_CHECK_ARG(k.row0 + k.n <= op.w_rows[c], "program op %d: block rows exceed image", i);
    }
  }
  return NERFB200_OK;
}

}  // namespace nerfb200
