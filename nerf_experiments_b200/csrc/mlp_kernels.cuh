// Shared pieces of the fused MLP kernels (forward, backward-data): shared-memory carve-up,
// kernel parameter blocks, sample loading and small math helpers.
#pragma once
#include "common.cuh"
#include "mlp.h"
#include "pe.cuh"
#include "tc.cuh"

namespace nerfb200 {

constexpr int kRowThreads = 128;   // warps 0-3: one tile row each
constexpr int kMmaWarp = 4;
constexpr int kProducerWarp = 5;
constexpr int kMlpThreads = 192;
constexpr uint32_t kTmemCols = 512;

// Dynamic shared memory of the MLP kernels.
struct MlpSmem {
  static constexpr uint32_t kSlabsOff = 0;
  static constexpr uint32_t kRingOff = NB_N_SLABS * NB_SLAB_BYTES;
  static constexpr uint32_t kCtrlOff = kRingOff + NB_RING_STAGES * NB_RING_STAGE_BYTES;
  static constexpr uint32_t kFloatsOff = kCtrlOff + 512;
  static constexpr uint32_t kMaxBiasFloats = 4096;   // packed bias slots of one network
  static constexpr uint32_t kBytes = kFloatsOff + kMaxBiasFloats * 4;

  uint8_t* base;
  uint64_t* full;      // [NB_RING_STAGES]
  uint64_t* empty;     // [NB_RING_STAGES]
  uint64_t* a_ready;
  uint64_t* acc_full;
  uint32_t* tmem_ptr;
  float* mask_pos;     // [kMaxLevels]
  float* mask_dir;     // [kMaxLevels]
  float* floats;       // [kMaxBiasFloats] bias-gradient accumulators (backward)

  __device__ explicit MlpSmem(uint8_t* b) : base(b) {
    uint8_t* c = b + kCtrlOff;
    full = reinterpret_cast<uint64_t*>(c);
    empty = full + NB_RING_STAGES;
    a_ready = empty + NB_RING_STAGES;
    acc_full = a_ready + 1;
    tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
    mask_pos = reinterpret_cast<float*>(c + 128);
    mask_dir = mask_pos + kMaxLevels;
    floats = reinterpret_cast<float*>(b + kFloatsOff);
  }
  __device__ uint8_t* slab(int i) const { return base + kSlabsOff + (uint32_t)i * NB_SLAB_BYTES; }
  __device__ uint8_t* ring(int s) const { return base + kRingOff + (uint32_t)s * NB_RING_STAGE_BYTES; }
};
static_assert(MlpSmem::kBytes <= 227 * 1024, "shared memory budget");
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
  for (int i = 0; i < prog.n_ops; ++i) {
    const NbOp& op = prog.ops[i];
    NB_CHECK_ARG(op.n_chunks >= 1 && op.n_chunks <= NB_MAX_CHUNKS, "program op %d: n_chunks=%d", i, op.n_chunks);
    NB_CHECK_ARG(op.n_blocks >= 1 && op.n_blocks <= NB_MAX_BLOCKS, "program op %d: n_blocks=%d", i, op.n_blocks);
    NB_CHECK_ARG(op.out_chunks >= 0 && op.out_chunks <= 4, "program op %d: out_chunks=%d", i, op.out_chunks);
    for (int c = 0; c < op.n_chunks; ++c) {
      NB_CHECK_ARG(op.a_src[c] >= 0 && op.a_src[c] < NB_N_SLABS, "program op %d: a_src=%d", i, op.a_src[c]);
      NB_CHECK_ARG(op.k16[c] >= 1 && op.k16[c] <= 4, "program op %d: k16=%d", i, op.k16[c]);
      NB_CHECK_ARG(op.w_rows[c] >= 8 && op.w_rows[c] * 128 <= NB_RING_STAGE_BYTES && op.w_rows[c] % 8 == 0,
                   "program op %d: w_rows=%d", i, op.w_rows[c]);
      NB_CHECK_ARG(op.w_off[c] >= 0, "program op %d: w_off=%d", i, op.w_off[c]);
    }
    for (int b = 0; b < op.n_blocks; ++b) {
      const NbBlock& k = op.blocks[b];
      NB_CHECK_ARG(k.n >= 16 && k.n <= 256 && k.n % 16 == 0, "program op %d: block n=%d", i, k.n);
      NB_CHECK_ARG(k.tmem_col >= 0 && k.tmem_col + k.n <= (int)kTmemCols, "program op %d: tmem_col=%d", i, k.tmem_col);
      NB_CHECK_ARG(k.row0 >= 0 && k.row0 % 8 == 0, "program op %d: row0=%d", i, k.row0);
      for (int c = 0; c < op.n_chunks; ++c)
        NB_CHECK_ARG(k.row0 + k.n <= op.w_rows[c], "program op %d: block rows exceed image", i);
    }
  }
  return NERFB200_OK;
}

}  // namespace nerfb200
