/*
 * nerfb200_garf.h — tile programs of the fused GARF field kernels (part of the C ABI of
 * libnerfb200.so, included by nerfb200.h).
 *
 * The GARF networks (reference garf/model_radiance.py:23-96, garf/model_proposal.py:22-56,
 * barf/gaussian.py:10-31) take RAW xyz (no positional encoding), are 1024 / 512 wide behind their
 * first layer, use the Gaussian activation y = exp(-z^2 (1/sigma^2 + 1e-6)) with a learnable width
 * per feature, add a residual (z1 + z2[:, :128]) and concatenate raw xyz / directions into two inner
 * layers. One 128-sample tile stays on chip for the whole network:
 *
 *   - the FIRST layer (3 -> 1024 / 512) is evaluated in fp32 in registers, 128 columns at a time
 *     (raw coordinates would lose the network's high frequencies in bf16), written as bf16 slabs;
 *   - every other Linear is a chain of tcgen05 bf16 MMAs (A = activation slabs in shared memory,
 *     B = packed weight images streamed by the TMA engine, D = fp32 in TMEM); layers wider than the
 *     256 columns a tile keeps in shared memory are column-blocked: a block of the wide layer is
 *     produced and immediately consumed as K chunks of the next layer, which accumulates in TMEM;
 *   - the raw xyz / direction columns of the two concatenating layers are rank-3 fp32 updates in
 *     the epilogue (again: no bf16 rounding of raw coordinates);
 *   - the epilogue applies bias + Gaussian (exp2 on the MUFU pipe) and, in training, writes the
 *     activation y (operand of the weight-gradient kernel) and the pre-activation z (needed by
 *     dy/dz = -2 z v y) to HBM stashes.
 *
 * A program is two interleaved lists: ops[k] (what the MMA warp issues) and steps[k] (what the
 * sixteen row warps do BEFORE op k may be issued; steps[n_ops] runs after the last op).
 * Step k first waits until op (k - 1 - wait_lag) has completed: wait_lag = 0 for a step that reads
 * op k-1's accumulator or rewrites slabs op k-1 read, 1 for a step that works one op ahead.
 */
#ifndef NERFB200_GARF_H_
#define NERFB200_GARF_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NG_MAX_OPS = 32,
  NG_MAX_CHUNKS = 6,        /* K chunks (64 columns = one slab each) of one op                 */
  NG_N_SLABS = 6,           /* shared-memory slabs: 0..3 work, 4..5 hold / small blocks        */
  NG_N_STAGES = 3,          /* weight ring stages of 32 KB                                     */
  NG_GEN_COLS = 128,        /* columns of the first layer one NG_STEP_GEN produces (2 slabs)   */
  NG_MAX_FLOATS = 7680      /* fp32 slots of the shared-memory region holding the packed values; a */
                            /* program may use 6900 (the kernels keep its step / op tables behind) */
};

/* step kinds (forward) */
enum {
  NG_STEP_NONE = 0,
  NG_STEP_GEN = 1,          /* y = G(W1 x + b1), columns gen_col0 .. +127, fp32 in registers   */
  NG_STEP_ACT = 2,          /* y = G(acc + bias [+ skip3]) -> slabs (+ y / z stashes)           */
  NG_STEP_LINEAR = 3,       /* y = acc + bias [+ residual slab] -> slabs; NG_F_SIGMA: density   */
  NG_STEP_RGB = 4,          /* rgb = sigmoid(acc[0:3] + bias) -> out                            */
  NG_STEP_SIGMA = 5,        /* sigma = softplus8(acc[0] + bias[0] + sigma_bias) -> out           */
  /* backward (data gradients) */
  NG_BSTEP_HEAD = 8,        /* gradients w.r.t. the output pre-activations -> slab 0             */
  NG_BSTEP_ACT = 9,         /* dz = (acc [+ hold]) * dG/dz(z) -> slabs / HBM                     */
  NG_BSTEP_PLAIN = 10       /* dz = acc -> slabs (NG_F_HOLD_SAVE: kept in registers too;          */
                            /* NG_F_SIGMA: d(sigma_pre) -> column 0 of slab out_slab + n_slabs)  */
};

enum {
  NG_F_SIGMA = 1,
  NG_F_HOLD_SAVE = 2,
  NG_F_HOLD_ADD = 4,
  NG_F_DIRECT = 8,          /* backward: dz goes straight to the HBM stash, no shared-memory slab */
  NG_F_FIRST_LAYER = 16     /* backward: this step differentiates the network's first layer: the  */
                            /* input gradient uses its fp32 weights (w1_off, gen_col0) directly   */
};

typedef struct {
  int8_t kind;
  int8_t wait_lag;
  int8_t n_slabs;           /* 64-column groups this step processes                             */
  int8_t out_slab;          /* first shared-memory slab written (-1: none)                      */
  int8_t res_slab;          /* NG_STEP_LINEAR: slab added to the result (-1: none)               */
  int8_t skip_src;          /* 0 none, 1 position, 2 direction                                  */
  int8_t flags;
  int8_t reserved;
  int16_t src_col;          /* first accumulator (TMEM) column read                             */
  int16_t sigma_col;        /* forward: TMEM column of the density pre-activation (NG_F_SIGMA);  */
                            /* backward: first of the 64 TMEM columns in which NG_F_HOLD_SAVE    */
                            /* parks the residual-path gradient for the NG_F_HOLD_ADD step        */
  int32_t bias_off;         /* packed floats: bias[64 * n_slabs]              (-1: none)        */
  int32_t coef_off;         /* packed floats: Gaussian coefficient -(1/s^2 + 1e-6) log2(e)      */
  int32_t skip_off;         /* packed floats: skip weights, [3][64 * n_slabs]  (-1: none)        */
  int32_t y_stash;          /* per-tile slab index the written slabs are copied to (-1: none):   */
                            /* forward: activation stash; backward: gradient (dz) stash          */
  int32_t z_stash;          /* per-tile slab index in the z stash: written (forward) / read      */
  int32_t gen_col0;         /* NG_STEP_GEN: first column of the first layer                      */
} NgStep;

typedef struct {
  int16_t tmem_col;         /* first accumulator column (0..511)                                */
  int16_t n;                /* MMA N (multiple of 16, <= 256)                                   */
  int16_t row0;             /* first row of the weight image used as B (multiple of 8)          */
  int16_t reserved;
} NgBlock;

typedef struct {
  int8_t n_chunks;          /* 0: no MMA (the op only marks a point in the completion order)    */
  int8_t n_blocks;          /* 1 or 2 N-blocks per K step                                       */
  int8_t accumulate;        /* 1: the first MMA adds to what the accumulator columns hold       */
  int8_t early;             /* 1: step k is a Gaussian step that publishes its slabs one by one   */
                            /* (slab_ready barriers) and op k writes none of the accumulator     */
                            /* columns step k reads: chunk c is issued as soon as ITS slab is     */
                            /* published, i.e. the MMAs of op k run under the epilogue of step k  */
  int8_t a_slab[NG_MAX_CHUNKS];
  int8_t k16[NG_MAX_CHUNKS];     /* 16-wide K steps of every chunk (1..4)                        */
  int16_t w_rows;                /* rows of every weight image of this op (<= 256)               */
  int16_t reserved2;
  int32_t w_off[NG_MAX_CHUNKS];  /* image offsets in the packed weight buffer, 1024 B units      */
  NgBlock blocks[2];
} NgOp;

typedef struct {
  int32_t n_ops;
  int32_t y_slabs_per_tile;      /* forward: activation stash; backward: dz stash                */
  int32_t z_slabs_per_tile;      /* z stash (written by the forward, read by the backward)       */
  int32_t n_floats;              /* packed fp32 values (biases, coefficients, skip weights)      */
  int64_t w1_off, b1_off, g1_off;/* first layer inside the flat fp32 parameter buffer (floats):  */
  int32_t n1;                    /* weight (n1, 3), bias (n1), inverse std (n1)                  */
  int32_t aux_pos_stash;         /* forward: per-tile slab receiving bf16 xyz (columns 0..2)     */
  int32_t aux_dir_stash;         /* forward: per-tile slab receiving bf16 directions, -1: none   */
  float sigma_bias;              /* added to the density pre-activation (radiance: -1)           */
  NgOp ops[NG_MAX_OPS];
  NgStep steps[NG_MAX_OPS + 1];
} NgProgram;

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_GARF_H_ */
