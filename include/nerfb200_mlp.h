/*
 * nerfb200_mlp.h — tile-program description shared by the host side and the fused MLP kernels
 * (part of the C ABI of libnerfb200.so, included by nerfb200.h).
 *
 * A network is compiled (by the host mirror, nerf_experiments_b200/mlp_program.py) into a short
 * program of GEMM ops executed on one 128-sample tile that stays on-chip:
 *   A operand  = up to NB_MAX_CHUNKS 64-wide K chunks taken from shared-memory slabs
 *   B operand  = one packed bf16 weight image per K chunk, streamed from L2 by the TMA engine
 *   D          = up to NB_MAX_BLOCKS N-blocks of the TMEM accumulator
 *   epilogue   = what the 128 row-threads do with D (activation, store as next A, outputs)
 * The same structure describes forward (A = activations, B = W) and backward-data
 * (A = dY, B = W^T) passes.
 */
#ifndef NERFB200_MLP_H_
#define NERFB200_MLP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  NB_MAX_OPS = 32,
  NB_MAX_CHUNKS = 12,
  NB_MAX_BLOCKS = 3,
  NB_TILE_ROWS = 128,
  NB_SLAB_BYTES = 128 * 128,          /* 128 rows x 64 bf16 */
  NB_N_SLABS = 6,                     /* 0..3 activations, 4 and 5 auxiliary (PE / extra) */
  NB_RING_STAGE_BYTES = 256 * 128,    /* largest weight image: 256 rows x 64 bf16 */
  NB_MAX_RING_STAGES = 4
};

/* forward epilogues */
enum {
  NB_EPI_RELU = 0,        /* y = relu(acc+b) -> act slabs                                      */
  NB_EPI_LINEAR = 1,      /* y = acc+b -> act slabs                                            */
  NB_EPI_LINEAR_SIGMA = 2,/* as LINEAR, and sigma = softplus8(extra block col 0 + b) -> out    */
  NB_EPI_RGB = 3,         /* rgb = sigmoid(acc[0:3]+b) -> out                                  */
  NB_EPI_RGB_SIGMA = 4,   /* as RGB, and sigma = softplus8(acc[3]+b[3]) -> out (delayed dens.) */
  NB_EPI_RELU_SIGMA = 5   /* y = relu(acc+b) -> act slabs, sigma from extra block              */
};

/* backward (data-gradient) epilogues */
enum {
  NB_BEPI_MASK = 0,       /* dY_prev = acc * relu_mask(bits) -> act slabs (+dY stash)          */
  NB_BEPI_PLAIN = 1,      /* dY_prev = acc -> act slabs (+dY stash)                            */
  NB_BEPI_PLAIN_SIGMA = 2,/* as PLAIN and d(sigma_pre) -> aux slab 4 col 0                     */
  NB_BEPI_MASK_SIGMA = 3, /* as MASK and d(sigma_pre) -> aux slab 4 col 0                      */
  NB_BEPI_NONE = 4,       /* nothing to store                                                  */
  NB_BEPI_PEGRAD_POS = 5, /* acc = d(position encoding) in the canonical column order (below): */
  NB_BEPI_PEGRAD_DIR = 6  /* pushed through the encoder into d(position) / d(direction)        */
};

/* Canonical column order of an encoding-gradient accumulator (NB_BEPI_PEGRAD_*): the gradient
 * w.r.t. cos(coordinate c, level j) sits in column 2*(NB_PE_CANON_LEVELS*c + j), the one w.r.t.
 * sin(c, j) in the next column, the identity columns in 60..62 — whatever the encoder's real
 * level count, so that the kernel can keep the 64 columns in registers. */
enum { NB_PE_CANON_LEVELS = 10, NB_PE_CANON_COLS = 64, NB_PE_CANON_IDENTITY = 60 };

typedef struct {
  int16_t tmem_col;   /* first accumulator column inside the op's 256-column buffer; values   */
                      /* >= 256 address column (tmem_col - 256) of the OTHER buffer            */
  int16_t n;          /* MMA N (multiple of 16, <= 256)                                        */
  int16_t row0;       /* first row of the weight image used as B (multiple of 8)               */
  int16_t accum_in;   /* 1: accumulate onto what an earlier op left in these columns           */
} NbBlock;

typedef struct {
  int8_t n_chunks;
  int8_t n_blocks;
  int8_t epi;
  int8_t out_chunks;                 /* 64-wide slabs the epilogue writes (act slabs 0..)       */
  int8_t a_src[NB_MAX_CHUNKS];       /* slab id of every K chunk                                */
  int8_t k16[NB_MAX_CHUNKS];         /* 16-wide MMA K steps in every chunk (1..4)               */
  int8_t blk_mask[NB_MAX_CHUNKS];    /* bit b set: this chunk's image feeds N-block b           */
  int8_t n_sub[NB_MAX_CHUNKS];       /* the chunk spans n_sub consecutive A slabs (a_src, +1..): */
                                     /* n_sub images of w_rows rows each, back to back          */
  int16_t w_rows[NB_MAX_CHUNKS];     /* rows of the weight image of every chunk                 */
  int32_t w_off[NB_MAX_CHUNKS];      /* offset of the image in the packed buffer, 1024 B units  */
  NbBlock blocks[NB_MAX_BLOCKS];
  int32_t bias_off;                  /* float index into the packed bias buffer (-1: none)      */
  int32_t stash_slab;                /* first stash slab (per tile) of the output, -1: none     */
  int32_t mask_word;                 /* first mask word row-group (per tile), -1: none          */
  int32_t out_width;                 /* real (unpadded) number of output features               */
} NbOp;

typedef struct {
  int32_t n_ops;
  int32_t stash_slabs_per_tile;      /* slabs of NB_SLAB_BYTES per tile in the stash            */
  int32_t mask_words_per_tile;       /* 32-column groups per tile (x128 rows x 4 B)             */
  int16_t n_slabs;                   /* shared-memory slabs in use: 5 or 6                      */
  int16_t n_stages;                  /* weight ring depth: 4 with 5 slabs, 3 with 6             */
  NbOp ops[NB_MAX_OPS];
} NbProgram;

/* positional encodings (reference barf/positional_encodings.py) */
enum { NB_PE_IDENTITY = 0, NB_PE_FOURIER = 1, NB_PE_INTEGRATED = 2 };

typedef struct {
  int32_t kind;                 /* NB_PE_*                                                    */
  int32_t levels;               /* L (0 allowed: identity only)                               */
  int32_t include_identity;
  int32_t use_mask;             /* BARF coarse-to-fine mask from *alpha                       */
  int32_t distribute_variance;  /* integrated PE only                                         */
  float scale;
  float pixel_width_sigma;      /* integrated PE only                                         */
  int32_t slab;                 /* destination slab id (4 or 5), -1: encoder unused           */
  int32_t stash_slab;           /* stash slab (per tile) of the encoding, -1: none            */
  int32_t encode_before_op;     /* forward: 0 = encoded at tile start; k > 0 = encoded while  */
                                /* op k's MMAs run (the slab is shared with an encoder whose  */
                                /* last use is op k-1 and this encoder's first use is > k)    */
} NbPeCfg;

/* Where the samples of a launch come from.  Either per-ray data (positions are computed in the
 * kernel: x = o + t_q d, reference barf/model_interpolation.py:279-312) or per-sample pos/dir
 * (the NerfModel.forward signature, barf/model_interpolation_architecture.py:96-104). */
typedef struct {
  int64_t N;                 /* samples = rays * S                                          */
  int32_t S;                 /* samples per ray; ray index of sample n is n / S               */
  int32_t t_mode;            /* 0: "left" (t_q = t_start), 1: "middle" ((t_start+t_end)/2)    */
  const float* ray_o;        /* (B,3) used when pos == NULL                                   */
  const float* ray_d;        /* (B,3)                                                         */
  const float* t_start;      /* (N,) or NULL                                                  */
  const float* t_end;        /* (N,) or NULL                                                  */
  const float* pixel_width;  /* (B,) / (N,) or NULL                                           */
  const float* pos;          /* (N,3) or NULL                                                 */
  const float* dir;          /* (N,3) or NULL (with pos)                                      */
  int32_t pixel_width_per_sample;
  int32_t reserved;
} NbMlpInputs;

/* A packed weight image: value(r,c) = params[base + r*row_stride + c*col_stride] for
 * r < n_rows, c < n_cols, zero elsewhere, is written for r < rows_padded to image row
 * dst_row0 + r*dst_row_step (rows of 64 bf16, swizzled slab layout). Rows no descriptor
 * writes keep the zeros the packed buffer was initialised with. */
typedef struct {
  int64_t base;
  int32_t row_stride, col_stride;
  int32_t n_rows, n_cols;
  int32_t rows_padded;
  int32_t dst_off;              /* 1024 B units into the packed buffer                        */
  int32_t dst_row0;             /* first image row written                                    */
  int32_t dst_row_step;         /* image-row step (>= 1)                                      */
  int32_t img_rows;             /* 0: one [rows][64] image in the 128B-swizzled slab layout;  */
                                /* > 0: four [img_rows][16] images (one per 16-column K step, */
                                /* img_rows * 32 B apart) in the 32B-swizzled K-major layout  */
                                /* the two-tile forward kernel streams one MMA at a time      */
  int32_t reserved;
} NbPackChunk;

/* A packed fp32 segment: dst[dst_off + i] = f(params[base + i * stride]) for i < n, 0 for
 * n <= i < n_padded; f = identity (NB_PACK_COPY) or the exponent coefficient of the Gaussian
 * activation, -(s^2 + 1e-6) log2(e) (NB_PACK_GAUSS: y = exp2(z^2 * coefficient), barf/gaussian.py:10-19) */
enum { NB_PACK_COPY = 0, NB_PACK_GAUSS = 1 };
typedef struct {
  int64_t base;
  int32_t n, n_padded;
  int32_t dst_off;
  int32_t kind;                 /* NB_PACK_*                                                   */
  int32_t stride;               /* element stride in params (0 is read as 1)                   */
  int32_t reserved;
} NbPackBias;

/* One work item of the weight-gradient kernel: dW[m0:m0+m_real, c0:c0+n_real] +=
 * dY[tiles, dy slabs]^T X[tiles, x slabs] over the tile range [tile_begin, tile_end). */
typedef struct {
  int32_t tile_begin, tile_end;
  int32_t n_dy_slabs;           /* 1..4 consecutive dY stash slabs (64 output features each)  */
  int32_t n_x_slabs;            /* 1..4 consecutive X stash slabs (64 input columns each)     */
  int32_t dy_slab, x_slab;      /* first slab inside a tile's dY / X stash                    */
  int32_t m_real, n_real;       /* real output features / input columns covered               */
  int64_t dst;                  /* float index of dW[m0, c0] in the flat gradient buffer      */
  int32_t ld;                   /* row stride of W (= in_features)                            */
  int32_t bias_dst;             /* float index of the bias gradient of output feature m0 (the */
                                /* column sums of the dY slabs, += over the tile range), -1:  */
                                /* another item of the same dY slabs carries it               */
  int32_t mode;                 /* NB_WGRAD_MMA, or NB_WGRAD_COLSUM: no weight block, only the */
                                /* column sums of the z duty below                             */
  int32_t coef_dst;             /* z duty: float index of the inverse-std parameter (and of   */
                                /* its gradient) of output feature m0, -1: none               */
  /* "z duty" (GARF): the item also streams n_z_slabs slabs of the Z stash (pre-activations of a      */
  /* Gaussian layer) — z slab j belongs to the item's dY slab z_first + j — and reduces, for those    */
  /* dY slabs, sum dz (bias gradient, at zbias_dst) and sum z*dz (Gaussian width, at coef_dst) per    */
  /* column while the slabs sit in shared memory for the MMAs. 2*ceil(n_dy/2) + n_x + n_z <= 9.       */
  /* NB_WGRAD_COLSUM items are z duty without a weight block: n_x_slabs = 0, n_z_slabs = n_dy_slabs.  */
  int32_t z_slab;               /* first slab inside a tile's Z stash, -1: no z duty          */
  int32_t n_z_slabs;            /* 0..4                                                        */
  int32_t z_first;              /* index (within the item) of the dY slab of z slab 0          */
  int32_t zbias_dst;            /* float index of the bias gradient of output feature m0, -1   */
  int32_t x2_slab;              /* >= 0: the item's LAST X slab is stash slab x2_slab instead  */
                                /* of x_slab + n_x_slabs - 1 (a concatenating layer whose two  */
                                /* input sources sit apart in the stash: one item, dY read     */
                                /* once); -1: the X slabs are consecutive                      */
} NbWgradItem;
enum { NB_WGRAD_MMA = 0, NB_WGRAD_COLSUM = 1 };

#ifdef __cplusplus
}
#endif
#endif /* NERFB200_MLP_H_ */
