// Fused GARF field forward on sm_100a: radiance network (reference garf/model_radiance.py:23-96,
// == barf/model_garf_radiance.py) and proposal network (garf/model_proposal.py:22-56) with the
// Gaussian activation of barf/gaussian.py:10-31, one persistent launch per network instead of
// 11 (4) cuBLAS GEMMs + activation kernels. Query positions x = o + t d are formed in registers
// (garf/model_garf.py:105,141 materialises them), the first layer runs in fp32 on the CUDA cores,
// every other Linear as tcgen05 bf16 MMAs on a 128-sample tile that never leaves the SM; see
// include/nerfb200_garf.h for the tile program and garf_kernels.cuh for the hand-off protocol.
#include "common.cuh"
#include "garf.h"
#include "garf_kernels.cuh"

namespace nerfb200 {
namespace {

using namespace tc;
using namespace garf;

struct GarfFwdParams {
  NgProgram prog;
  const uint8_t* wpack;
  const float* floats;       // packed biases / Gaussian coefficients / skip weights
  const float* params;       // flat fp32 master parameters (first layer)
  NbMlpInputs in;
  int N;
  float* out_sigma;
  float* out_rgb;            // NULL: density-only network
  uint8_t* y_stash;          // NULL: inference (no stashes, no copies)
  uint8_t* z_stash;
};

constexpr float kLog2e = 1.4426950408889634f;

__global__ void __launch_bounds__(kThreadsG, 1)
garf_fwd_kernel(const __grid_constant__ GarfFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GarfSmem sm(smem_raw);
  const NgProgram& prog = p.prog;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  const int n_ops = prog.n_ops;
  const bool training = (p.y_stash != nullptr);

  if (threadIdx.x == 0) sm.init_barriers();
  for (int i = threadIdx.x; i < prog.n_floats; i += blockDim.x) sm.floats[i] = p.floats[i];
  sm.load_tables(prog);
  if (warp == kMmaWarpG) tmem_alloc(sm.tmem_ptr, kTmemColsG);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *sm.tmem_ptr;

  if (warp == kProducerWarpG) {
    regs_helper();
    if (lane == 0) producer_loop(prog, p.wpack, sm, n_tiles);
  } else if (warp == kMmaWarpG) {
    regs_helper();
    mma_loop(prog, sm, tmem_base, n_tiles);
  } else if (warp == kStashWarpG) {
    regs_helper();
    if (training && lane == 0) stash_copier_loop(prog, sm, n_tiles, p.y_stash, prog.y_slabs_per_tile);
  } else if (warp >= kRowWarpsG) {
    regs_helper();      // the idle warp of the helper warpgroup
  } else {
    regs_row();
    // ---------------- row threads ----------------
    const int row = threadIdx.x & 127;                  // tile row = TMEM lane
    const int cq = threadIdx.x >> 7;                    // 16-column quarter of every 64-column slab
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t off0 = (uint32_t)row * 128u + ((uint32_t)((2 * cq) ^ (row & 7)) << 4);       // this thread's two
    const uint32_t off1 = (uint32_t)row * 128u + ((uint32_t)((2 * cq + 1) ^ (row & 7)) << 4);   // 16-byte chunks of a slab
    const uint32_t sec_off = off0 < off1 ? off0 : off1;   // the 32-byte sector holding both chunks
    const bool odd_row = (row & 1) != 0;
    const uint32_t slab_base = smem_u32(sm.slab(0));
    RowSync rs;
    uint32_t g0 = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      uint8_t* ztile = training ? p.z_stash + (size_t)tile * (size_t)prog.z_slabs_per_tile * NB_SLAB_BYTES : nullptr;

      // ---- query position / direction of every row (fp32, shared by all four column quarters) ----
      named_bar_sync(1, kRowThreadsG);        // every reader of the previous tile's rows is done
      if (cq == 0) {
        PeSample ps;
        load_sample(p.in, n, ps);
        sm.pos[row] = make_float4(ps.x[0], ps.x[1], ps.x[2], 0.f);
        sm.dir[row] = make_float4(ps.dir[0], ps.dir[1], ps.dir[2], 0.f);
      }
      named_bar_sync(1, kRowThreadsG);
      const float4 mp = sm.pos[row], md = sm.dir[row];
      if (training) {
        // bf16 xyz / direction slabs: X operands of the weight gradients of the first layer and of
        // the two concatenating layers (columns 0..2, zero elsewhere)
        uint8_t* ytile = p.y_stash + (size_t)tile * (size_t)prog.y_slabs_per_tile * NB_SLAB_BYTES;
        const uint32_t px0 = cq == 0 ? pack_bf16(mp.x, mp.y) : 0u, px1 = cq == 0 ? pack_bf16(mp.z, 0.f) : 0u;
        const uint32_t pd0 = cq == 0 ? pack_bf16(md.x, md.y) : 0u, pd1 = cq == 0 ? pack_bf16(md.z, 0.f) : 0u;
        if (prog.aux_pos_stash >= 0) {
          uint8_t* s = ytile + (size_t)prog.aux_pos_stash * NB_SLAB_BYTES;
          const uint32_t a[8] = {px0, px1, 0u, 0u, 0u, 0u, 0u, 0u};
          stg256_row(s + sec_off, odd_row, a);
        }
        if (prog.aux_dir_stash >= 0) {
          uint8_t* s = ytile + (size_t)prog.aux_dir_stash * NB_SLAB_BYTES;
          const uint32_t a[8] = {pd0, pd1, 0u, 0u, 0u, 0u, 0u, 0u};
          stg256_row(s + sec_off, odd_row, a);
        }
      }

      for (int k = 0; k <= n_ops; ++k) {
        const NgStep& st = sm.steps[k];
        const uint32_t g = g0 + (uint32_t)k;
        const int through = (int)g - 1 - st.wait_lag;
        const int kind = st.kind, nsl = st.n_slabs;
        // first-layer parameters of this lane's two columns: requested BEFORE the barrier waits (with 227 KB of
        // shared memory carved out there is no L1 to speak of; each of these is an L2 round trip)
        float gen_w[10];
        if (kind == NG_STEP_GEN) {
          const int col = st.gen_col0 + 64 * (warp & 1) + 2 * lane;
          const float* W = p.params + prog.w1_off + (long long)col * 3;
#pragma unroll
          for (int i = 0; i < 6; ++i) gen_w[i] = __ldg(W + i);
          gen_w[6] = __ldg(p.params + prog.b1_off + col);
          gen_w[7] = __ldg(p.params + prog.b1_off + col + 1);
          gen_w[8] = __ldg(p.params + prog.g1_off + col);
          gen_w[9] = __ldg(p.params + prog.g1_off + col + 1);
        }
        rs.acc_through(sm, through);
        if (training) rs.drain_through(sm, through);
        tcgen05_fence_after();

        if (kind == NG_STEP_GEN) {
          // first layer, fp32: this warp covers 16 rows x one slab of the 128-column block, a lane
          // owns two adjacent columns (weights in registers, positions broadcast from shared memory)
          const int sib = warp & 1, rg = warp >> 1;
          const float w00 = gen_w[0], w01 = gen_w[1], w02 = gen_w[2], w10 = gen_w[3], w11 = gen_w[4], w12 = gen_w[5];
          const float b0 = gen_w[6], b1 = gen_w[7], s0 = gen_w[8], s1 = gen_w[9];
          const float c0 = -(s0 * s0 + 1e-6f) * kLog2e, c1 = -(s1 * s1 + 1e-6f) * kLog2e;
          uint8_t* slab = sm.slab(st.out_slab + sib);
          uint32_t zp[16];
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            const int rr = rg * 16 + r;
            const float4 x = sm.pos[rr];
            const float z0 = fmaf(w02, x.z, fmaf(w01, x.y, fmaf(w00, x.x, b0)));
            const float z1 = fmaf(w12, x.z, fmaf(w11, x.y, fmaf(w10, x.x, b1)));
            const float y0 = ex2f(z0 * z0 * c0), y1 = ex2f(z1 * z1 * c1);
            *reinterpret_cast<uint32_t*>(slab + slab_offset((uint32_t)rr, (uint32_t)(2 * lane))) = pack_bf16(y0, y1);
            zp[r] = pack_bf16(z0, z1);
          }
          publish_step(sm, g, true, lane);
#ifndef NG_EXP_NO_Z
          if (training && st.z_stash >= 0) {
#else
          if (false) {
#endif
            uint8_t* zs = ztile + (size_t)(st.z_stash + sib) * NB_SLAB_BYTES;
#pragma unroll
            for (int r = 0; r < 16; ++r)
              *reinterpret_cast<uint32_t*>(zs + zstash_offset((uint32_t)(rg * 16 + r), (uint32_t)(2 * lane))) = zp[r];
          }
        } else if (kind == NG_STEP_ACT) {
          const int ncols = 64 * nsl;
          const float4* bias4 = reinterpret_cast<const float4*>(sm.floats + st.bias_off + 16 * cq);
          const float4* coef4 = reinterpret_cast<const float4*>(sm.floats + st.coef_off + 16 * cq);
          const bool has_skip = st.skip_src != 0;
          const float* skip = sm.floats + (has_skip ? st.skip_off : 0) + 16 * cq;
          const float sx = st.skip_src == 2 ? md.x : mp.x, sy = st.skip_src == 2 ? md.y : mp.y,
                      sz = st.skip_src == 2 ? md.z : mp.z;
          const uint32_t acc_q = tmem_lane + (uint32_t)(st.src_col + 16 * cq);
          const bool early_step = k < n_ops && sm.ops[k].early != 0;
#ifndef NG_EXP_NO_Z
          const bool z_now = training && st.z_stash >= 0;
#else
          const bool z_now = false;
#endif
          uint32_t va[16], vb[16];
          // one 16-column group: bias (+ rank-3 fp32 skip), Gaussian, bf16 packs of y (-> slab) and z (-> stash)
          auto group = [&](const uint32_t (&v)[16], int j) {
            uint32_t yp[8], zp[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b = bias4[16 * j + q], c = coef4[16 * j + q];
              float z0 = __uint_as_float(v[4 * q]) + b.x, z1 = __uint_as_float(v[4 * q + 1]) + b.y;
              float z2 = __uint_as_float(v[4 * q + 2]) + b.z, z3 = __uint_as_float(v[4 * q + 3]) + b.w;
              if (has_skip) {
                const float4 kx = *reinterpret_cast<const float4*>(skip + 64 * j + 4 * q);
                const float4 ky = *reinterpret_cast<const float4*>(skip + ncols + 64 * j + 4 * q);
                const float4 kz = *reinterpret_cast<const float4*>(skip + 2 * ncols + 64 * j + 4 * q);
                z0 = fmaf(kz.x, sz, fmaf(ky.x, sy, fmaf(kx.x, sx, z0)));
                z1 = fmaf(kz.y, sz, fmaf(ky.y, sy, fmaf(kx.y, sx, z1)));
                z2 = fmaf(kz.z, sz, fmaf(ky.z, sy, fmaf(kx.z, sx, z2)));
                z3 = fmaf(kz.w, sz, fmaf(ky.w, sy, fmaf(kx.w, sx, z3)));
              }
              const float y0 = ex2f(z0 * z0 * c.x), y1 = ex2f(z1 * z1 * c.y);
              const float y2 = ex2f(z2 * z2 * c.z), y3 = ex2f(z3 * z3 * c.w);
              yp[2 * q] = pack_bf16(y0, y1);
              yp[2 * q + 1] = pack_bf16(y2, y3);
              zp[2 * q] = pack_bf16(z0, z1);
              zp[2 * q + 1] = pack_bf16(z2, z3);
            }
            const uint32_t sb = slab_base + (uint32_t)(st.out_slab + j) * NB_SLAB_BYTES;
            sts128g(sb + off0, yp[0], yp[1], yp[2], yp[3]);
            sts128g(sb + off1, yp[4], yp[5], yp[6], yp[7]);
            if (early_step) publish_slab(sm, st.out_slab + j, lane);   // the op's chunk on this slab may go
            // Early steps publish without a proxy fence in the row warps (the consumers fence), so the z stores
            // leave right away instead of sitting in 32 registers until the step is published. (Behind a step
            // that is not early, the writer-side fence of publish_step would wait for their acknowledgement:
            // slower, still correct — every Gaussian step of the two GARF programs is early.)
            if (z_now) stg256(ztile + (size_t)(st.z_stash + j) * NB_SLAB_BYTES + sec_off, zp);
          };
          // the TMEM load of group j + 1 is in flight during the math of group j
          tmem_ld16(acc_q, va);
#pragma unroll
          for (int j = 0; j < 4; j += 2) {
            if (j < nsl) {
              tmem_ld_wait16(va);
              if (j + 1 < nsl) tmem_ld16(acc_q + (uint32_t)(64 * (j + 1)), vb);
              group(va, j);
            }
            if (j + 1 < nsl) {
              tmem_ld_wait16(vb);
              if (j + 2 < nsl) tmem_ld16(acc_q + (uint32_t)(64 * (j + 2)), va);
              group(vb, j + 1);
            }
          }
          publish_step(sm, g, !early_step, lane);
        } else if (kind == NG_STEP_LINEAR) {
          const float* bias = sm.floats + st.bias_off + 16 * cq;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nsl) {
              uint32_t v[16], yp[8];
              tmem_ld16(tmem_lane + (uint32_t)(st.src_col + 64 * j + 16 * cq), v);
              tmem_ld_wait16(v);
              uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
              if (st.res_slab >= 0) {   // residual: the bf16 activations another layer left in a hold slab
                const uint8_t* rsb = sm.slab(st.res_slab + j);
                r0 = *reinterpret_cast<const uint4*>(rsb + off0);
                r1 = *reinterpret_cast<const uint4*>(rsb + off1);
              }
              const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const float a0 = __uint_as_float(v[i]) + bias[64 * j + i] + bf_lo(rr[i >> 1]);
                const float a1 = __uint_as_float(v[i + 1]) + bias[64 * j + i + 1] + bf_hi(rr[i >> 1]);
                yp[i >> 1] = pack_bf16(a0, a1);
              }
              const uint32_t sb = slab_base + (uint32_t)(st.out_slab + j) * NB_SLAB_BYTES;
              sts128g(sb + off0, yp[0], yp[1], yp[2], yp[3]);
              sts128g(sb + off1, yp[4], yp[5], yp[6], yp[7]);
            }
          }
          if ((st.flags & NG_F_SIGMA) && cq == 0) {
            uint32_t e[16];
            tmem_ld16(tmem_lane + (uint32_t)st.sigma_col, e);
            tmem_ld_wait16(e);
            const float pre = __uint_as_float(e[0]) + sm.floats[st.bias_off + 64 * nsl];
            if (valid) p.out_sigma[n] = softplus8(pre + prog.sigma_bias);
          }
          publish_step(sm, g, true, lane);
        } else if (kind == NG_STEP_RGB || kind == NG_STEP_SIGMA) {
          if (cq == 0) {
            uint32_t v[16];
            tmem_ld16(tmem_lane + (uint32_t)st.src_col, v);
            tmem_ld_wait16(v);
            const float* bias = sm.floats + st.bias_off;
            if (valid) {
              if (kind == NG_STEP_RGB) {
                p.out_rgb[n * 3 + 0] = sigmoidf(__uint_as_float(v[0]) + bias[0]);
                p.out_rgb[n * 3 + 1] = sigmoidf(__uint_as_float(v[1]) + bias[1]);
                p.out_rgb[n * 3 + 2] = sigmoidf(__uint_as_float(v[2]) + bias[2]);
              } else {
                p.out_sigma[n] = softplus8(__uint_as_float(v[0]) + bias[0] + prog.sigma_bias);
              }
            }
          }
          if (k < n_ops) publish_step(sm, g, false, lane);
        } else {
          if (k < n_ops) publish_step(sm, g, false, lane);
        }
      }
      g0 += (uint32_t)n_ops;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarpG) tmem_dealloc(tmem_base, kTmemColsG);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_garf_workspace_bytes(const void* program_host, long long n_samples,
                                             long long* y_stash_bytes, long long* z_stash_bytes) {
  NB_CHECK_ARG(program_host && n_samples >= 0, "garf_workspace_bytes: bad arguments");
  const NgProgram* prog = reinterpret_cast<const NgProgram*>(program_host);
  const long long n_tiles = (n_samples + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  if (y_stash_bytes) *y_stash_bytes = n_tiles * prog->y_slabs_per_tile * (long long)NB_SLAB_BYTES;
  if (z_stash_bytes) *z_stash_bytes = n_tiles * prog->z_slabs_per_tile * (long long)NB_SLAB_BYTES;
  return NERFB200_OK;
}

extern "C" int nerfb200_garf_fwd(const void* program_host, const void* wpack, const float* floats,
                                 const float* params, const NbMlpInputs* in_host, float* out_sigma,
                                 float* out_rgb, void* y_stash, void* z_stash, void* stream) {
  NB_CHECK_ARG(program_host && wpack && floats && params && in_host && out_sigma, "garf_fwd: null pointer");
  const NgProgram* prog = reinterpret_cast<const NgProgram*>(program_host);
  NB_CHECK_ARG(in_host->N >= 0 && in_host->S >= 1, "garf_fwd: bad shape N=%lld S=%d", (long long)in_host->N, in_host->S);
  NB_CHECK_ARG((y_stash == nullptr) == (z_stash == nullptr), "garf_fwd: the two stashes go together");
  int rc = garf::validate_garf_program(*prog, false);
  if (rc != NERFB200_OK) return rc;
  for (int k = 0; k <= prog->n_ops; ++k)
    NB_CHECK_ARG(prog->steps[k].kind != NG_STEP_RGB || out_rgb != nullptr, "garf_fwd: the program writes rgb but out_rgb is NULL");
  if (in_host->N == 0) return NERFB200_OK;

  GarfFwdParams p;
  p.prog = *prog;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack);
  p.floats = floats;
  p.params = params;
  p.in = *in_host;
  p.N = (int)in_host->N;
  p.out_sigma = out_sigma;
  p.out_rgb = out_rgb;
  p.y_stash = reinterpret_cast<uint8_t*>(y_stash);
  p.z_stash = reinterpret_cast<uint8_t*>(z_stash);

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(garf_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)garf::GarfSmem::bytes()));
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  garf_fwd_kernel<<<grid, garf::kThreadsG, garf::GarfSmem::bytes(), (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
