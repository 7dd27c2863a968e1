// Fused GARF field backward (data gradients) on sm_100a: the mirror image of garf_fwd.cu. One
// 128-sample tile walks the layers in reverse; dz lives in shared memory (A operand), transposed
// weight images stream through the TMA ring, dy accumulates in TMEM and the epilogue multiplies by
// the Gaussian's derivative dy/dz = -2 z v y (reference barf/gaussian.py:21-34), with z read back
// from the forward stash and y recomputed from it. Every dz tile also goes to HBM (bf16 slabs) for the
// weight-gradient kernel (mlp_wgrad.cu), which also reduces the bias and Gaussian-width gradients
// from the same slabs. Gradients w.r.t. the raw xyz / direction inputs (the path to the camera poses,
// garf/model_camera_calibration.py) are accumulated in fp32 in the epilogues of the three layers that
// see them.
#include "common.cuh"
#include "garf.h"
#include "garf_kernels.cuh"

namespace nerfb200 {
namespace {

using namespace tc;
using namespace garf;

struct GarfBwdParams {
  NgProgram prog;
  const uint8_t* wpack;      // transposed weight images
  const float* floats;       // packed Gaussian coefficients / skip weights (backward layout)
  const float* params;       // flat fp32 master parameters (first-layer weights for d(position))
  NbMlpInputs in;
  int N;
  const float* sigma;        // forward outputs
  const float* rgb;          // NULL: density-only network
  const float* g_sigma;      // upstream gradients (NULL = zero)
  const float* g_rgb;
  const uint8_t* z_stash;
  uint8_t* dy_stash;
  int want_input_grads;
  float* d_ray_o;            // (B,3) += (rays mode)
  float* d_ray_d;
  float* d_pos;              // (N,3) += (samples mode)
  float* d_dir;
};

constexpr float kTwoLn2 = 1.3862943611198906f;

__global__ void __launch_bounds__(kThreadsG, 1)
garf_bwd_kernel(const __grid_constant__ GarfBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GarfSmem sm(smem_raw);
  const NgProgram& prog = p.prog;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + NB_TILE_ROWS - 1) / NB_TILE_ROWS;
  const int n_ops = prog.n_ops;

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
    if (lane == 0) stash_copier_loop(prog, sm, n_tiles, p.dy_stash, prog.y_slabs_per_tile);
  } else if (warp >= kRowWarpsG) {
    regs_helper();      // the idle warp of the helper warpgroup
  } else {
    regs_row();
    const int row = threadIdx.x & 127;
    const int cq = threadIdx.x >> 7;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t off0 = (uint32_t)row * 128u + ((uint32_t)((2 * cq) ^ (row & 7)) << 4);
    const uint32_t off1 = (uint32_t)row * 128u + ((uint32_t)((2 * cq + 1) ^ (row & 7)) << 4);
    const uint32_t sec_off = off0 < off1 ? off0 : off1;   // the 32-byte sector holding both chunks
    const bool odd_row = (row & 1) != 0;
    const uint32_t slab_base = smem_u32(sm.slab(0));
    const bool want = p.want_input_grads != 0;
    RowSync rs;
    uint32_t g0 = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long n_raw = (long long)tile * NB_TILE_ROWS + row;
      const bool valid = n_raw < p.N;
      const long long n = valid ? n_raw : (long long)p.N - 1;
      const uint8_t* ztile = p.z_stash + (size_t)tile * (size_t)prog.z_slabs_per_tile * NB_SLAB_BYTES;
      uint8_t* dytile = p.dy_stash + (size_t)tile * (size_t)prog.y_slabs_per_tile * NB_SLAB_BYTES;

      // gradient w.r.t. the density pre-activation: softplus'(x) = sigmoid(x) = 1 - exp(-softplus(x))
      float dsp = 0.f;
      if (cq == 0 && valid && p.g_sigma != nullptr) {
        const float sg = p.sigma[n];
        dsp = p.g_sigma[n] * (sg > 8.f ? 1.f : -expm1f(-sg));
      }
      float4 mp = make_float4(0.f, 0.f, 0.f, 0.f), md = mp;
      float tq = 0.f;
      if (want) {
        PeSample ps;
        load_sample(p.in, n, ps);
        mp = make_float4(ps.x[0], ps.x[1], ps.x[2], 0.f);
        md = make_float4(ps.dir[0], ps.dir[1], ps.dir[2], 0.f);
        tq = p.in.t_mode == 0 ? ps.t0 : (ps.t0 + ps.t1) * 0.5f;
      }
      float dpos[3] = {0.f, 0.f, 0.f}, ddir[3] = {0.f, 0.f, 0.f};

      for (int k = 0; k <= n_ops; ++k) {
        const NgStep& st = sm.steps[k];
        const uint32_t g = g0 + (uint32_t)k;
        const int through = (int)g - 1 - st.wait_lag;
        const int kind = st.kind, nsl = st.n_slabs, flags = st.flags;
        // the forward's pre-activations of this layer: requested before the accumulator is waited for
        // (two 64-column groups in flight; the next two are requested while these are consumed)
        Sector32 zq[2];
        const uint8_t* zbase = ztile + (size_t)(st.z_stash < 0 ? 0 : st.z_stash) * NB_SLAB_BYTES;
        if (kind == NG_BSTEP_ACT) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (j < nsl) {
              zq[j] = ldg256_stream(zbase + (size_t)j * NB_SLAB_BYTES + sec_off);
            }
          }
        }
        // ... and the lines of the NEXT Gaussian step are pulled into L2 now (one 128-byte row segment per
        // slab and row, by the quarter-0 thread): its loads then cost an L2 hit instead of an HBM round trip
        if (cq == 0) {
          for (int kn = k + 1; kn <= n_ops; ++kn) {
            const NgStep& sn = sm.steps[kn];
            if (sn.kind != NG_BSTEP_ACT) continue;
            const uint8_t* zn = ztile + (size_t)sn.z_stash * NB_SLAB_BYTES + (uint32_t)row * 128u;
            for (int j = 0; j < sn.n_slabs; ++j)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(zn + (size_t)j * NB_SLAB_BYTES));
            break;
          }
        }
        rs.acc_through(sm, through);
        rs.drain_through(sm, through);
        tcgen05_fence_after();

        if (kind == NG_BSTEP_HEAD) {
          // gradients w.r.t. the pre-activations of the output layer -> columns 0.. of slab out_slab
          uint32_t h0 = 0u, h1 = 0u;
          if (cq == 0) {
            if (flags & NG_F_SIGMA) {
              h0 = pack_bf16(dsp, 0.f);
            } else if (valid && p.g_rgb != nullptr) {
              float d3[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float y = p.rgb[n * 3 + c];
                d3[c] = p.g_rgb[n * 3 + c] * y * (1.f - y);
              }
              h0 = pack_bf16(d3[0], d3[1]);
              h1 = pack_bf16(d3[2], 0.f);
            }
          }
          const uint32_t sb = slab_base + (uint32_t)st.out_slab * NB_SLAB_BYTES;
          sts128g(sb + off0, h0, h1, 0u, 0u);      // cq == 0: logical chunk 0 holds the head; zero elsewhere
          sts128g(sb + off1, 0u, 0u, 0u, 0u);
          publish_step(sm, g, true, lane);
        } else if (kind == NG_BSTEP_ACT) {
          const int ncols = 64 * nsl;
          const float4* coef4 = reinterpret_cast<const float4*>(sm.floats + st.coef_off + 16 * cq);
          const bool direct = (flags & NG_F_DIRECT) != 0;
          const bool early_step = k < n_ops && sm.ops[k].early != 0;
          const bool first = (flags & NG_F_FIRST_LAYER) != 0;
          const bool grads = want && (st.skip_src != 0);
          const bool add_hold = (flags & NG_F_HOLD_ADD) != 0;
          const float* skip = sm.floats + ((grads && !first) ? st.skip_off : 0) + 16 * cq;
          const uint32_t acc_q = tmem_lane + (uint32_t)(st.src_col + 16 * cq);
          float acc3[3] = {0.f, 0.f, 0.f};
          uint32_t va[16], vb[16];
          // The residual-path gradient (d(z1 + z2[:, :128]), packed bf16, this thread's 2 x 16 columns) was parked in
          // TMEM columns sigma_col .. + 63 by the NG_F_HOLD_SAVE step: sixteen registers held across ten steps cost
          // the backward 9 % (spills, 2.65 -> 2.42 ms without them). TMEM lane = tile row, so every thread reads
          // back exactly what it wrote; the compiler guarantees no op in between touches those columns.
          auto group = [&](uint32_t (&v)[16], int j) {
            uint32_t dp[8];
            if (add_hold && j < 2) {      // fetched where it is used and added to the accumulator values up front:
              uint32_t hold[8];           // the element loop below carries no trace of it for the other steps
              tmem_ld8(tmem_lane + (uint32_t)(st.sigma_col + 32 * j + 8 * cq), hold);
              tmem_ld_wait8(hold);
#pragma unroll
              for (int i = 0; i < 16; ++i)
                v[i] = __float_as_uint(__uint_as_float(v[i]) + ((i & 1) ? bf_hi(hold[i >> 1]) : bf_lo(hold[i >> 1])));
            }
            uint32_t zz[8];      // the z stash keeps a thread's 16 columns in natural order (garf_kernels.cuh)
#pragma unroll
            for (int i = 0; i < 8; ++i) zz[i] = zq[j & 1].w[i];
            if (j + 2 < nsl)       // the slot is free again: request the group after the next
              zq[j & 1] = ldg256_stream(zbase + (size_t)(j + 2) * NB_SLAB_BYTES + sec_off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 c = coef4[16 * j + q];
              const float cc[4] = {c.x, c.y, c.z, c.w};
              float dz[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * q + e;
                const float ga = __uint_as_float(v[i]);
                const float z = (e & 1) ? bf_hi(zz[i >> 1]) : bf_lo(zz[i >> 1]);
                const float t = z * cc[e];
                dz[e] = ga * ex2f(z * t) * (t * kTwoLn2);                 // g * y * (-2 v z)
              }
              dp[2 * q] = pack_bf16(dz[0], dz[1]);
              dp[2 * q + 1] = pack_bf16(dz[2], dz[3]);
              if (grads) {
                if (first) {
                  const float* W = p.params + prog.w1_off + (long long)(st.gen_col0 + 64 * j + 16 * cq + 4 * q) * 3;
#pragma unroll
                  for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int c3 = 0; c3 < 3; ++c3) acc3[c3] = fmaf(dz[e], __ldg(W + 3 * e + c3), acc3[c3]);
                } else {
#pragma unroll
                  for (int c3 = 0; c3 < 3; ++c3) {
                    const float4 k = *reinterpret_cast<const float4*>(skip + c3 * ncols + 64 * j + 4 * q);
                    acc3[c3] = fmaf(dz[3], k.w, fmaf(dz[2], k.z, fmaf(dz[1], k.y, fmaf(dz[0], k.x, acc3[c3]))));
                  }
                }
              }
            }
            if (direct) {
              uint8_t* ds = dytile + (size_t)(st.y_stash + j) * NB_SLAB_BYTES;
              stg256_row(ds + sec_off, odd_row, dp);
            } else {
              const uint32_t sb = slab_base + (uint32_t)(st.out_slab + j) * NB_SLAB_BYTES;
              sts128g(sb + off0, dp[0], dp[1], dp[2], dp[3]);
              sts128g(sb + off1, dp[4], dp[5], dp[6], dp[7]);
              if (early_step) publish_slab(sm, st.out_slab + j, lane);   // the op's chunk on this slab may go
            }
          };
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
          if (grads) {
            if (st.skip_src == 2) { ddir[0] += acc3[0]; ddir[1] += acc3[1]; ddir[2] += acc3[2]; }
            else { dpos[0] += acc3[0]; dpos[1] += acc3[1]; dpos[2] += acc3[2]; }
          }
          if (k < n_ops) publish_step(sm, g, !direct && !early_step, lane);
        } else if (kind == NG_BSTEP_PLAIN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nsl) {
              uint32_t v[16], dp[8];
              tmem_ld16(tmem_lane + (uint32_t)(st.src_col + 64 * j + 16 * cq), v);
              tmem_ld_wait16(v);
#pragma unroll
              for (int i = 0; i < 16; i += 2) dp[i >> 1] = pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
              if ((flags & NG_F_HOLD_SAVE) && j < 2)      // parked in TMEM until the NG_F_HOLD_ADD step (see there)
                tmem_st8(tmem_lane + (uint32_t)(st.sigma_col + 32 * j + 8 * cq), dp);
              const uint32_t sb = slab_base + (uint32_t)(st.out_slab + j) * NB_SLAB_BYTES;
              sts128g(sb + off0, dp[0], dp[1], dp[2], dp[3]);
              sts128g(sb + off1, dp[4], dp[5], dp[6], dp[7]);
            }
          }
          if (flags & NG_F_SIGMA) {   // d(sigma_pre) = column 0 of the slab behind the main ones (a 16-wide K step)
            const uint32_t sb = slab_base + (uint32_t)(st.out_slab + nsl) * NB_SLAB_BYTES;
            sts128g(sb + off0, cq == 0 ? pack_bf16(dsp, 0.f) : 0u, 0u, 0u, 0u);
            sts128g(sb + off1, 0u, 0u, 0u, 0u);
          }
          if (flags & NG_F_HOLD_SAVE) tmem_st_wait();
          publish_step(sm, g, true, lane);
        } else {
          if (k < n_ops) publish_step(sm, g, false, lane);
        }
      }
      g0 += (uint32_t)n_ops;

      // ---- d(position), d(direction) of the samples -> rays (the four quarters add up) ----
      if (want) {
        if (p.in.pos != nullptr) {
          if (valid) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              atomicAdd(p.d_pos + n * 3 + c, dpos[c]);
              if (p.d_dir != nullptr) atomicAdd(p.d_dir + n * 3 + c, ddir[c]);
            }
          }
        } else {
          float go[3], gd[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            go[c] = valid ? dpos[c] : 0.f;
            gd[c] = valid ? (tq * dpos[c] + ddir[c]) : 0.f;
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
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarpG) tmem_dealloc(tmem_base, kTmemColsG);
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_garf_bwd(const void* program_host, const void* wpack_t, const float* floats,
                                 const float* params, const NbMlpInputs* in_host, const float* sigma,
                                 const float* rgb, const float* g_sigma, const float* g_rgb,
                                 const void* z_stash, void* dy_stash, int want_input_grads,
                                 float* d_ray_o, float* d_ray_d, float* d_pos, float* d_dir, void* stream) {
  NB_CHECK_ARG(program_host && wpack_t && floats && params && in_host, "garf_bwd: null pointer");
  NB_CHECK_ARG(sigma && z_stash && dy_stash, "garf_bwd: null buffer");
  const NgProgram* prog = reinterpret_cast<const NgProgram*>(program_host);
  if (want_input_grads) {
    if (in_host->pos != nullptr) NB_CHECK_ARG(d_pos != nullptr, "garf_bwd: d_pos required");
    else NB_CHECK_ARG(d_ray_o && d_ray_d, "garf_bwd: d_ray_o / d_ray_d required");
  }
  int rc = garf::validate_garf_program(*prog, true);
  if (rc != NERFB200_OK) return rc;
  for (int k = 0; k <= prog->n_ops; ++k) {
    const NgStep& st = prog->steps[k];
    if (st.kind == NG_BSTEP_HEAD && !(st.flags & NG_F_SIGMA))
      NB_CHECK_ARG(rgb != nullptr || g_rgb == nullptr, "garf_bwd: rgb outputs required for their gradient");
  }
  if (in_host->N == 0) return NERFB200_OK;

  GarfBwdParams p;
  p.prog = *prog;
  p.wpack = reinterpret_cast<const uint8_t*>(wpack_t);
  p.floats = floats;
  p.params = params;
  p.in = *in_host;
  p.N = (int)in_host->N;
  p.sigma = sigma;
  p.rgb = rgb;
  p.g_sigma = g_sigma;
  p.g_rgb = g_rgb;
  p.z_stash = reinterpret_cast<const uint8_t*>(z_stash);
  p.dy_stash = reinterpret_cast<uint8_t*>(dy_stash);
  p.want_input_grads = want_input_grads ? 1 : 0;
  p.d_ray_o = d_ray_o;
  p.d_ray_d = d_ray_d;
  p.d_pos = d_pos;
  p.d_dir = d_dir;

  static bool configured = false;
  if (!configured) {
    NB_CHECK_CUDA(cudaFuncSetAttribute(garf_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)garf::GarfSmem::bytes()));
    configured = true;
  }
  const int n_tiles = ceil_div(p.N, NB_TILE_ROWS);
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  garf_bwd_kernel<<<grid, garf::kThreadsG, garf::GarfSmem::bytes(), (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
