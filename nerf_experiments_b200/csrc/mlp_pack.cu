// Refreshes the bf16 weight images (and padded fp32 biases) the fused MLP kernels consume from
// the fp32 master parameters. One launch per optimiser step; every image is written directly
// in the 128B-swizzled slab layout so that a single cp.async.bulk brings it into shared memory.
#include "common.cuh"
#include "mlp.h"
#include "tc.cuh"

namespace nerfb200 {
namespace {

__global__ void __launch_bounds__(256)
mlp_pack_kernel(const float* __restrict__ params, const NbPackChunk* __restrict__ chunks,
                int n_chunks, uint8_t* __restrict__ wpack, const NbPackBias* __restrict__ biases,
                int n_biases, float* __restrict__ bias_out) {
  // blockIdx.y = descriptor; threads cover (row, 16-byte piece)
  const int d = blockIdx.y;
  if (d < n_chunks) {
    const NbPackChunk ch = chunks[d];
    uint8_t* dst = wpack + (size_t)ch.dst_off * 1024u;
    const int pieces = ch.rows_padded * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pieces; i += gridDim.x * blockDim.x) {
      const int r = i >> 3, q = i & 7;
      uint32_t w[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = q * 8 + h * 2 + e;
          v[e] = (r < ch.n_rows && c < ch.n_cols)
                     ? params[ch.base + (long long)r * ch.row_stride + (long long)c * ch.col_stride]
                     : 0.f;
        }
        w[h] = tc::pack_bf16(v[0], v[1]);
      }
      const int R = ch.dst_row0 + r * ch.dst_row_step;   // image row
      if (ch.img_rows > 0) {
        // one [img_rows][16] image per K step, rows of 32 B, 16-byte halves swizzled with row bit 2
        // (canonical UMMA SWIZZLE_32B K-major layout)
        const uint32_t off = (uint32_t)(q >> 1) * (uint32_t)ch.img_rows * 32u + (uint32_t)R * 32u +
                             ((uint32_t)((q & 1) ^ ((R >> 2) & 1)) << 4);
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(w[0], w[1], w[2], w[3]);
      } else {
        *reinterpret_cast<uint4*>(dst + (size_t)R * 128u + ((uint32_t)(q ^ (R & 7)) << 4)) =
            make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  } else if (d - n_chunks < n_biases) {
    const NbPackBias b = biases[d - n_chunks];
    const long long stride = b.stride > 0 ? b.stride : 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < b.n_padded; i += gridDim.x * blockDim.x) {
      float v = (i < b.n) ? params[b.base + (long long)i * stride] : 0.f;
      if (b.kind == NB_PACK_GAUSS && i < b.n) v = -(v * v + 1e-6f) * 1.4426950408889634f;
      bias_out[b.dst_off + i] = v;
    }
  }
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_mlp_pack(const float* params, const NbPackChunk* chunks_dev, int n_chunks,
                                 void* wpack, const NbPackBias* biases_dev, int n_biases,
                                 float* bias_out, void* stream) {
  NB_CHECK_ARG(params && n_chunks >= 0 && n_biases >= 0, "mlp_pack: bad arguments");
  NB_CHECK_ARG(n_chunks == 0 || (chunks_dev && wpack), "mlp_pack: null chunk buffers");
  NB_CHECK_ARG(n_biases == 0 || (biases_dev && bias_out), "mlp_pack: null bias buffers");
  if (n_chunks + n_biases == 0) return NERFB200_OK;
  dim3 grid(2, n_chunks + n_biases);
  mlp_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      params, chunks_dev, n_chunks, reinterpret_cast<uint8_t*>(wpack), biases_dev, n_biases, bias_out);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
