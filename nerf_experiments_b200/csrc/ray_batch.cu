// GPU-resident ray batcher: builds one training batch — ray origins and
// directions with the raw and the noisy camera poses, target colours (optionally the blur-pyramid
// interpolation), image index and pixel width — from flat ray indices, in one launch.
// Replaces ImagePoseDataset.__getitem__ + the DataLoader's collate + the host-to-device copy
// (reference barf/dataset.py:613-637, ray generation :407-482) and
// ImagePoseDataModule.get_blurred_pixel_colors (barf/data_module.py:276-369): the reference
// stores every ray direction of every image (N*H*W*3 floats, twice) and indexes them from Python
// one ray at a time; here the direction is recomputed from the pixel index in registers.
#include "common.cuh"

namespace nerfb200 {
namespace {

struct RayBatchParams {
  const long long* ray_index;   // (B) flat index = image * H*W + pixel
  int B;
  const float* c2w_raw;         // (N,4,4) row-major
  const float* c2w_noisy;       // (N,4,4)
  const float* images;          // (N, H*W, n_sigmas, 3)
  const int* image_id_map;      // (N) dataset image -> global image index, or NULL (identity)
  int n_images, H, W, n_sigmas;
  float focal, pixel_width;
  int blur_low, blur_high;      // blur_low < 0: copy all n_sigmas levels
  float blur_coef;
  float* o_raw; float* o_noisy; float* d_raw; float* d_noisy;   // (B,3)
  float* colors;                // (B,2,3) or (B,n_sigmas,3)
  long long* img_idx;           // (B)
  float* pixel_width_out;       // (B)
};

__global__ void __launch_bounds__(256) ray_batch_kernel(const RayBatchParams p) {
  const long long hw = (long long)p.H * p.W;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < p.B; b += gridDim.x * blockDim.x) {
    const long long idx = p.ray_index[b];
    const int img = (int)(idx / hw);                  // dataset.py:615
    const long long pix = idx - (long long)img * hw;  // dataset.py:627
    const int i = (int)(pix / p.W), j = (int)(pix - (long long)i * p.W);
    // pixel-centre direction in camera space, camera looks along -z, y flipped (dataset.py:443-451)
    const float x = __fdiv_rn((float)j - 0.5f * (float)(p.W - 1), p.focal);
    const float y = -__fdiv_rn((float)i - 0.5f * (float)(p.H - 1), p.focal);
    const float norm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), 1.f));
    const float m[3] = {__fdiv_rn(x, norm), __fdiv_rn(y, norm), __fdiv_rn(-1.f, norm)};
    const float* Pr = p.c2w_raw + (size_t)img * 16;
    const float* Pn = p.c2w_noisy + (size_t)img * 16;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      // world direction = R m (dataset.py:479-482), origin = translation column (:401)
      p.d_raw[(size_t)b * 3 + r] = __ldg(Pr + 4 * r) * m[0] + __ldg(Pr + 4 * r + 1) * m[1] + __ldg(Pr + 4 * r + 2) * m[2];
      p.d_noisy[(size_t)b * 3 + r] = __ldg(Pn + 4 * r) * m[0] + __ldg(Pn + 4 * r + 1) * m[1] + __ldg(Pn + 4 * r + 2) * m[2];
      p.o_raw[(size_t)b * 3 + r] = __ldg(Pr + 4 * r + 3);
      p.o_noisy[(size_t)b * 3 + r] = __ldg(Pn + 4 * r + 3);
    }
    const float* c = p.images + ((size_t)img * hw + pix) * p.n_sigmas * 3;
    if (p.blur_low < 0) {
      for (int k = 0; k < p.n_sigmas * 3; ++k) p.colors[(size_t)b * p.n_sigmas * 3 + k] = ld_stream(c + k);
    } else {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float lo = ld_stream(c + p.blur_low * 3 + ch), hi = ld_stream(c + p.blur_high * 3 + ch);
        // interpolation = c[low] * coef + c[high] * (1 - coef)   (data_module.py:356)
        p.colors[(size_t)b * 6 + ch] = __fadd_rn(__fmul_rn(lo, p.blur_coef), __fmul_rn(hi, 1.f - p.blur_coef));
        p.colors[(size_t)b * 6 + 3 + ch] = ld_stream(c + (p.n_sigmas - 1) * 3 + ch);   // original pixel
      }
    }
    p.img_idx[b] = p.image_id_map ? (long long)p.image_id_map[img] : (long long)img;
    p.pixel_width_out[b] = p.pixel_width;
  }
}

}  // namespace
}  // namespace nerfb200

using namespace nerfb200;

extern "C" int nerfb200_ray_batch(const long long* ray_index, int B, const float* c2w_raw,
                                  const float* c2w_noisy, const float* images,
                                  const int* image_id_map, int n_images, int H, int W, int n_sigmas,
                                  float focal, float pixel_width, int blur_low, int blur_high,
                                  float blur_coef, float* o_raw, float* o_noisy, float* d_raw,
                                  float* d_noisy, float* colors, long long* img_idx,
                                  float* pixel_width_out, void* stream) {
  NB_CHECK_ARG(B >= 0 && n_images >= 1 && H >= 1 && W >= 1 && n_sigmas >= 1,
               "ray_batch: bad shape B=%d N=%d H=%d W=%d n_sigmas=%d", B, n_images, H, W, n_sigmas);
  NB_CHECK_ARG(blur_low < n_sigmas && blur_high < n_sigmas && (blur_low < 0 || blur_high >= 0),
               "ray_batch: blur levels (%d, %d) out of range", blur_low, blur_high);
  if (B == 0) return NERFB200_OK;
  NB_CHECK_ARG(ray_index && c2w_raw && c2w_noisy && images && o_raw && o_noisy && d_raw && d_noisy &&
               colors && img_idx && pixel_width_out, "ray_batch: null pointer");
  RayBatchParams p{ray_index, B, c2w_raw, c2w_noisy, images, image_id_map, n_images, H, W, n_sigmas,
                   focal, pixel_width, blur_low, blur_high, blur_coef, o_raw, o_noisy, d_raw, d_noisy,
                   colors, img_idx, pixel_width_out};
  int blocks = ceil_div(B, 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  ray_batch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p);
  count_launch();
  NB_CHECK_LAUNCH();
  return NERFB200_OK;
}
