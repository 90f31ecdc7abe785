// texture_io.cu -- the two texture-atlas kernels of the SoftRas extension family.
//
//   load_textures         third_party/softras/soft_renderer/cuda/load_textures_cuda_kernel.cu:8-66
//                         (called from functional/load_obj.py:100): sample a texture image bilinearly at the
//                         R*R barycentric texel centres of every face -> 'surface' textures [F, R*R, 3]
//   create_texture_image  cuda/create_texture_image_cuda_kernel.cu:10-73 (functional/save_obj.py:27):
//                         the inverse for OBJ export: paint each face's texels into its cell of an atlas image
//
// Both are one thread per output element, HBM-bound and tiny (a 1024^2 atlas is 12 MB); nothing to tile.
// The reference instantiates them for float (and double); LASR / SoftRas only ever pass float32.  The float
// instantiation mixes precisions -- literals like `1. / 3.` and `1.` promote sub-expressions to double
// (load_textures :28-35, create_texture_image :54) -- and nvcc contracts the float sums of products; both are
// written out below so the texel weights, positions and the `<= 1` texel-half decision round as in the
// reference (association read off `cuobjdump -sass` of oracle/_ref/*_kernel.o).
//
// Deliberate deviations: the reference reads the image one row / column past the end when a texture coordinate
// is exactly 1.0 (weight 0, load_textures :50-53); those reads are clamped here.  CUDA launch errors are
// returned, not printed.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "lasr_b200.h"
#include "host_common.h"

namespace lasr {

__global__ void __launch_bounds__(256) load_textures_kernel(const float* __restrict__ image,      // [H,W,3]
                                                            const float* __restrict__ faces,      // [F,3,2] uv
                                                            const int32_t* __restrict__ is_update,  // [F]
                                                            float* __restrict__ textures,         // [F,R*R,3]
                                                            long long num_texels, int R, int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_texels) return;
  const int fn = (int)(i / (R * R));
  const int w_y = (int)((i % (R * R)) / R);
  const int w_x = (int)(i % R);
  if (is_update[fn] == 0) return;
  // barycentric coordinate of the texel centre, computed in double and rounded to float (:28-35)
  float w0, w1, w2;
  if (w_x + w_y < R) {
    w0 = (float)(((double)w_x + 1. / 3.) / (double)R);
    w1 = (float)(((double)w_y + 1. / 3.) / (double)R);
  } else {
    w0 = (float)((((double)R - 1. - (double)w_x) + 2. / 3.) / (double)R);
    w1 = (float)((((double)R - 1. - (double)w_y) + 2. / 3.) / (double)R);
  }
  w2 = (float)(1. - (double)w0 - (double)w1);
  const float* face = faces + (long long)fn * 6;
  // plain float expressions of the reference's shape (:40-43): nvcc contracts a*b + c*d + e*f the same way
  // for both builds (fma(e, f, fma(a, b, c*d))), then one multiply by (float)(size - 1)
  const float pos_x = (face[0] * w0 + face[2] * w1 + face[4] * w2) * (float)(W - 1);
  const float pos_y = (face[1] * w0 + face[3] * w1 + face[5] * w2) * (float)(H - 1);
  const int x0 = (int)pos_x, y0 = (int)pos_y;
  const int y1 = (int)(pos_y + 1.f);
  const int x1 = x0 + 1;
  const float wx1 = pos_x - (float)x0, wx0 = 1.f - wx1;
  const float wy1 = pos_y - (float)y0, wy0 = 1.f - wy1;
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  float* texture = textures + i * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float c = 0.f;                                                   // :48-54, same accumulation order
    c += image[((long long)cy0 * W + cx0) * 3 + k] * (wx0 * wy0);
    c += image[((long long)cy1 * W + cx0) * 3 + k] * (wx0 * wy1);
    c += image[((long long)cy0 * W + cx1) * 3 + k] * (wx1 * wy0);
    c += image[((long long)cy1 * W + cx1) * 3 + k] * (wx1 * wy1);
    texture[k] = c;
  }
}

__global__ void __launch_bounds__(256) create_texture_image_kernel(const float* __restrict__ faces,     // [F,3,2] px
                                                                   const float* __restrict__ textures,  // [F,R*R,3]
                                                                   float* __restrict__ image,           // [H,W,3]
                                                                   long long num_pixels, int num_faces, int R,
                                                                   int res_out, int tile_width, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_pixels) return;
  const int row_len = tile_width * res_out;
  const int x = (int)(i % row_len);
  const int y = (int)(i / row_len);
  const int row = x / res_out;
  const int column = y / res_out;
  const int fn = row + column * tile_width;
  if (fn >= num_faces) return;
  const float* texture = textures + (long long)fn * R * R * 3;
  const float* p0 = faces + (long long)fn * 6;
  const float* p1 = p0 + 2;
  const float* p2 = p0 + 4;
  // adjugate over determinant (:39-47), written as plain float expressions of the reference's shape so that
  // nvcc contracts them identically in both builds; IEEE division
  float inv[9] = {p1[1] - p2[1], p2[0] - p1[0], p1[0] * p2[1] - p2[0] * p1[1],
                  p2[1] - p0[1], p0[0] - p2[0], p2[0] * p0[1] - p0[0] * p2[1],
                  p0[1] - p1[1], p1[0] - p0[0], p0[0] * p1[1] - p1[0] * p0[1]};
  const float den = p2[0] * (p0[1] - p1[1]) + p0[0] * (p1[1] - p2[1]) + p1[0] * (p2[1] - p0[1]);
  const float den_eps = den + eps;
#pragma unroll
  for (int k = 0; k < 9; ++k) inv[k] = __fdiv_rn(inv[k], den_eps);
  // w = face_inv * (x, y, 1), clamped in double (the `1.` / `0.` literals, :54), renormalised (:52-58)
  float w[3], w_sum = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float v = inv[3 * k] * x + inv[3 * k + 1] * y + inv[3 * k + 2];
    w[k] = (float)fmax(fmin((double)v, 1.), 0.);
    w_sum += w[k];
  }
  const float ws = w_sum + eps;
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = __fdiv_rn(w[k], ws);
  const int w_x = w[0] * R;
  const int w_y = w[1] * R;
  int texel = ((w[0] + w[1]) * R - w_x - w_y <= 1) ? (w_y * R + w_x) : ((R - 1 - w_y) * R + (R - 1 - w_x));   // :62
  texel = min(max(texel, 0), R * R - 1);   // w[0] == 1 gives w_x == R in the reference (reads the next face's texel)
#pragma unroll
  for (int k = 0; k < 3; ++k) image[i * 3 + k] = texture[texel * 3 + k];
}

}  // namespace lasr

using namespace lasr;

extern "C" int lasr_b200_load_textures(const float* image, const float* faces, const int32_t* is_update,
                                       float* textures, int num_faces, int texture_res, int image_height,
                                       int image_width, void* stream) {
  if (!image || !faces || !is_update || !textures) return lasr_fail("load_textures: null tensor pointer");
  if (num_faces <= 0 || texture_res <= 0 || image_height <= 0 || image_width <= 0)
    return lasr_fail("load_textures: num_faces, texture_res and the image size must be positive");
  const long long n = (long long)num_faces * texture_res * texture_res;
  load_textures_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      image, faces, is_update, textures, n, texture_res, image_height, image_width);
  return lasr_check_launch("load_textures_kernel");
}

extern "C" int lasr_b200_create_texture_image(const float* faces, const float* textures, float* image,
                                              int num_faces, int texture_res_in, int image_height,
                                              int image_width, float eps, void* stream) {
  if (!faces || !textures || !image) return lasr_fail("create_texture_image: null tensor pointer");
  if (num_faces <= 0 || texture_res_in <= 0 || image_height <= 0 || image_width <= 0)
    return lasr_fail("create_texture_image: num_faces, texture_res and the image size must be positive");
  const int tile_width = (int)sqrt((double)(num_faces - 1)) + 1;    // create_texture_image_cuda_kernel.cu:84
  const int res_out = image_width / tile_width;                     // :85
  if (res_out <= 0) return lasr_fail("create_texture_image: image narrower than one pixel per face");
  const long long n = (long long)image_height * image_width;
  create_texture_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      faces, textures, image, n, num_faces, texture_res_in, res_out, tile_width, eps);
  return lasr_check_launch("create_texture_image_kernel");
}
