// staging.cu -- dataloader -> device staging: from decoded video frames to the tensors
// LASR's set_input() hands to the model, computed on the GPU instead of per item on a dataloader worker.
//
// Replaces the arithmetic of BaseDataset.__getitem__ (dataloader/vidbase.py:52-193) between file decode and
// collate: complement-colour background (:78-82), mask bounding box -> square crop (:96-106), integer-offset
// cv2.remap with constant border (:108-123), cv2.resize to img_size (:128-139; bilinear for image / flow /
// occlusion, nearest for the mask), flow re-expressed between the two crops and normalised (:141-160), the
// focal scale 1/alp (:175-183) and the principal-point offsets (:172-173).  File decode (cv2.imread, readPFM)
// stays on the host; the distance transforms and contour samples the reference also builds (:187-189) are not
// read by LASR.forward and are not staged.
//
// A "pair" is one dataset item: frame 0 and its partner frame n (each with the flow TO the other).  Layout:
//   in   frames_bgr [P,2,H,W,3] u8 (cv2 channel order), masks [P,2,H,W] u8, flows [P,2,H,W,3] f32
//        (u, v, valid), occs [P,2,H,W] f32
//   out  img [P,2,3,S,S] (RGB, 0..1), mask [P,2,S,S] (0/1), flow [P,2,3,S,S], occ [P,2,S,S],
//        cam_focal [P,2] (= 1/alp), pps [P,2,2], boxes [P,2,4] int32 (cx, cy, L, mask pixel count)
//
// Two kernels: (1) per-frame mask statistics -- bounding box, pixel count and the exact integer sums of the
// masked B, G, R bytes -- reduced per warp and merged with integer atomics (order independent, so deterministic);
// (2) one thread per output pixel: the crop is never materialised, each of the four bilinear taps is read from
// the source frame at (tap + crop offset) or replaced by the border value.  cv2.resize's arithmetic is
// followed step by step (scale = 1/(dst/src) in double, fx = float((dx+.5)*scale-.5), floor, clamp with zero
// weight, float32 weights, horizontal then vertical pass; nearest: min(floor(dx*scale), src-1)), and the
// float64 steps of the flow correction (:141-142) are done in double.  HBM-bound and small: a 480p frame pair
// is 16 MB in, 0.5 MB out at S = 256.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include <string>

#include "lasr_b200.h"
#include "host_common.h"

namespace lasr {

struct FrameStats {           // 48 bytes per frame
  int xmin, xmax, ymin, ymax;
  unsigned long long count, sum_b, sum_g, sum_r;
};

__global__ void stage_init_kernel(FrameStats* st, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st[i].xmin = INT_MAX; st[i].ymin = INT_MAX; st[i].xmax = -1; st[i].ymax = -1;
  st[i].count = 0; st[i].sum_b = 0; st[i].sum_g = 0; st[i].sum_r = 0;
}

// one CTA per strip of pixels of one frame; gridDim.y = frames.  A thread takes 16 consecutive pixels at a time: one
// 128-bit load of their mask bytes (a 480p frame is mostly background: all-zero words end the work there) and, where
// any is set, three 128-bit loads of their 48 BGR bytes -- the byte-per-thread form of round 1 moved 32 B per warp
// instruction.  Frames whose pixel count is not a multiple of 16 (the 16-byte alignment of a frame's first mask byte)
// take the scalar loop.
__global__ void __launch_bounds__(256) stage_stats_kernel(const uint8_t* __restrict__ frames,
                                                          const uint8_t* __restrict__ masks, FrameStats* st, int H,
                                                          int W, int px_per_block) {
  const int f = blockIdx.y;
  const long long HW = (long long)H * W;
  const uint8_t* mk = masks + f * HW;
  const uint8_t* fr = frames + f * HW * 3;
  int xmin = INT_MAX, xmax = -1, ymin = INT_MAX, ymax = -1;
  unsigned cnt = 0, sb = 0, sg = 0, sr = 0;
  const long long i0 = (long long)blockIdx.x * px_per_block;     // px_per_block is a multiple of 16
  const long long i1 = min(i0 + px_per_block, HW);
  auto take = [&](int x, int y, unsigned b, unsigned g, unsigned r) {
    xmin = min(xmin, x); xmax = max(xmax, x); ymin = min(ymin, y); ymax = max(ymax, y);
    ++cnt; sb += b; sg += g; sr += r;
  };
  if ((HW & 15) == 0) {
    for (long long i = i0 + 16 * (long long)threadIdx.x; i < i1; i += 16 * (long long)blockDim.x) {
      const uint4 m4 = __ldg(reinterpret_cast<const uint4*>(mk + i));
      if ((m4.x | m4.y | m4.z | m4.w) == 0u) continue;
      const unsigned mw[4] = {m4.x, m4.y, m4.z, m4.w};
      const uint4* c4 = reinterpret_cast<const uint4*>(fr + 3 * i);
      const uint4 ca = __ldg(c4), cb = __ldg(c4 + 1), cc = __ldg(c4 + 2);
      const unsigned cw[12] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w, cc.x, cc.y, cc.z, cc.w};
      int y = (int)(i / W), x = (int)(i - (long long)y * W);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if ((mw[k >> 2] >> (8 * (k & 3))) & 0xffu) {
          const int o = 3 * k;                                    // byte offset of this pixel's B, G, R
          take(x, y, (cw[o >> 2] >> (8 * (o & 3))) & 0xffu, (cw[(o + 1) >> 2] >> (8 * ((o + 1) & 3))) & 0xffu,
               (cw[(o + 2) >> 2] >> (8 * ((o + 2) & 3))) & 0xffu);
        }
        if (++x == W) { x = 0; ++y; }
      }
    }
  } else {
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      if (mk[i] > 0) {
        const int y = (int)(i / W), x = (int)(i - (long long)y * W);
        take(x, y, fr[3 * i], fr[3 * i + 1], fr[3 * i + 2]);
      }
    }
  }
  xmin = __reduce_min_sync(0xffffffffu, xmin); ymin = __reduce_min_sync(0xffffffffu, ymin);
  xmax = __reduce_max_sync(0xffffffffu, xmax); ymax = __reduce_max_sync(0xffffffffu, ymax);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  sb = __reduce_add_sync(0xffffffffu, sb); sg = __reduce_add_sync(0xffffffffu, sg); sr = __reduce_add_sync(0xffffffffu, sr);
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    FrameStats* s = st + f;
    atomicMin(&s->xmin, xmin); atomicMax(&s->xmax, xmax); atomicMin(&s->ymin, ymin); atomicMax(&s->ymax, ymax);
    atomicAdd(&s->count, (unsigned long long)cnt);
    atomicAdd(&s->sum_b, (unsigned long long)sb); atomicAdd(&s->sum_g, (unsigned long long)sg);
    atomicAdd(&s->sum_r, (unsigned long long)sr);
  }
}

struct Box { int cx, cy, L; };

__device__ __forceinline__ Box crop_box(const FrameStats& s) {     // vidbase.py:96-106
  Box b;
  b.cx = (s.xmax + s.xmin) / 2; b.cy = (s.ymax + s.ymin) / 2;      // non-negative: // == /
  const int lx = (s.xmax - s.xmin) / 2, ly = (s.ymax - s.ymin) / 2;
  b.L = (int)(1.2 * (double)max(lx, ly));
  return b;
}

// cv2.resize INTER_LINEAR source tap for destination index d: lower index and float weight of the upper tap
__device__ __forceinline__ void linear_tap(int d, double scale, int src, int& s0, int& s1, float& w) {
  float fx = (float)(((double)d + 0.5) * scale - 0.5);
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { sx = 0; fx = 0.f; }
  if (sx >= src - 1) { sx = src - 1; fx = 0.f; }
  s0 = sx; s1 = min(sx + 1, src - 1); w = fx;
}

__device__ __forceinline__ float lerp2(float p00, float p01, float p10, float p11, float wx, float wy) {
  const float top = p00 * (1.f - wx) + p01 * wx;      // horizontal pass
  const float bot = p10 * (1.f - wx) + p11 * wx;
  return top * (1.f - wy) + bot * wy;                 // vertical pass
}

__global__ void __launch_bounds__(256) stage_resample_kernel(const uint8_t* __restrict__ frames,
                                                             const uint8_t* __restrict__ masks,
                                                             const float* __restrict__ flows,
                                                             const float* __restrict__ occs,
                                                             const FrameStats* __restrict__ st, float* __restrict__ img,
                                                             float* __restrict__ mask_out, float* __restrict__ flow_out,
                                                             float* __restrict__ occ_out, float* cam_focal, float* pps,
                                                             int32_t* boxes, int H, int W, int S) {
  const int slot = blockIdx.y, pair = blockIdx.z;
  const int f = pair * 2 + slot, fo = pair * 2 + (1 - slot);
  const FrameStats s_own = st[f], s_oth = st[fo];
  const Box b = crop_box(s_own), bo = crop_box(s_oth);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    boxes[4 * f] = b.cx; boxes[4 * f + 1] = b.cy; boxes[4 * f + 2] = b.L;
    boxes[4 * f + 3] = (int)min(s_own.count, (unsigned long long)INT_MAX);
  }
  if (s_own.count == 0 || s_oth.count == 0 || b.L <= 0 || bo.L <= 0) return;   // host raises (empty mask)
  const int src = 2 * b.L;
  const double alp = (double)src / (double)S, alpn = (double)(2 * bo.L) / (double)S;   // :141 (orisize[0] / maxw)
  if (i == 0) {
    cam_focal[f] = (float)(1.0 / alp);                                                   // :182
    pps[2 * f] = (float)(b.cx - b.L); pps[2 * f + 1] = (float)(b.cy - b.L);             // :172
  }
  if (i >= S * S) return;
  const int dy = i / S, dx = i - dy * S;
  const double scale = 1.0 / ((double)S / (double)src);     // cv2: inv_scale = dst/src; scale = 1/inv_scale
  int x0, x1, y0, y1;
  float wx, wy;
  linear_tap(dx, scale, src, x0, x1, wx);
  linear_tap(dy, scale, src, y0, y1, wy);
  const int ox = b.cx - b.L, oy = b.cy - b.L;               // crop origin in the frame
  const int X0 = x0 + ox, X1 = x1 + ox, Y0 = y0 + oy, Y1 = y1 + oy;
  const bool vx0 = X0 >= 0 && X0 < W, vx1 = X1 >= 0 && X1 < W, vy0 = Y0 >= 0 && Y0 < H, vy1 = Y1 >= 0 && Y1 < H;
  const long long HW = (long long)H * W;
  const uint8_t* mk = masks + f * HW;
  const uint8_t* fr = frames + f * HW * 3;
  const float* fl = flows + f * HW * 3;
  const float* oc = occs + f * HW;
  const long long SS = (long long)S * S;
  const long long q00 = (long long)Y0 * W + X0, q01 = (long long)Y0 * W + X1, q10 = (long long)Y1 * W + X0,
                  q11 = (long long)Y1 * W + X1;
  const bool v00 = vy0 && vx0, v01 = vy0 && vx1, v10 = vy1 && vx0, v11 = vy1 && vx1;

  // image: RGB / 255 where the mask is set, the complement of the masked mean colour elsewhere and outside
  const bool m00 = v00 && mk[q00] > 0, m01 = v01 && mk[q01] > 0, m10 = v10 && mk[q10] > 0, m11 = v11 && mk[q11] > 0;
  const unsigned long long sums[3] = {s_own.sum_r, s_own.sum_g, s_own.sum_b};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float color = (float)(1.0 - ((double)sums[c] / 255.0) / (double)s_own.count);   // :78
    const int ch = 2 - c;                                                                 // BGR -> RGB (:56)
    const float p00 = m00 ? (float)fr[3 * q00 + ch] / 255.f : color;
    const float p01 = m01 ? (float)fr[3 * q01 + ch] / 255.f : color;
    const float p10 = m10 ? (float)fr[3 * q10 + ch] / 255.f : color;
    const float p11 = m11 ? (float)fr[3 * q11 + ch] / 255.f : color;
    img[((long long)f * 3 + c) * SS + i] = lerp2(p00, p01, p10, p11, wx, wy);
  }
  // occlusion map and flow: bilinear, zero border
  const float occ = lerp2(v00 ? oc[q00] : 0.f, v01 ? oc[q01] : 0.f, v10 ? oc[q10] : 0.f, v11 ? oc[q11] : 0.f, wx, wy);
  occ_out[(long long)f * SS + i] = occ;
  float fv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    fv[c] = lerp2(v00 ? fl[3 * q00 + c] : 0.f, v01 ? fl[3 * q01 + c] : 0.f, v10 ? fl[3 * q10 + c] : 0.f,
                  v11 ? fl[3 * q11 + c] : 0.f, wx, wy);
  // :141-149 -- the shift between the two crops and the scale change are added in float64, as numpy does
  const double dxs = (double)(ox - (bo.cx - bo.L)) + (double)dx * (alp - alpn);
  const double dys = (double)(oy - (bo.cy - bo.L)) + (double)dy * (alp - alpn);
  const float a_other = (float)alpn;
  float u = (float)((double)fv[0] + dxs), v = (float)((double)fv[1] + dys);
  u = __fdiv_rn(u, a_other); v = __fdiv_rn(v, a_other);
  u = 2.f * __fdiv_rn(u, (float)S); v = 2.f * __fdiv_rn(v, (float)S);
  float* fo_ = flow_out + (long long)f * 3 * SS + i;
  fo_[0] = u; fo_[SS] = v;
  fo_[2 * SS] = (__fdiv_rn(fv[2], a_other) != 0.f && occ < 10.f) ? 1.f : 0.f;            // :151
  // mask: nearest neighbour, zero border
  const int nx = min((int)floor((double)dx * scale), src - 1) + ox, ny = min((int)floor((double)dy * scale), src - 1) + oy;
  const bool vn = nx >= 0 && nx < W && ny >= 0 && ny < H;
  mask_out[(long long)f * SS + i] = (vn && mk[(long long)ny * W + nx] > 0) ? 1.f : 0.f;
}

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_stage_workspace_bytes(int num_pairs) {
  return num_pairs > 0 ? (size_t)num_pairs * 2 * sizeof(FrameStats) : 0;
}

extern "C" int lasr_b200_stage_pairs(const uint8_t* frames_bgr, const uint8_t* masks, const float* flows,
                                     const float* occs, float* img, float* mask_out, float* flow_out,
                                     float* occ_out, float* cam_focal, float* pps, int32_t* boxes,
                                     void* workspace, size_t workspace_bytes, int num_pairs, int height,
                                     int width, int image_size, void* stream) {
  if (!frames_bgr || !masks || !flows || !occs || !img || !mask_out || !flow_out || !occ_out || !cam_focal || !pps || !boxes)
    return lasr_fail("stage_pairs: null tensor pointer");
  if (num_pairs <= 0 || height <= 0 || width <= 0 || image_size <= 0)
    return lasr_fail("stage_pairs: num_pairs, height, width, image_size must be positive");
  if (num_pairs > 32767 || image_size > 16384) return lasr_fail("stage_pairs: num_pairs > 32767 or image_size > 16384");
  if (!workspace || workspace_bytes < lasr_b200_stage_workspace_bytes(num_pairs))
    return lasr_fail("stage_pairs: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  FrameStats* stats = reinterpret_cast<FrameStats*>(workspace);
  const int nf = 2 * num_pairs;
  stage_init_kernel<<<(nf + 127) / 128, 128, 0, st>>>(stats, nf);
  if (lasr_check_launch("stage_init_kernel")) return 1;
  const long long HW = (long long)height * width;
  const int px_per_block = 256 * 32;
  dim3 g1((unsigned)((HW + px_per_block - 1) / px_per_block), nf);
  stage_stats_kernel<<<g1, 256, 0, st>>>(frames_bgr, masks, stats, height, width, px_per_block);
  if (lasr_check_launch("stage_stats_kernel")) return 1;
  dim3 g2((unsigned)(((long long)image_size * image_size + 255) / 256), 2, num_pairs);
  stage_resample_kernel<<<g2, 256, 0, st>>>(frames_bgr, masks, flows, occs, stats, img, mask_out, flow_out, occ_out,
                                             cam_focal, pps, boxes, height, width, image_size);
  return lasr_check_launch("stage_resample_kernel");
}
