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

// Per-frame constants of the resample pass.  They are float64 expressions of the frame's statistics (numpy evaluates
// them in double, vidbase.py:78,141-149); thread 0 of every block evaluates them once and the block reads them from
// shared memory -- round 1 had every one of the S*S threads of a frame run the same eight double divisions.
struct ResampleFrame {
  double scale, alp, alpn, shift_x, shift_y;
  float color[3], a_other;
  int src, ox, oy, ok;
};

__global__ void __launch_bounds__(256) stage_resample_kernel(const uint8_t* __restrict__ frames,
                                                             const uint8_t* __restrict__ masks,
                                                             const float* __restrict__ flows,
                                                             const float* __restrict__ occs,
                                                             const FrameStats* __restrict__ st, float* __restrict__ img,
                                                             float* __restrict__ mask_out, float* __restrict__ flow_out,
                                                             float* __restrict__ occ_out, float* cam_focal, float* pps,
                                                             int32_t* boxes, int H, int W, int S) {
  __shared__ ResampleFrame s_c;
  const int slot = blockIdx.y, pair = blockIdx.z;
  const int f = pair * 2 + slot, fo = pair * 2 + (1 - slot);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0) {
    const FrameStats s_own = st[f], s_oth = st[fo];
    const Box b = crop_box(s_own), bo = crop_box(s_oth);
    ResampleFrame c;
    c.ok = !(s_own.count == 0 || s_oth.count == 0 || b.L <= 0 || bo.L <= 0);       // host raises (empty mask)
    c.src = 2 * b.L;
    c.alp = (double)c.src / (double)S;                                            // :141 (orisize[0] / maxw)
    c.alpn = (double)(2 * bo.L) / (double)S;
    c.scale = 1.0 / ((double)S / (double)c.src);     // cv2: inv_scale = dst/src; scale = 1/inv_scale
    c.ox = b.cx - b.L; c.oy = b.cy - b.L;            // crop origin in the frame
    c.shift_x = (double)(c.ox - (bo.cx - bo.L));
    c.shift_y = (double)(c.oy - (bo.cy - bo.L));
    c.a_other = (float)c.alpn;
    const unsigned long long sums[3] = {s_own.sum_r, s_own.sum_g, s_own.sum_b};
#pragma unroll
    for (int k = 0; k < 3; ++k)
      c.color[k] = c.ok ? (float)(1.0 - ((double)sums[k] / 255.0) / (double)s_own.count) : 0.f;   // :78
    s_c = c;
    if (i == 0) {
      boxes[4 * f] = b.cx; boxes[4 * f + 1] = b.cy; boxes[4 * f + 2] = b.L;
      boxes[4 * f + 3] = (int)min(s_own.count, (unsigned long long)INT_MAX);
      if (c.ok) {
        cam_focal[f] = (float)(1.0 / c.alp);                                       // :182
        pps[2 * f] = (float)c.ox; pps[2 * f + 1] = (float)c.oy;                   // :172
      }
    }
  }
  __syncthreads();
  if (!s_c.ok || i >= S * S) return;
  const int src = s_c.src, ox = s_c.ox, oy = s_c.oy;
  const double scale = s_c.scale;
  const int dy = i / S, dx = i - dy * S;
  int x0, x1, y0, y1;
  float wx, wy;
  linear_tap(dx, scale, src, x0, x1, wx);
  linear_tap(dy, scale, src, y0, y1, wy);
  const int X0 = x0 + ox, X1 = x1 + ox, Y0 = y0 + oy, Y1 = y1 + oy;
  const bool vx0 = X0 >= 0 && X0 < W, vx1 = X1 >= 0 && X1 < W, vy0 = Y0 >= 0 && Y0 < H, vy1 = Y1 >= 0 && Y1 < H;
  const long long HW = (long long)H * W;
  const uint8_t* mk = masks + f * HW;
  const uint8_t* fr = frames + f * HW * 3;
  const float* fl = flows + f * HW * 3;
  const float* oc = occs + f * HW;
  const long long SS = (long long)S * S;
  const bool v[4] = {vy0 && vx0, vy0 && vx1, vy1 && vx0, vy1 && vx1};
  // taps outside the frame read element 0 (always there) and are discarded: every load below is unconditional, so the
  // 4 mask + 12 colour + 4 occlusion + 12 flow reads of a pixel are in flight together instead of mask -> colour chains
  const long long q[4] = {v[0] ? (long long)Y0 * W + X0 : 0, v[1] ? (long long)Y0 * W + X1 : 0,
                          v[2] ? (long long)Y1 * W + X0 : 0, v[3] ? (long long)Y1 * W + X1 : 0};
  // mask: nearest neighbour, zero border
  const int nx = min((int)floor((double)dx * scale), src - 1) + ox, ny = min((int)floor((double)dy * scale), src - 1) + oy;
  const bool vn = nx >= 0 && nx < W && ny >= 0 && ny < H;
  const unsigned mnn = __ldg(mk + (vn ? (long long)ny * W + nx : 0));
  unsigned mq[4], cq[4][3];
  float oq[4], fq[4][3];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    mq[t] = __ldg(mk + q[t]);
    oq[t] = __ldg(oc + q[t]);
#pragma unroll
    for (int c = 0; c < 3; ++c) { cq[t][c] = __ldg(fr + 3 * q[t] + c); fq[t][c] = __ldg(fl + 3 * q[t] + c); }
  }
  // image: RGB / 255 where the mask is set, the complement of the masked mean colour elsewhere and outside
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float color = s_c.color[c];
    const int ch = 2 - c;                                                                 // BGR -> RGB (:56)
    float p[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) p[t] = (v[t] && mq[t] > 0) ? (float)cq[t][ch] / 255.f : color;
    img[((long long)f * 3 + c) * SS + i] = lerp2(p[0], p[1], p[2], p[3], wx, wy);
  }
  // occlusion map and flow: bilinear, zero border
  const float occ = lerp2(v[0] ? oq[0] : 0.f, v[1] ? oq[1] : 0.f, v[2] ? oq[2] : 0.f, v[3] ? oq[3] : 0.f, wx, wy);
  occ_out[(long long)f * SS + i] = occ;
  float fv[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    fv[c] = lerp2(v[0] ? fq[0][c] : 0.f, v[1] ? fq[1][c] : 0.f, v[2] ? fq[2][c] : 0.f, v[3] ? fq[3][c] : 0.f, wx, wy);
  // :141-149 -- the shift between the two crops and the scale change are added in float64, as numpy does
  const double dxs = s_c.shift_x + (double)dx * (s_c.alp - s_c.alpn);
  const double dys = s_c.shift_y + (double)dy * (s_c.alp - s_c.alpn);
  const float a_other = s_c.a_other;
  float u = (float)((double)fv[0] + dxs), w = (float)((double)fv[1] + dys);
  u = __fdiv_rn(u, a_other); w = __fdiv_rn(w, a_other);
  u = 2.f * __fdiv_rn(u, (float)S); w = 2.f * __fdiv_rn(w, (float)S);
  float* fo_ = flow_out + (long long)f * 3 * SS + i;
  fo_[0] = u; fo_[SS] = w;
  fo_[2 * SS] = (__fdiv_rn(fv[2], a_other) != 0.f && occ < 10.f) ? 1.f : 0.f;            // :151
  mask_out[(long long)f * SS + i] = (vn && mnn > 0) ? 1.f : 0.f;
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
