// losses.cu -- fused masked reductions of LASR's three image losses, sm_100a.
//
// The reference computes them inline in LASR.forward (nnutils/mesh_net.py:372-447) with Python
// double loops over (frame, hypothesis): boolean-mask indexing + .mean() per pair and a host
// synchronisation per pair (`if mask[i,j].sum()==0`, :412) -- 2bs*H tiny kernels and syncs per loss.
// Here ONE pass over the images produces, per (frame i, hypothesis j), the six masked sums all three
// losses are made of; a one-block kernel turns sums into the per-pair losses and their means
// (losses_finalize_kernel) and its adjoint turns upstream gradients into gradients of the sums
// (losses_coef_kernel); the backward kernel then produces per-pixel gradients in one elementwise pass.
// (lasr_b200/losses.py also offers the three losses separately with the scalar step in PyTorch.)
//
//   v  = occ != 0                      (pixel inside the original frame)
//   k  = !bgmask & v & (mask > 0)      (flow visibility)
//   s  = sigmoid(-occ)
//   obs = img * (mask > 0),  white = 1 - (mask > 0) + obs
//
//   sums[i,j,0] = sum v (mask_pred - mask)^2                    -> silhouette loss (:374-390)
//   sums[i,j,1] = sum k s |flow_rd - flow_obs|_2                -> flow loss       (:393-416)
//   sums[i,j,2] = sum k s        sums[i,j,3] = sum k
//   sums[i,j,4] = sum v mean_c |obs_c   - tex_c * fg|           -> texture L1      (:420-441)
//   sums[i,j,5] = sum v mean_c |white_c - tex_c|
//   vcount[i]   = sum v
//
// Deterministic: fixed-order block reductions into per-block partials, summed in block order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "host_common.h"
#include "lasr_b200.h"

namespace lasr {

constexpr int kLossThreads = 256;
constexpr int kLossPixPerBlock = 4096;   // pixels handled by one block (16 per thread)
constexpr int kNumSums = 7;              // 6 sums + the valid-pixel count

struct LossArgs {
  const float* mask_pred;   // [N,H,S,S]
  const float* flow_rd;     // [N,H,S,S,2]
  const uint8_t* bgmask;    // [N,H,S,S]
  const float* tex;         // [N,H,3,S,S]
  const float* fg;          // [N,H,S,S]
  const float* imgs;        // [N,3,S,S]
  const float* masks;       // [N,S,S]
  const float* occ;         // [N,S,S]
  const float* flow_obs;    // [N,3,S,S] (channels 0,1 used)
  int N, H, SS;             // SS = S*S
  int has_flow, has_tex;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// V consecutive pixels per thread and iteration: V = 4 turns every plane access into one 128-bit load (13 of them in
// flight per thread before the first use) when S*S is a multiple of 4 -- the kernel is HBM-bound and the scalar form
// (V = 1, kept for odd image sizes) reached 31 % of the copy bandwidth; all loads are unconditional so that they can be
// issued together (the masked-out values are simply not accumulated).
template <int V> struct PixVec;
template <> struct PixVec<1> {
  static __device__ __forceinline__ void ld(const float* p, float* o) { o[0] = __ldg(p); }
  static __device__ __forceinline__ void ldb(const uint8_t* p, unsigned* o) { o[0] = __ldg(p); }
  static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; }
};
template <> struct PixVec<4> {
  static __device__ __forceinline__ void ld(const float* p, float* o) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  }
  static __device__ __forceinline__ void ldb(const uint8_t* p, unsigned* o) {
    const unsigned q = __ldg(reinterpret_cast<const unsigned*>(p));
    o[0] = q & 0xffu; o[1] = (q >> 8) & 0xffu; o[2] = (q >> 16) & 0xffu; o[3] = q >> 24;
  }
  static __device__ __forceinline__ void st(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <int V>
__global__ void __launch_bounds__(kLossThreads) losses_forward_kernel(LossArgs A, float* partial, int nblk) {
  __shared__ float s_w[kLossThreads / 32][kNumSums];
  const int ij = blockIdx.y;            // frame * H + hypothesis
  const int i = ij / A.H;
  const long long SS = A.SS;
  float acc[kNumSums];
#pragma unroll
  for (int q = 0; q < kNumSums; ++q) acc[q] = 0.f;
  const int p0 = blockIdx.x * kLossPixPerBlock;
  const int p1 = min(p0 + kLossPixPerBlock, A.SS);
  for (int p = p0 + threadIdx.x * V; p < p1; p += kLossThreads * V) {
    float occ[V], m[V], mp[V], fx[V], fy[V], ox[V], oy[V], fgv[V], img[3][V], tx[3][V];
    unsigned bgm[V];
    PixVec<V>::ld(A.occ + i * SS + p, occ);
    PixVec<V>::ld(A.masks + i * SS + p, m);
    PixVec<V>::ld(A.mask_pred + ij * SS + p, mp);
    if (A.has_flow) {
      PixVec<V>::ldb(A.bgmask + ij * SS + p, bgm);
      float f2[2 * V];                                    // interleaved (x, y) per pixel
      PixVec<V>::ld(A.flow_rd + 2 * (ij * SS + p), f2);
      if constexpr (V == 4) PixVec<V>::ld(A.flow_rd + 2 * (ij * SS + p) + 4, f2 + 4); else f2[1] = __ldg(A.flow_rd + 2 * (ij * SS + p) + 1);
#pragma unroll
      for (int k = 0; k < V; ++k) { fx[k] = f2[2 * k]; fy[k] = f2[2 * k + 1]; }
      PixVec<V>::ld(A.flow_obs + (i * 3 + 0) * SS + p, ox);
      PixVec<V>::ld(A.flow_obs + (i * 3 + 1) * SS + p, oy);
    }
    if (A.has_tex) {
      PixVec<V>::ld(A.fg + ij * SS + p, fgv);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        PixVec<V>::ld(A.imgs + (i * 3 + c) * SS + p, img[c]);
        PixVec<V>::ld(A.tex + ((long long)ij * 3 + c) * SS + p, tx[c]);
      }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const bool v = occ[k] != 0.f;
      const bool fgm = m[k] > 0.f;
      if (v) {
        acc[6] += 1.f;
        const float d = mp[k] - m[k];
        acc[0] += d * d;
      }
      if (A.has_flow && v && fgm && bgm[k] == 0) {
        const float dx = fx[k] - ox[k];
        const float dy = fy[k] - oy[k];
        const float sg = 1.f / (1.f + expf(occ[k]));
        acc[1] += sg * sqrtf(dx * dx + dy * dy);
        acc[2] += sg;
        acc[3] += 1.f;
      }
      if (A.has_tex && v) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float obs = fgm ? img[c][k] : 0.f;
          const float white = fgm ? img[c][k] : 1.f;
          a += fabsf(obs - tx[c][k] * fgv[k]);
          b += fabsf(white - tx[c][k]);
        }
        acc[4] += a * (1.f / 3.f);
        acc[5] += b * (1.f / 3.f);
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < kNumSums; ++q) {
    const float w = wsum(acc[q]);
    if (lane == 0) s_w[warp][q] = w;
  }
  __syncthreads();
  if (threadIdx.x < kNumSums) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) t += s_w[w][threadIdx.x];
    partial[((long long)ij * nblk + blockIdx.x) * kNumSums + threadIdx.x] = t;
  }
}

// Sum of the nblk block partials of output t = (frame*H + hypothesis, q), by one warp: lane l adds blocks l, l + 32, ...
// in order, then a butterfly -- a fixed order (deterministic), and one or two rounds of memory latency instead of nblk.
__device__ __forceinline__ float reduce_partials_warp(const float* __restrict__ partial, int nblk, int t, int lane) {
  const int ij = t / kNumSums, q = t % kNumSums;
  const float* src = partial + (long long)ij * nblk * kNumSums + q;
  float acc = 0.f;
#pragma unroll 4
  for (int b = lane; b < nblk; b += 32) acc += __ldg(src + (long long)b * kNumSums);
  return wsum(acc);
}

__global__ void __launch_bounds__(256) losses_reduce_kernel(const float* __restrict__ partial, int nblk, int NH,
                                                            float* sums) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= NH * kNumSums) return;
  const float acc = reduce_partials_warp(partial, nblk, t, lane);
  if (lane == 0) sums[t] = acc;
}

// gsums [N,H,7] (gradient of the scalar objective w.r.t. each sum; entry 2,3,6 ignored: those sums
// do not depend on differentiable inputs) -> per-pixel gradients.
template <int V>
__global__ void __launch_bounds__(kLossThreads) losses_backward_kernel(LossArgs A, const float* gsums,
                                                                       float* g_mask_pred, float* g_flow_rd,
                                                                       float* g_tex, float* g_fg) {
  const int ij = blockIdx.y;
  const int i = ij / A.H;
  const long long SS = A.SS;
  const int p = (blockIdx.x * kLossThreads + threadIdx.x) * V;
  if (p >= A.SS) return;
  const float g0 = gsums[ij * kNumSums + 0], g1 = gsums[ij * kNumSums + 1];
  const float g4 = gsums[ij * kNumSums + 4], g5 = gsums[ij * kNumSums + 5];
  float occ[V], m[V], mp[V], fx[V], fy[V], ox[V], oy[V], fgv[V], img[3][V], tx[3][V];
  unsigned bgm[V];
  PixVec<V>::ld(A.occ + i * SS + p, occ);
  PixVec<V>::ld(A.masks + i * SS + p, m);
  if (g_mask_pred) PixVec<V>::ld(A.mask_pred + ij * SS + p, mp);
  const bool do_flow = A.has_flow && g_flow_rd, do_tex = A.has_tex && g_tex;
  if (do_flow) {
    PixVec<V>::ldb(A.bgmask + ij * SS + p, bgm);
    float f2[2 * V];
    PixVec<V>::ld(A.flow_rd + 2 * (ij * SS + p), f2);
    if constexpr (V == 4) PixVec<V>::ld(A.flow_rd + 2 * (ij * SS + p) + 4, f2 + 4); else f2[1] = __ldg(A.flow_rd + 2 * (ij * SS + p) + 1);
#pragma unroll
    for (int k = 0; k < V; ++k) { fx[k] = f2[2 * k]; fy[k] = f2[2 * k + 1]; }
    PixVec<V>::ld(A.flow_obs + (i * 3 + 0) * SS + p, ox);
    PixVec<V>::ld(A.flow_obs + (i * 3 + 1) * SS + p, oy);
  }
  if (do_tex) {
    PixVec<V>::ld(A.fg + ij * SS + p, fgv);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      PixVec<V>::ld(A.imgs + (i * 3 + c) * SS + p, img[c]);
      PixVec<V>::ld(A.tex + ((long long)ij * 3 + c) * SS + p, tx[c]);
    }
  }
  float o_mp[V], o_flow[2 * V], o_tex[3][V], o_fg[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const bool v = occ[k] != 0.f;
    const bool fgm = m[k] > 0.f;
    o_mp[k] = (g_mask_pred && v) ? 2.f * (mp[k] - m[k]) * g0 : 0.f;
    o_flow[2 * k] = 0.f; o_flow[2 * k + 1] = 0.f;
    if (do_flow && v && fgm && bgm[k] == 0) {
      const float dx = fx[k] - ox[k];
      const float dy = fy[k] - oy[k];
      const float nrm = sqrtf(dx * dx + dy * dy);
      if (nrm > 0.f) {
        const float c = g1 / (1.f + expf(occ[k])) / nrm;
        o_flow[2 * k] = c * dx; o_flow[2 * k + 1] = c * dy;
      }
    }
    float gfg = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gt = 0.f;
      if (do_tex && v) {
        const float obs = fgm ? img[c][k] : 0.f;
        const float white = fgm ? img[c][k] : 1.f;
        const float t = tx[c][k];
        const float ra = obs - t * fgv[k], rb = white - t;
        const float sa = ra > 0.f ? 1.f : (ra < 0.f ? -1.f : 0.f);   // d|x|/dx, 0 at 0 like torch.abs
        const float sb = rb > 0.f ? 1.f : (rb < 0.f ? -1.f : 0.f);
        gt = (-sa * fgv[k] * g4 - sb * g5) * (1.f / 3.f);
        gfg += -sa * t * g4 * (1.f / 3.f);
      }
      o_tex[c][k] = gt;
    }
    o_fg[k] = gfg;
  }
  if (g_mask_pred) PixVec<V>::st(g_mask_pred + ij * SS + p, o_mp);
  if (do_flow) {
    PixVec<V>::st(g_flow_rd + 2 * (ij * SS + p), o_flow);
    if constexpr (V == 4) PixVec<V>::st(g_flow_rd + 2 * (ij * SS + p) + 4, o_flow + 4); else g_flow_rd[2 * (ij * SS + p) + 1] = o_flow[1];
  }
  if (do_tex) {
#pragma unroll
    for (int c = 0; c < 3; ++c) PixVec<V>::st(g_tex + ((long long)ij * 3 + c) * SS + p, o_tex[c]);
    if (g_fg) PixVec<V>::st(g_fg + ij * SS + p, o_fg);
  }
}

// sums [N,H,7] -> per-(frame,hypothesis) losses sub[3][N*H] (mask, flow, texture) and their means
// out3[3], following nnutils/mesh_net.py:388-390, 404-416, 440-445.  One block; fixed-order reduction.
__device__ __forceinline__ void losses_finalize_block(const float* sums, int N, int H, float l1tex_wt, float* sub,
                                                      float* out3, bool with_total) {   // the first 256 threads work
  __shared__ float s_red[3][256 / 32];
  __shared__ float s_out[3];
  const int NH = N * H;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int t = threadIdx.x; t < NH && threadIdx.x < 256; t += 256) {
    const int i = t / H;
    const float* q = sums + (long long)t * kNumSums;
    const float V = q[6];
    float Bs = 0.f, Cs = 0.f;
    for (int j = 0; j < H; ++j) { Bs += sums[((long long)i * H + j) * kNumSums + 2]; Cs += sums[((long long)i * H + j) * kNumSums + 3]; }
    const float m = 0.5f * q[0] / V;
    const float wmean = Bs / Cs;
    const float f = q[3] == 0.f ? 0.f : 0.5f * q[1] / wmean / q[3];
    const float x = 0.25f * ((q[4] + q[5]) / V * 2.f * l1tex_wt);
    sub[t] = m; sub[NH + t] = f; sub[2 * NH + t] = x;
    acc[0] += m; acc[1] += f; acc[2] += x;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float w = wsum(acc[k]);
    if (lane == 0 && warp < 256 / 32) s_red[k][warp] = w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < 256 / 32; ++w) t += s_red[threadIdx.x][w];
    out3[threadIdx.x] = s_out[threadIdx.x] = t / (float)NH;
  }
  if (with_total) {        // out3 has a fourth entry: (mask + flow) + texture, the sum LASR forms of them (mesh_net.py:448)
    __syncthreads();
    if (threadIdx.x == 0) out3[3] = __fadd_rn(__fadd_rn(s_out[0], s_out[1]), s_out[2]);
  }
}

__global__ void __launch_bounds__(256) losses_finalize_kernel(const float* sums, int N, int H, float l1tex_wt,
                                                              float* sub, float* out3) {
  losses_finalize_block(sums, N, H, l1tex_wt, sub, out3, false);
}

// Block partials -> sums [N,H,7] -> per-pair losses and the three means, in ONE block (the two steps used to be two
// launches behind the pixel pass; at LASR's batch sizes the whole head is launch-latency bound).  Same summation orders as
// losses_reduce_kernel and losses_finalize_kernel: same bits.
__global__ void __launch_bounds__(1024) losses_reduce_finalize_kernel(const float* __restrict__ partial, int nblk, int N,
                                                                     int H, float l1tex_wt, float* sums, float* sub,
                                                                     float* out3) {
  const int tot = N * H * kNumSums;
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x >> 5; t < tot; t += blockDim.x >> 5) {
    const float acc = reduce_partials_warp(partial, nblk, t, lane);
    if (lane == 0) sums[t] = acc;
  }
  __syncthreads();                       // this block's own global writes are visible to it after the barrier
  losses_finalize_block(sums, N, H, l1tex_wt, sub, out3, true);
}

// upstream gradients of the three means (g3[3]) and of the per-pair losses (gsub[3][N*H], may be null)
// -> gradient of every sum (only sums 0, 1, 4, 5 carry gradient to the images)
__global__ void losses_coef_kernel(const float* sums, const float* g3, const float* gsub, int N, int H,
                                   float l1tex_wt, float* gsums, int with_total) {
  const int NH = N * H;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= NH) return;
  const int i = t / H;
  const float* q = sums + (long long)t * kNumSums;
  const float V = q[6];
  float Bs = 0.f, Cs = 0.f;
  for (int j = 0; j < H; ++j) { Bs += sums[((long long)i * H + j) * kNumSums + 2]; Cs += sums[((long long)i * H + j) * kNumSums + 3]; }
  const float inv = 1.f / (float)NH;
  const float gt = with_total ? g3[3] : 0.f;          // gradient of the fourth entry, the sum of the three
  const float gm = (g3[0] + gt) * inv + (gsub ? gsub[t] : 0.f);
  const float gf = (g3[1] + gt) * inv + (gsub ? gsub[NH + t] : 0.f);
  const float gx = (g3[2] + gt) * inv + (gsub ? gsub[2 * NH + t] : 0.f);
  float* o = gsums + (long long)t * kNumSums;
  o[0] = gm * 0.5f / V;
  o[1] = q[3] == 0.f ? 0.f : gf * 0.5f / (Bs / Cs) / q[3];
  o[2] = 0.f; o[3] = 0.f;
  o[4] = gx * 0.5f * l1tex_wt / V;
  o[5] = o[4];
  o[6] = 0.f;
}

// ---------------------------------------------------------------------------------------------
// Flow epilogue of render_flow_soft_2 (nnutils/mesh_net.py:88-104): the rasterizer rendered the 3-D
// positions of frame 0 and frame 1 as colours (planar [2n,4,S,S]: first n = frame-0 positions, last
// n = frame-1 positions); re-project both with their intrinsics and subtract.
//   bg   = pos0.z < 1e-9 | pos1.z < 1e-9           (both set to 10 there, as the reference does)
//   flow = (pp1 + pos1.xy * f1 / pos1.z) - (pp0 + pos0.xy * f0 / pos0.z)      [n,S,S,2]
// Gradient flows to frame-1 positions only, and not at background pixels (`p0.detach()`, :102-103).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flow_epilogue_forward_kernel(const float* __restrict__ px, const float* pp0,
                                                                    const float* pp1, const float* f0, const float* f1,
                                                                    float* flow, uint8_t* bgmask, int n, int SS) {
  const int i = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= SS) return;
  const float* a = px + (long long)i * 4 * SS + p;         // frame-0 positions
  const float* b = px + (long long)(n + i) * 4 * SS + p;   // frame-1 positions
  float x0 = a[0], y0 = a[SS], z0 = a[2 * (long long)SS];
  float x1 = b[0], y1 = b[SS], z1 = b[2 * (long long)SS];
  const bool bg = z0 < 1e-9f || z1 < 1e-9f;
  if (bg) { x0 = y0 = z0 = 10.f; x1 = y1 = z1 = 10.f; }
  const float u0 = pp0[2 * i] + x0 * f0[i] / z0, v0 = pp0[2 * i + 1] + y0 * f0[i] / z0;
  const float u1 = pp1[2 * i] + x1 * f1[i] / z1, v1 = pp1[2 * i + 1] + y1 * f1[i] / z1;
  reinterpret_cast<float2*>(flow)[(long long)i * SS + p] = make_float2(u1 - u0, v1 - v0);
  bgmask[(long long)i * SS + p] = bg ? 1 : 0;
}

// grad_flow [n,S,S,2] -> grad_px [2n,4,S,S] (zero except xyz of the frame-1 half at foreground pixels),
// plus per-frame partial sums for pp1 (x,y) and f1: partial [n][nblk][3]
__global__ void __launch_bounds__(256) flow_epilogue_backward_kernel(const float* __restrict__ px, const float* f1,
                                                                     const float* __restrict__ gflow, float* gpx,
                                                                     float* partial, int n, int SS, int nblk) {
  __shared__ float s_w[8][3];
  const int i = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  if (p < SS) {
    const float* a = px + (long long)i * 4 * SS + p;
    const float* b = px + (long long)(n + i) * 4 * SS + p;
    const float z0 = a[2 * (long long)SS];
    const float x1 = b[0], y1 = b[SS], z1 = b[2 * (long long)SS];
    const bool bg = z0 < 1e-9f || z1 < 1e-9f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (!bg) {
      const float2 g = reinterpret_cast<const float2*>(gflow)[(long long)i * SS + p];
      const float f = f1[i], iz = 1.f / z1;
      gx = g.x * f * iz;
      gy = g.y * f * iz;
      gz = -(g.x * x1 + g.y * y1) * f * iz * iz;
      acc[0] = g.x; acc[1] = g.y; acc[2] = (g.x * x1 + g.y * y1) * iz;
    }
    float* o0 = gpx + (long long)i * 4 * SS + p;
    float* o1 = gpx + (long long)(n + i) * 4 * SS + p;
#pragma unroll
    for (int c = 0; c < 4; ++c) o0[c * (long long)SS] = 0.f;
    o1[0] = gx; o1[SS] = gy; o1[2 * (long long)SS] = gz; o1[3 * (long long)SS] = 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float w = wsum(acc[k]);
    if (lane == 0) s_w[warp][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w][threadIdx.x];
    partial[((long long)i * nblk + blockIdx.x) * 3 + threadIdx.x] = t;
  }
}

// One warp per output (frame, component): lane l adds blocks l, l + 32, ... in order, then a butterfly over the lanes --
// a fixed order, so the result is deterministic.  (One thread per output walked the nblk partials one dependent load
// after another: 12 us for 24 outputs of 256 blocks, as long as the head's backward pass itself at LASR's batch size.)
__global__ void __launch_bounds__(128) flow_epilogue_reduce_kernel(const float* __restrict__ partial, int n, int nblk,
                                                                   float* gpp1, float* gf1, int shift = 0) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n * 3) return;
  const int i = t / 3, k = t % 3;
  const float* src = partial + (long long)i * nblk * 3 + k;
  float acc = 0.f;
#pragma unroll 4
  for (int b = lane; b < nblk; b += 32) acc += __ldg(src + (long long)b * 3);
  acc = wsum(acc);
  if (lane == 0) {
    const int o = i + shift >= n ? i + shift - n : i + shift;     // the head's pair_shift: the entry frame i read
    if (k < 2) gpp1[2 * o + k] = acc; else gf1[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Fused image head for the three-set render (lasr_b200_render_forward with num_sets = 3): the image
// [N*H,10,S,S] holds, per frame, the camera-space position of the frame's own pose (planes 0-2), of
// its pair's pose (3-5), the texture (6-8) and the silhouette (9).  One pass does the work of the flow
// epilogue above for BOTH flow directions (frame i of the first half pairs with i+bs and vice versa;
// the caller passes per-frame own / pair intrinsics) and of losses_forward_kernel, without the
// flow / bgmask / texture / mask intermediates in HBM; its adjoint writes the whole image gradient.
// Arithmetic is that of flow_epilogue_* and losses_* above, expression for expression.
// ---------------------------------------------------------------------------------------------
constexpr int kHeadPlanes = 10;
constexpr int kHeadPixPerBlock = 1024;   // 4 pixels per thread: enough blocks in flight to cover the 17 loads per pixel

struct HeadArgs {
  const float* image;      // [N*H,10,S,S]
  const float* pp_own;     // [N*H,2]
  const float* pp_pair;    // [N*H,2]
  const float* f_own;      // [N*H]
  const float* f_pair;     // [N*H]
  const float* imgs;       // [N,3,S,S]
  const float* masks;      // [N,S,S]
  const float* occ;        // [N,S,S]
  const float* flow_obs;   // [N,3,S,S]
  int N, H, SS;
  int pair_shift;          // frame ij reads (and its gradient is written to) entry (ij + pair_shift) mod N*H of the PAIR arrays
};

__device__ __forceinline__ int head_pair_index(const HeadArgs& A, int ij) {
  const int j = ij + A.pair_shift;
  return j >= A.N * A.H ? j - A.N * A.H : j;
}

struct HeadFlow { float fx, fy, x1, y1, z1; bool bg; };

__device__ __forceinline__ HeadFlow head_flow(const HeadArgs& A, const float* img, int ij, int p) {
  const long long SS = A.SS;
  float x0 = img[p], y0 = img[SS + p], z0 = img[2 * SS + p];
  float x1 = img[3 * SS + p], y1 = img[4 * SS + p], z1 = img[5 * SS + p];
  HeadFlow r;
  r.bg = z0 < 1e-9f || z1 < 1e-9f;
  if (r.bg) { x0 = y0 = z0 = 10.f; x1 = y1 = z1 = 10.f; }
  const float f0 = A.f_own[ij], f1 = A.f_pair[head_pair_index(A, ij)];
  const float u0 = A.pp_own[2 * ij] + x0 * f0 / z0, v0 = A.pp_own[2 * ij + 1] + y0 * f0 / z0;
  const float u1 = A.pp_pair[2 * head_pair_index(A, ij)] + x1 * f1 / z1, v1 = A.pp_pair[2 * head_pair_index(A, ij) + 1] + y1 * f1 / z1;
  r.fx = u1 - u0; r.fy = v1 - v0; r.x1 = x1; r.y1 = y1; r.z1 = z1;
  return r;
}

// per-pixel body of the forward head: planes x[0..9], observations, accumulators
__device__ __forceinline__ void head_pixel(const HeadArgs& A, int ij, const float* x, float occ, float m,
                                           const float* im3, float fobs_x, float fobs_y, float* acc, bool want_flow,
                                           float2* flow, uint8_t* bgo) {
  const bool v = occ != 0.f;
  const bool fgm = m > 0.f;
  const float alpha = x[9];
  if (v) {
    acc[6] += 1.f;
    const float d = alpha - m;
    acc[0] += d * d;
  }
  if ((v && fgm) || want_flow) {
    float x0 = x[0], y0 = x[1], z0 = x[2], x1 = x[3], y1 = x[4], z1 = x[5];
    const bool bg = z0 < 1e-9f || z1 < 1e-9f;
    if (bg) { x0 = y0 = z0 = 10.f; x1 = y1 = z1 = 10.f; }
    const float f0 = A.f_own[ij], f1 = A.f_pair[head_pair_index(A, ij)];
    const float u0 = A.pp_own[2 * ij] + x0 * f0 / z0, v0 = A.pp_own[2 * ij + 1] + y0 * f0 / z0;
    const float u1 = A.pp_pair[2 * head_pair_index(A, ij)] + x1 * f1 / z1, v1 = A.pp_pair[2 * head_pair_index(A, ij) + 1] + y1 * f1 / z1;
    const float fx = u1 - u0, fy = v1 - v0;
    if (want_flow) { *flow = make_float2(fx, fy); *bgo = bg ? 1 : 0; }
    if (v && fgm && !bg) {
      const float dx = fx - fobs_x, dy = fy - fobs_y;
      const float s = 1.f / (1.f + expf(occ));
      acc[1] += s * sqrtf(dx * dx + dy * dy);
      acc[2] += s;
      acc[3] += 1.f;
    }
  }
  if (v) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float obs = fgm ? im3[c] : 0.f;
      const float white = fgm ? im3[c] : 1.f;
      const float t = x[6 + c];
      a += fabsf(obs - t * alpha);
      b += fabsf(white - t);
    }
    acc[4] += a * (1.f / 3.f);
    acc[5] += b * (1.f / 3.f);
  }
}

// kVec = 4: each thread takes 4 consecutive pixels with 128-bit loads (needs S*S % 4 == 0): 17 independent
// float4 loads in flight per thread instead of 17 dependent-latency scalar rounds
template <int kVec>
__global__ void __launch_bounds__(kLossThreads) head_forward_kernel(HeadArgs A, float* partial, int nblk,
                                                                    float* flow_out, uint8_t* bg_out) {
  __shared__ float s_w[kLossThreads / 32][kNumSums];
  const int ij = blockIdx.y;
  const int i = ij / A.H;
  const long long SS = A.SS;
  const float* img = A.image + (long long)ij * kHeadPlanes * SS;
  float acc[kNumSums];
#pragma unroll
  for (int q = 0; q < kNumSums; ++q) acc[q] = 0.f;
  const int p0 = blockIdx.x * kHeadPixPerBlock;
  const int p1 = min(p0 + kHeadPixPerBlock, A.SS);
  for (int p = p0 + threadIdx.x * kVec; p < p1; p += kLossThreads * kVec) {
    float x[kHeadPlanes][kVec], occ[kVec], m[kVec], im[3][kVec], fo[2][kVec];
    if (kVec == 4) {
#pragma unroll
      for (int c = 0; c < kHeadPlanes; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(img + c * SS + p);
        x[c][0] = t.x; x[c][1] = t.y; x[c][2] = t.z; x[c][kVec - 1] = t.w;
      }
      auto ld4 = [&](const float* base, float* o) {
        const float4 t = *reinterpret_cast<const float4*>(base + p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[kVec - 1] = t.w;
      };
      ld4(A.occ + i * SS, occ); ld4(A.masks + i * SS, m);
#pragma unroll
      for (int c = 0; c < 3; ++c) ld4(A.imgs + (i * 3 + c) * SS, im[c]);
      ld4(A.flow_obs + (i * 3 + 0) * SS, fo[0]); ld4(A.flow_obs + (i * 3 + 1) * SS, fo[1]);
    } else {
#pragma unroll
      for (int c = 0; c < kHeadPlanes; ++c) x[c][0] = img[c * SS + p];
      occ[0] = A.occ[i * SS + p]; m[0] = A.masks[i * SS + p];
#pragma unroll
      for (int c = 0; c < 3; ++c) im[c][0] = A.imgs[(i * 3 + c) * SS + p];
      fo[0][0] = A.flow_obs[(i * 3 + 0) * SS + p]; fo[1][0] = A.flow_obs[(i * 3 + 1) * SS + p];
    }
#pragma unroll
    for (int e = 0; e < kVec; ++e) {
      float xe[kHeadPlanes], im3[3];
#pragma unroll
      for (int c = 0; c < kHeadPlanes; ++c) xe[c] = x[c][e];
#pragma unroll
      for (int c = 0; c < 3; ++c) im3[c] = im[c][e];
      head_pixel(A, ij, xe, occ[e], m[e], im3, fo[0][e], fo[1][e], acc, flow_out != nullptr,
                 flow_out ? reinterpret_cast<float2*>(flow_out) + ij * SS + p + e : nullptr,
                 bg_out ? bg_out + ij * SS + p + e : nullptr);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < kNumSums; ++q) {
    const float w = wsum(acc[q]);
    if (lane == 0) s_w[warp][q] = w;
  }
  __syncthreads();
  if (threadIdx.x < kNumSums) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) t += s_w[w][threadIdx.x];
    partial[((long long)ij * nblk + blockIdx.x) * kNumSums + threadIdx.x] = t;
  }
}

// gsums [N,H,7] -> grad_image [N*H,10,S,S] (every plane written) + per-block partials of the pair
// intrinsics' gradients: partial [N*H][nblk][3] = (d pp_pair.x, d pp_pair.y, d f_pair)
// kPacked: instead of the ten planar gradient planes, write what the rasterizer's backward needs from a pixel in packed
// form -- three float4 per pixel, planes [N*H][3][S*S]:
//   q0 = { g_alpha * (1 - alpha),  sum_k g_k * image_k over the six live colour planes,  1 / softmax sum,  softmax max }
//   q1 = { g3, g4, g5, g6 }     q2 = { g7, g8, 0, 0 }          (planes 0-2, the frame's own positions, are detached)
// so that a (pixel, face) pair costs it three coalesced 128-bit loads instead of sixteen scalar plane loads
// (raster_backward_pair.inc, kPacked).  `aggrs` is the forward's [N*H,2,S,S] (softmax sum, max).
template <bool kPacked>
__global__ void __launch_bounds__(256) head_backward_kernel(HeadArgs A, const float* gsums, float* gimage,
                                                            float* partial, int nblk, const float* aggrs = nullptr,
                                                            float4* packed = nullptr) {
  __shared__ float s_w[8][3];
  const int ij = blockIdx.y;
  const int i = ij / A.H;
  const long long SS = A.SS;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  if (p < A.SS) {
    const float* img = A.image + (long long)ij * kHeadPlanes * SS;
    float* go = kPacked ? nullptr : gimage + (long long)ij * kHeadPlanes * SS + p;
    const float g0 = gsums[ij * kNumSums + 0], g1 = gsums[ij * kNumSums + 1];
    const float g4 = gsums[ij * kNumSums + 4], g5 = gsums[ij * kNumSums + 5];
    const float occ = A.occ[i * SS + p];
    const float m = A.masks[i * SS + p];
    const bool v = occ != 0.f;
    const bool fgm = m > 0.f;
    const float alpha = img[9 * SS + p];
    float galpha = v ? 2.f * (alpha - m) * g0 : 0.f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (v && fgm) {
      const HeadFlow fl = head_flow(A, img, ij, p);
      if (!fl.bg) {
        const float dx = fl.fx - A.flow_obs[(i * 3 + 0) * SS + p];
        const float dy = fl.fy - A.flow_obs[(i * 3 + 1) * SS + p];
        const float nrm = sqrtf(dx * dx + dy * dy);
        if (nrm > 0.f) {
          const float c = g1 / (1.f + expf(occ)) / nrm;
          const float ox = c * dx, oy = c * dy;
          const float f = A.f_pair[head_pair_index(A, ij)], iz = 1.f / fl.z1;
          gx = ox * f * iz;
          gy = oy * f * iz;
          gz = -(ox * fl.x1 + oy * fl.y1) * f * iz * iz;
          acc[0] = ox; acc[1] = oy; acc[2] = (ox * fl.x1 + oy * fl.y1) * iz;
        }
      }
    }
    if (!kPacked) {
      go[0] = 0.f; go[SS] = 0.f; go[2 * SS] = 0.f;
      go[3 * SS] = gx; go[4 * SS] = gy; go[5 * SS] = gz;
    }
    float gfg = 0.f, gtex[3], tsum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gt = 0.f;
      if (v) {
        const float im = A.imgs[(i * 3 + c) * SS + p];
        const float obs = fgm ? im : 0.f;
        const float white = fgm ? im : 1.f;
        const float t = img[(6 + c) * SS + p];
        const float ra = obs - t * alpha, rb = white - t;
        const float sa = ra > 0.f ? 1.f : (ra < 0.f ? -1.f : 0.f);
        const float sb = rb > 0.f ? 1.f : (rb < 0.f ? -1.f : 0.f);
        gt = (-sa * alpha * g4 - sb * g5) * (1.f / 3.f);
        gfg += -sa * t * g4 * (1.f / 3.f);
        tsum += gt * t;
      }
      gtex[c] = gt;
      if (!kPacked) go[(6 + c) * SS] = gt;
    }
    if (!kPacked) {
      go[9 * SS] = galpha + gfg;
    } else {
      if (gx != 0.f || gy != 0.f || gz != 0.f)       // the pair-frame positions rendered at this pixel (planes 3-5)
        tsum += gx * img[3 * SS + p] + gy * img[4 * SS + p] + gz * img[5 * SS + p];
      const float* ag = aggrs + (long long)ij * 2 * SS + p;
      float4* pk = packed + (long long)ij * 3 * SS + p;
      pk[0] = make_float4((galpha + gfg) * (1.f - alpha), tsum, __fdividef(1.f, ag[0]), ag[SS]);
      pk[SS] = make_float4(gx, gy, gz, gtex[0]);
      pk[2 * SS] = make_float4(gtex[1], gtex[2], 0.f, 0.f);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float w = wsum(acc[k]);
    if (lane == 0) s_w[warp][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w][threadIdx.x];
    partial[((long long)ij * nblk + blockIdx.x) * 3 + threadIdx.x] = t;
  }
}

static int fill_head(HeadArgs& A, const float* image, const float* pp_own, const float* pp_pair, const float* f_own,
                     const float* f_pair, const float* imgs, const float* masks, const float* occ,
                     const float* flow_obs, int N, int H, int S, int pair_shift = 0) {
  if (N <= 0 || H <= 0 || S <= 0) return lasr_fail("head: N, H, S must be positive");
  if (pair_shift < 0 || pair_shift >= N * H) return lasr_fail("head: pair_shift must lie in [0, N*H)");
  A.pair_shift = pair_shift;
  if (!image || !pp_own || !pp_pair || !f_own || !f_pair || !imgs || !masks || !occ || !flow_obs)
    return lasr_fail("head: null tensor pointer");
  A.image = image; A.pp_own = pp_own; A.pp_pair = pp_pair; A.f_own = f_own; A.f_pair = f_pair;
  A.imgs = imgs; A.masks = masks; A.occ = occ; A.flow_obs = flow_obs;
  A.N = N; A.H = H; A.SS = S * S;
  return 0;
}

static int fill_loss(LossArgs& A, const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                     const float* tex, const float* fg, const float* imgs, const float* masks, const float* occ,
                     const float* flow_obs, int N, int H, int S) {
  if (N <= 0 || H <= 0 || S <= 0) return lasr_fail("losses: N, H, S must be positive");
  if (!mask_pred || !masks || !occ) return lasr_fail("losses: mask_pred, masks, occ are required");
  A.has_flow = flow_rd != nullptr;
  A.has_tex = tex != nullptr;
  if (A.has_flow && (!bgmask || !flow_obs)) return lasr_fail("losses: flow loss needs bgmask and flow_obs");
  if (A.has_tex && (!fg || !imgs)) return lasr_fail("losses: texture loss needs fg and imgs");
  A.mask_pred = mask_pred; A.flow_rd = flow_rd; A.bgmask = bgmask; A.tex = tex; A.fg = fg; A.imgs = imgs;
  A.masks = masks; A.occ = occ; A.flow_obs = flow_obs;
  A.N = N; A.H = H; A.SS = S * S;
  return 0;
}

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_losses_workspace_bytes(int N, int H, int S) {
  if (N <= 0 || H <= 0 || S <= 0) return 0;
  const size_t nblk = ((size_t)S * S + kLossPixPerBlock - 1) / kLossPixPerBlock;
  return (size_t)N * H * nblk * kNumSums * sizeof(float) + 256;
}

// partials -> sums (and, when `sub`/`out3` are given, the per-pair losses and their means in the same launch)
static int reduce_tail(const float* partial, int nblk, int N, int H, float* sums, float* sub, float* out3, float l1tex_wt,
                       cudaStream_t st) {
  if ((sub == nullptr) != (out3 == nullptr)) return lasr_fail("losses: sub and out3 go together");
  if (sub) {
    losses_reduce_finalize_kernel<<<1, 1024, 0, st>>>(partial, nblk, N, H, l1tex_wt, sums, sub, out3);
    return lasr_check_launch("losses_reduce_finalize_kernel");
  }
  const int tot = N * H * kNumSums;
  losses_reduce_kernel<<<(tot + 7) / 8, 256, 0, st>>>(partial, nblk, N * H, sums);
  return lasr_check_launch("losses_reduce_kernel");
}

extern "C" int lasr_b200_losses_forward_total(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                                              const float* tex, const float* fg, const float* imgs, const float* masks,
                                              const float* occ, const float* flow_obs, float* sums, float* sub,
                                              float* out3, float l1tex_wt, void* workspace, size_t workspace_bytes,
                                              int N, int H, int S, void* stream) {
  LossArgs A{};
  if (fill_loss(A, mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, N, H, S)) return 1;
  if (!sums || !workspace || workspace_bytes < lasr_b200_losses_workspace_bytes(N, H, S))
    return lasr_fail("losses_forward: null output or workspace too small");
  const int nblk = (A.SS + kLossPixPerBlock - 1) / kLossPixPerBlock;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  if (A.SS % 4 == 0) losses_forward_kernel<4><<<dim3(nblk, N * H), kLossThreads, 0, st>>>(A, partial, nblk);
  else losses_forward_kernel<1><<<dim3(nblk, N * H), kLossThreads, 0, st>>>(A, partial, nblk);
  if (lasr_check_launch("losses_forward_kernel")) return 1;
  return reduce_tail(partial, nblk, N, H, sums, sub, out3, l1tex_wt, st);
}

extern "C" int lasr_b200_losses_forward(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                                        const float* tex, const float* fg, const float* imgs, const float* masks,
                                        const float* occ, const float* flow_obs, float* sums, void* workspace,
                                        size_t workspace_bytes, int N, int H, int S, void* stream) {
  return lasr_b200_losses_forward_total(mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, sums, nullptr,
                                        nullptr, 0.f, workspace, workspace_bytes, N, H, S, stream);
}

extern "C" int lasr_b200_losses_backward(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                                         const float* tex, const float* fg, const float* imgs, const float* masks,
                                         const float* occ, const float* flow_obs, const float* grad_sums,
                                         float* grad_mask_pred, float* grad_flow_rd, float* grad_tex, float* grad_fg,
                                         int N, int H, int S, void* stream) {
  LossArgs A{};
  if (fill_loss(A, mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, N, H, S)) return 1;
  if (!grad_sums) return lasr_fail("losses_backward: grad_sums required");
  cudaStream_t st = (cudaStream_t)stream;
  if (A.SS % 4 == 0)
    losses_backward_kernel<4><<<dim3((A.SS / 4 + kLossThreads - 1) / kLossThreads, N * H), kLossThreads, 0, st>>>(
        A, grad_sums, grad_mask_pred, grad_flow_rd, grad_tex, grad_fg);
  else
    losses_backward_kernel<1><<<dim3((A.SS + kLossThreads - 1) / kLossThreads, N * H), kLossThreads, 0, st>>>(
        A, grad_sums, grad_mask_pred, grad_flow_rd, grad_tex, grad_fg);
  return lasr_check_launch("losses_backward_kernel");
}

extern "C" int lasr_b200_head_backward_packed_shift(const float* image, const float* pp_own, const float* pp_pair,
                                                    const float* f_own, const float* f_pair, const float* imgs,
                                                    const float* masks, const float* occ, const float* flow_obs,
                                                    const float* grad_sums, const float* aggrs_info, float* packed_grad,
                                                    float* grad_pp_pair, float* grad_f_pair, void* workspace,
                                                    size_t workspace_bytes, int N, int H, int S, int pair_shift,
                                                    void* stream) {
  HeadArgs A{};
  if (fill_head(A, image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, N, H, S, pair_shift)) return 1;
  if (!grad_sums || !aggrs_info || !packed_grad || !grad_pp_pair || !grad_f_pair)
    return lasr_fail("head_backward_packed: null pointer");
  if (!workspace || workspace_bytes < lasr_b200_head_workspace_bytes(N, H, S))
    return lasr_fail("head_backward_packed: workspace too small");
  const int nblk = (A.SS + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  head_backward_kernel<true><<<dim3(nblk, N * H), 256, 0, st>>>(A, grad_sums, nullptr, partial, nblk, aggrs_info,
                                                                  reinterpret_cast<float4*>(packed_grad));
  if (lasr_check_launch("head_backward_kernel<packed>")) return 1;
  flow_epilogue_reduce_kernel<<<(N * H * 3 + 3) / 4, 128, 0, st>>>(partial, N * H, nblk, grad_pp_pair, grad_f_pair,
                                                                   pair_shift);
  return lasr_check_launch("flow_epilogue_reduce_kernel");
}

extern "C" int lasr_b200_head_backward_packed(const float* image, const float* pp_own, const float* pp_pair,
                                              const float* f_own, const float* f_pair, const float* imgs,
                                              const float* masks, const float* occ, const float* flow_obs,
                                              const float* grad_sums, const float* aggrs_info, float* packed_grad,
                                              float* grad_pp_pair, float* grad_f_pair, void* workspace,
                                              size_t workspace_bytes, int N, int H, int S, void* stream) {
  return lasr_b200_head_backward_packed_shift(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, grad_sums,
                                              aggrs_info, packed_grad, grad_pp_pair, grad_f_pair, workspace,
                                              workspace_bytes, N, H, S, 0, stream);
}

extern "C" int lasr_b200_losses_finalize(const float* sums, float* sub, float* out3, int N, int H, float l1tex_wt,
                                         void* stream) {
  if (!sums || !sub || !out3 || N <= 0 || H <= 0) return lasr_fail("losses_finalize: bad arguments");
  losses_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(sums, N, H, l1tex_wt, sub, out3);
  return lasr_check_launch("losses_finalize_kernel");
}

static int coefficients_impl(const float* sums, const float* grad_out, const float* grad_sub, float* grad_sums, int N,
                             int H, float l1tex_wt, int with_total, void* stream) {
  if (!sums || !grad_out || !grad_sums || N <= 0 || H <= 0) return lasr_fail("losses_coefficients: bad arguments");
  const int NH = N * H;
  losses_coef_kernel<<<(NH + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, grad_out, grad_sub, N, H, l1tex_wt,
                                                                         grad_sums, with_total);
  return lasr_check_launch("losses_coef_kernel");
}

extern "C" int lasr_b200_losses_coefficients(const float* sums, const float* grad_out3, const float* grad_sub,
                                             float* grad_sums, int N, int H, float l1tex_wt, void* stream) {
  return coefficients_impl(sums, grad_out3, grad_sub, grad_sums, N, H, l1tex_wt, 0, stream);
}

extern "C" int lasr_b200_losses_coefficients_total(const float* sums, const float* grad_out4, const float* grad_sub,
                                                   float* grad_sums, int N, int H, float l1tex_wt, void* stream) {
  return coefficients_impl(sums, grad_out4, grad_sub, grad_sums, N, H, l1tex_wt, 1, stream);
}

extern "C" size_t lasr_b200_flow_epilogue_workspace_bytes(int n, int S) {
  if (n <= 0 || S <= 0) return 0;
  return (size_t)n * (((size_t)S * S + 255) / 256) * 3 * sizeof(float) + 256;
}

extern "C" int lasr_b200_flow_epilogue_forward(const float* px, const float* pp0, const float* pp1, const float* f0,
                                               const float* f1, float* flow, uint8_t* bgmask, int n, int S,
                                               void* stream) {
  if (!px || !pp0 || !pp1 || !f0 || !f1 || !flow || !bgmask || n <= 0 || S <= 0)
    return lasr_fail("flow_epilogue_forward: bad arguments");
  const int SS = S * S;
  flow_epilogue_forward_kernel<<<dim3((SS + 255) / 256, n), 256, 0, (cudaStream_t)stream>>>(px, pp0, pp1, f0, f1, flow,
                                                                                             bgmask, n, SS);
  return lasr_check_launch("flow_epilogue_forward_kernel");
}

extern "C" int lasr_b200_flow_epilogue_backward(const float* px, const float* f1, const float* grad_flow,
                                                float* grad_px, float* grad_pp1, float* grad_f1, void* workspace,
                                                size_t workspace_bytes, int n, int S, void* stream) {
  if (!px || !f1 || !grad_flow || !grad_px || !grad_pp1 || !grad_f1 || n <= 0 || S <= 0)
    return lasr_fail("flow_epilogue_backward: bad arguments");
  if (!workspace || workspace_bytes < lasr_b200_flow_epilogue_workspace_bytes(n, S))
    return lasr_fail("flow_epilogue_backward: workspace too small");
  const int SS = S * S, nblk = (SS + 255) / 256;
  float* partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  flow_epilogue_backward_kernel<<<dim3(nblk, n), 256, 0, st>>>(px, f1, grad_flow, grad_px, partial, n, SS, nblk);
  if (lasr_check_launch("flow_epilogue_backward_kernel")) return 1;
  flow_epilogue_reduce_kernel<<<(n * 3 + 3) / 4, 128, 0, st>>>(partial, n, nblk, grad_pp1, grad_f1);
  return lasr_check_launch("flow_epilogue_reduce_kernel");
}

extern "C" size_t lasr_b200_head_workspace_bytes(int N, int H, int S) {
  if (N <= 0 || H <= 0 || S <= 0) return 0;
  const size_t SS = (size_t)S * S;
  const size_t a = (size_t)N * H * ((SS + kHeadPixPerBlock - 1) / kHeadPixPerBlock) * kNumSums * sizeof(float);
  const size_t b = (size_t)N * H * ((SS + 255) / 256) * 3 * sizeof(float);
  return (a > b ? a : b) + 256;
}

extern "C" int lasr_b200_head_forward_total(const float* image, const float* pp_own, const float* pp_pair,
                                            const float* f_own, const float* f_pair, const float* imgs,
                                            const float* masks, const float* occ, const float* flow_obs, float* sums,
                                            float* sub, float* out3, float l1tex_wt, float* flow_rd, uint8_t* bgmask,
                                            void* workspace, size_t workspace_bytes, int N, int H, int S,
                                            int pair_shift, void* stream) {
  HeadArgs A{};
  if (fill_head(A, image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, N, H, S, pair_shift)) return 1;
  if (!sums || !workspace || workspace_bytes < lasr_b200_head_workspace_bytes(N, H, S))
    return lasr_fail("head_forward: null output or workspace too small");
  if ((flow_rd == nullptr) != (bgmask == nullptr)) return lasr_fail("head_forward: flow_rd and bgmask go together");
  const int nblk = (A.SS + kHeadPixPerBlock - 1) / kHeadPixPerBlock;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  if (A.SS % 4 == 0) head_forward_kernel<4><<<dim3(nblk, N * H), kLossThreads, 0, st>>>(A, partial, nblk, flow_rd, bgmask);
  else head_forward_kernel<1><<<dim3(nblk, N * H), kLossThreads, 0, st>>>(A, partial, nblk, flow_rd, bgmask);
  if (lasr_check_launch("head_forward_kernel")) return 1;
  return reduce_tail(partial, nblk, N, H, sums, sub, out3, l1tex_wt, st);
}

extern "C" int lasr_b200_head_forward(const float* image, const float* pp_own, const float* pp_pair,
                                      const float* f_own, const float* f_pair, const float* imgs,
                                      const float* masks, const float* occ, const float* flow_obs, float* sums,
                                      float* flow_rd, uint8_t* bgmask, void* workspace, size_t workspace_bytes, int N,
                                      int H, int S, void* stream) {
  return lasr_b200_head_forward_total(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, sums, nullptr,
                                      nullptr, 0.f, flow_rd, bgmask, workspace, workspace_bytes, N, H, S, 0, stream);
}

extern "C" int lasr_b200_head_backward(const float* image, const float* pp_own, const float* pp_pair,
                                       const float* f_own, const float* f_pair, const float* imgs,
                                       const float* masks, const float* occ, const float* flow_obs,
                                       const float* grad_sums, float* grad_image, float* grad_pp_pair,
                                       float* grad_f_pair, void* workspace, size_t workspace_bytes, int N, int H,
                                       int S, void* stream) {
  HeadArgs A{};
  if (fill_head(A, image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, N, H, S)) return 1;
  if (!grad_sums || !grad_image || !grad_pp_pair || !grad_f_pair) return lasr_fail("head_backward: null pointer");
  if (!workspace || workspace_bytes < lasr_b200_head_workspace_bytes(N, H, S))
    return lasr_fail("head_backward: workspace too small");
  const int nblk = (A.SS + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  head_backward_kernel<false><<<dim3(nblk, N * H), 256, 0, st>>>(A, grad_sums, grad_image, partial, nblk);
  if (lasr_check_launch("head_backward_kernel")) return 1;
  flow_epilogue_reduce_kernel<<<(N * H * 3 + 3) / 4, 128, 0, st>>>(partial, N * H, nblk, grad_pp_pair, grad_f_pair);
  return lasr_check_launch("flow_epilogue_reduce_kernel");
}
