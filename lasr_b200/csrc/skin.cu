// skin.cu -- Gaussian-mixture skinning weights, sm_100a.
//
// The reference computes them inline in LASR.forward (nnutils/mesh_net.py:264-267):
//   dis  = (ctl_ts[h,b] - mean_v[h,v]) . R(ctl_rs[h,b])          (kornia quaternion_to_rotation_matrix, XYZW)
//   skin = softmax_b( -10 * sum_d exp(log_ctl[h,b,d]) * dis_d^2 )
// as ~45 one-block PyTorch kernels for the quaternion algebra plus a [H,B-1,V,3] matmul, pow, exp, sum and
// softmax, and about as many again in the backward.  Here: one kernel per direction over (hypothesis,
// vertex) with the B-1 bone frames staged in shared memory, and a per-bone finishing kernel that sums the
// per-block partials in block order (deterministic) and carries the gradient through the quaternion
// normalisation.  mean_v is detached by the caller, as in the reference (:266).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "host_common.h"
#include "lasr_b200.h"
#include "raster_common.cuh"   // warp_sum16

namespace lasr {

constexpr int kSkinThreads = 256;
constexpr int kSkinMaxBones = 128;   // bone frames staged per block (16 floats each)
constexpr int kSkinVals = 15;        // per-bone gradient sums: ctl_ts (3) | exp-scale (3) | R (9)

struct BoneFrame { float c[3]; float e[3]; float R[9]; float pad; };   // 16 floats

// normalised-quaternion rotation matrix, same expression tree as lasr_b200.geom_utils.quaternion_to_rotation_matrix
__device__ __forceinline__ void quat_to_matrix(const float* q, float* R, float* qn_out, float* inv_norm_out) {
  const float nrm = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  const float x = q[0] * inv, y = q[1] * inv, z = q[2] * inv, w = q[3] * inv;
  const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
  if (qn_out) { qn_out[0] = x; qn_out[1] = y; qn_out[2] = z; qn_out[3] = w; }
  if (inv_norm_out) *inv_norm_out = inv;
}

__device__ __forceinline__ void stage_bones(BoneFrame* s_bone, const float* ctl_ts, const float* ctl_rs,
                                            const float* log_ctl, int h, int Bp) {
  for (int b = threadIdx.x; b < Bp; b += blockDim.x) {
    const long long hb = (long long)h * Bp + b;
    BoneFrame f;
#pragma unroll
    for (int d = 0; d < 3; ++d) { f.c[d] = ctl_ts[hb * 3 + d]; f.e[d] = expf(log_ctl[hb * 3 + d]); }
    quat_to_matrix(ctl_rs + hb * 4, f.R, nullptr, nullptr);
    f.pad = 0.f;
    s_bone[b] = f;
  }
  __syncthreads();
}

// bone-frame offsets of a vertex and the logit; d_j = sum_i (c_i - v_i) R[i][j] accumulated i = 0,1,2
__device__ __forceinline__ float bone_logit(const BoneFrame& f, const float* v, float* d) {
  const float a0 = f.c[0] - v[0], a1 = f.c[1] - v[1], a2 = f.c[2] - v[2];
#pragma unroll
  for (int j = 0; j < 3; ++j) d[j] = fmaf(a2, f.R[6 + j], fmaf(a1, f.R[3 + j], a0 * f.R[j]));
  const float dist = (f.e[0] * (d[0] * d[0]) + f.e[1] * (d[1] * d[1])) + f.e[2] * (d[2] * d[2]);
  return -10.f * dist;
}

__global__ void __launch_bounds__(kSkinThreads) skin_forward_kernel(const float* __restrict__ ctl_ts,
                                                                    const float* __restrict__ ctl_rs,
                                                                    const float* __restrict__ log_ctl,
                                                                    const float* __restrict__ verts, float* skin,
                                                                    int Bp, int V) {
  __shared__ BoneFrame s_bone[kSkinMaxBones];
  const int h = blockIdx.y;
  stage_bones(s_bone, ctl_ts, ctl_rs, log_ctl, h, Bp);
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* vp = verts + ((long long)h * V + v) * 3;
  const float vv[3] = {vp[0], vp[1], vp[2]};
  float d[3];
  float mx = -INFINITY;
  for (int b = 0; b < Bp; ++b) mx = fmaxf(mx, bone_logit(s_bone[b], vv, d));
  float sum = 0.f;
  float* out = skin + (long long)h * Bp * V + v;
  for (int b = 0; b < Bp; ++b) {
    const float e = expf(bone_logit(s_bone[b], vv, d) - mx);
    out[(long long)b * V] = e;
    sum += e;
  }
  for (int b = 0; b < Bp; ++b) out[(long long)b * V] = out[(long long)b * V] / sum;
}

// grad_skin [H,Bp,V] -> per-block partial sums [H][nblk][Bp][15]
// (A) one thread per vertex: the softmax adjoint's dot product, vertex and dot staged in shared memory; (B) one WARP per
// bone: lanes stride over the block's vertices accumulating the bone's 15 sums in registers, one recursive-halving
// reduction per bone.  (Round 1 reduced bone after bone across the block: two __syncthreads and 31 shuffles per bone in
// every warp -- 17 us for 25 bones on 642 vertices, where three blocks are all the parallelism there is.)
__global__ void __launch_bounds__(kSkinThreads) skin_backward_kernel(const float* __restrict__ ctl_ts,
                                                                     const float* __restrict__ ctl_rs,
                                                                     const float* __restrict__ log_ctl,
                                                                     const float* __restrict__ verts,
                                                                     const float* __restrict__ skin,
                                                                     const float* __restrict__ gskin, float* partial,
                                                                     int Bp, int V) {
  __shared__ BoneFrame s_bone[kSkinMaxBones];
  __shared__ float s_v[4][kSkinThreads];     // x, y, z, sum_k g_k s_k
  const int h = blockIdx.y;
  stage_bones(s_bone, ctl_ts, ctl_rs, log_ctl, h, Bp);
  const int v0 = blockIdx.x * kSkinThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* sp = skin + (long long)h * Bp * V + v0;
  const float* gp = gskin + (long long)h * Bp * V + v0;
  {
    const int v = v0 + threadIdx.x;
    float vv[3] = {0.f, 0.f, 0.f}, dot = 0.f;
    if (v < V) {
      const float* vp = verts + ((long long)h * V + v) * 3;
      vv[0] = vp[0]; vv[1] = vp[1]; vv[2] = vp[2];
      for (int b = 0; b < Bp; ++b) dot += gp[(long long)b * V + threadIdx.x] * sp[(long long)b * V + threadIdx.x];
    }
    s_v[0][threadIdx.x] = vv[0]; s_v[1][threadIdx.x] = vv[1]; s_v[2][threadIdx.x] = vv[2]; s_v[3][threadIdx.x] = dot;
  }
  __syncthreads();
  const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  float* pout = partial + ((long long)h * gridDim.x + blockIdx.x) * Bp * kSkinVals;
  for (int b = warp; b < Bp; b += kSkinThreads / 32) {
    const BoneFrame& f = s_bone[b];
    float val[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) val[k] = 0.f;
    float sw[kSkinThreads / 32], gw[kSkinThreads / 32];
#pragma unroll
    for (int i = 0; i < kSkinThreads / 32; ++i) {        // all loads of the bone first: one round trip, not eight
      const int t = lane + 32 * i;
      const bool live = v0 + t < V;
      sw[i] = live ? __ldg(sp + (long long)b * V + t) : 0.f;
      gw[i] = live ? __ldg(gp + (long long)b * V + t) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kSkinThreads / 32; ++i) {
      const int t = lane + 32 * i;
      if (v0 + t < V) {
        const float vv[3] = {s_v[0][t], s_v[1][t], s_v[2][t]};
        float d[3];
        bone_logit(f, vv, d);
        const float gl = sw[i] * (gw[i] - s_v[3][t]);          // d loss / d logit
        const float gdist = -10.f * gl;
        float gd[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          val[3 + j] += gdist * d[j] * d[j];                   // d / d exp(log_ctl_j)
          gd[j] = 2.f * gdist * f.e[j] * d[j];
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          val[q] += gd[0] * f.R[3 * q] + gd[1] * f.R[3 * q + 1] + gd[2] * f.R[3 * q + 2];   // d / d ctl_ts_q
          const float a = f.c[q] - vv[q];
#pragma unroll
          for (int j = 0; j < 3; ++j) val[6 + 3 * q + j] += a * gd[j];                       // d / d R[q][j]
        }
      }
    }
    warp_sum16(val, lane);
    if ((lane & 1) == 0 && vidx < kSkinVals) pout[b * kSkinVals + vidx] = val[0];
  }
}

// one thread per (hypothesis, bone): sum the block partials in block order, then carry the matrix gradient
// through the unit-quaternion matrix and the normalisation
__global__ void skin_finish_kernel(const float* __restrict__ partial, const float* __restrict__ ctl_rs,
                                   const float* __restrict__ log_ctl, float* g_ts, float* g_rs, float* g_log, int H,
                                   int Bp, int nblk) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= H * Bp) return;
  const int h = t / Bp, b = t % Bp;
  float a[kSkinVals];
#pragma unroll
  for (int k = 0; k < kSkinVals; ++k) a[k] = 0.f;
  for (int blk = 0; blk < nblk; ++blk) {
    const float* p = partial + (((long long)h * nblk + blk) * Bp + b) * kSkinVals;
#pragma unroll
    for (int k = 0; k < kSkinVals; ++k) a[k] += p[k];
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g_ts[t * 3 + d] = a[d];
    g_log[t * 3 + d] = a[3 + d] * expf(log_ctl[t * 3 + d]);
  }
  const float* G = a + 6;   // dL/dR, row-major
  float R[9], q[4], inv;
  quat_to_matrix(ctl_rs + (long long)t * 4, R, q, &inv);
  const float x = q[0], y = q[1], z = q[2], w = q[3];
  // R = [[1-2(yy+zz), 2(xy-zw), 2(xz+yw)], [2(xy+zw), 1-2(xx+zz), 2(yz-xw)], [2(xz-yw), 2(yz+xw), 1-2(xx+yy)]]
  float gq[4];
  gq[0] = 2.f * (y * (G[1] + G[3]) + z * (G[2] + G[6]) - 2.f * x * (G[4] + G[8]) + w * (G[7] - G[5]));
  gq[1] = 2.f * (x * (G[1] + G[3]) + z * (G[5] + G[7]) - 2.f * y * (G[0] + G[8]) + w * (G[2] - G[6]));
  gq[2] = 2.f * (x * (G[2] + G[6]) + y * (G[5] + G[7]) - 2.f * z * (G[0] + G[4]) + w * (G[3] - G[1]));
  gq[3] = 2.f * (x * (G[7] - G[5]) + y * (G[2] - G[6]) + z * (G[3] - G[1]));
  // q_n = q / |q|  =>  dq = (dq_n - q_n (q_n . dq_n)) / |q|
  const float dotq = gq[0] * x + gq[1] * y + gq[2] * z + gq[3] * w;
#pragma unroll
  for (int k = 0; k < 4; ++k) g_rs[t * 4 + k] = (gq[k] - q[k] * dotq) * inv;
}

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_skin_workspace_bytes(int H, int Bp, int V) {
  if (H <= 0 || Bp <= 0 || V <= 0) return 0;
  const size_t nblk = ((size_t)V + kSkinThreads - 1) / kSkinThreads;
  return (size_t)H * nblk * Bp * kSkinVals * sizeof(float) + 256;
}

static int skin_check(const void* a, const void* b, const void* c, const void* d, int H, int Bp, int V) {
  if (H <= 0 || Bp <= 0 || V <= 0) return lasr_fail("skin_weights: H, bones, V must be positive");
  if (Bp > kSkinMaxBones) return lasr_fail("skin_weights: more than 128 part bones per hypothesis unsupported");
  if (!a || !b || !c || !d) return lasr_fail("skin_weights: null tensor pointer");
  return 0;
}

extern "C" int lasr_b200_skin_weights_forward(const float* ctl_ts, const float* ctl_rs, const float* log_ctl,
                                              const float* verts, float* skin, int H, int Bp, int V, void* stream) {
  if (skin_check(ctl_ts, ctl_rs, log_ctl, verts, H, Bp, V) || !skin)
    return skin ? 1 : lasr_fail("skin_weights_forward: null output");
  const dim3 grid((V + kSkinThreads - 1) / kSkinThreads, H);
  skin_forward_kernel<<<grid, kSkinThreads, 0, (cudaStream_t)stream>>>(ctl_ts, ctl_rs, log_ctl, verts, skin, Bp, V);
  return lasr_check_launch("skin_forward_kernel");
}

extern "C" int lasr_b200_skin_weights_backward(const float* ctl_ts, const float* ctl_rs, const float* log_ctl,
                                               const float* verts, const float* skin, const float* grad_skin,
                                               float* grad_ctl_ts, float* grad_ctl_rs, float* grad_log_ctl,
                                               void* workspace, size_t workspace_bytes, int H, int Bp, int V,
                                               void* stream) {
  if (skin_check(ctl_ts, ctl_rs, log_ctl, verts, H, Bp, V)) return 1;
  if (!skin || !grad_skin || !grad_ctl_ts || !grad_ctl_rs || !grad_log_ctl)
    return lasr_fail("skin_weights_backward: null tensor pointer");
  if (!workspace || workspace_bytes < lasr_b200_skin_workspace_bytes(H, Bp, V))
    return lasr_fail("skin_weights_backward: workspace too small");
  const int nblk = (V + kSkinThreads - 1) / kSkinThreads;
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  skin_backward_kernel<<<dim3(nblk, H), kSkinThreads, 0, st>>>(ctl_ts, ctl_rs, log_ctl, verts, skin, grad_skin, partial,
                                                               Bp, V);
  if (lasr_check_launch("skin_backward_kernel")) return 1;
  skin_finish_kernel<<<(H * Bp + 127) / 128, 128, 0, st>>>(partial, ctl_rs, log_ctl, grad_ctl_ts, grad_ctl_rs,
                                                           grad_log_ctl, H, Bp, nblk);
  return lasr_check_launch("skin_finish_kernel");
}
