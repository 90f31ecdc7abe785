// lbs.cu -- fused linear-blend skinning + rigid body transform + pinhole projection + batch depth
// window, forward and backward, for sm_100a.
//
// Replaces, on CUDA tensors, the reference's dense PyTorch glue
//   obj_to_cam   nnutils/geom_utils.py:45-71   (B-1 batched matmuls + cat + weighted sum; the
//                                               [N,B-1,V,3] intermediate is 25x the output)
//   pinhole_cam  nnutils/geom_utils.py:27-34   (6 clones + 2 in-place slice assigns)
//   near/far     nnutils/mesh_net.py:304-311   (2 reductions + a device->host sync at the pybind call)
// with one pass over the vertices: every thread owns one (frame, vertex), the frame's bone
// transforms sit in shared memory, skinning weights are read coalesced ([B-1, V] rows), and the
// posed vertex never leaves registers between skinning, the body transform and the projection.
// Row-vector convention of the reference:  p_k = v R_k + T_k,  vs = sum_k w_k p_k,
// cam = vs R_0 + T_0,  px = (pp_x + cam_x f / cam_z, pp_y + cam_y f / cam_z, cam_z, 1).
//
// Backward is deterministic: per-vertex gradients are written per frame; the per-bone (R, T) and
// per-frame (pp, f) gradients are block-reduced with fixed-order shuffles into a partials buffer
// and summed by a second tiny kernel (no float atomics).
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>
#include <string>

#include "host_common.h"
#include "lasr_b200.h"
#include "raster_common.cuh"   // warp_sum16

namespace lasr {

constexpr int kLbsThreads = 256;

struct LbsArgs {
  const float* verts;   // [Nv, V, 3]   Nv = N or 1 (shared mesh)
  const float* Rmat;    // [N, B, 3, 3]
  const float* Tmat;    // [N, B, 3]
  const float* skin;    // [Ns, B-1, V] Ns = N or 1 (shared weights); may be null when B == 1
  const float* pp;      // [N, 2] or null (no projection)
  const float* fl;      // [N]    or null
  float* cam;           // [N, V, 3] camera-space (or skinned-only when !tocam)
  float* px;            // [N, V, 4] projected (x, y, z, 1); null when no projection
  unsigned* zminmax;    // [2] order-preserving encodings of min / max z (null: skip)
  int N, V, B;
  int verts_shared, skin_shared, tocam;
  // Optional composition of the part transforms about their rest positions (nnutils/mesh_net.py:278-281), done while the
  // frame's bones are staged in shared memory instead of by ~30 small PyTorch kernels per step:
  //   R'_k = R_k^T,   T'_k = -R_k rest_k + T_k + rest_k      (k >= 1; the body transform k = 0 is used as given)
  const float* rest;    // [n_hypo*(B-1), 3] joint centres, shared by the frames; null: Rmat / Tmat are used as given
  int n_hypo;           // frames are laid out [.., n_hypo]: frame n uses the bones of hypothesis n % n_hypo
  // Optional third output: the projected vertex as LASR hands it to the rasterizer -- + eye offset, y flip, then the
  // renderer's look_at(eye) with the eye on the z axis (mesh_net.py:81-83, 354-357; transform.py:41-48): (x, -y, (z + e) - e),
  // the float additions kept so that the values are the PyTorch chain's bit for bit (three small kernels forward, four back).
  float* view;          // [N, V, 3] or null
  float eye_z;
};

// the frame's B transforms -> shared memory, [B][12]: R (9, row-major) then T (3)
__device__ __forceinline__ void stage_bones(const LbsArgs& A, int n, float* s_rt, int nthreads) {
  for (int i = threadIdx.x; i < A.B * 12; i += nthreads) {
    const int k = i / 12, j = i % 12;
    const float* R = A.Rmat + ((long long)n * A.B + k) * 9;
    const float* T = A.Tmat + ((long long)n * A.B + k) * 3;
    float val;
    if (A.rest == nullptr || k == 0) {
      val = j < 9 ? R[j] : T[j - 9];
    } else {
      const float* q = A.rest + ((long long)(n % A.n_hypo) * (A.B - 1) + (k - 1)) * 3;
      if (j < 9) {
        val = R[(j % 3) * 3 + j / 3];
      } else {
        const int c = j - 9;
        val = (-(R[3 * c] * q[0] + R[3 * c + 1] * q[1] + R[3 * c + 2] * q[2]) + T[c]) + q[c];
      }
    }
    s_rt[i] = val;
  }
}

// order-preserving float <-> uint encoding for atomicMin/atomicMax
__device__ __forceinline__ unsigned enc(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void lbs_init_minmax_kernel(unsigned* zminmax) {
  zminmax[0] = 0xffffffffu;  // min slot starts at +max
  zminmax[1] = 0u;           // max slot starts at -max
}

// near = dmin - (dmax - dmin)/2, far = dmax + (dmax - dmin)/2   (mesh_net.py:304-307)
__global__ void lbs_depth_window_kernel(const unsigned* zminmax, float* near_far) {
  const float dmin = dec(zminmax[0]), dmax = dec(zminmax[1]);
  const float half = (dmax - dmin) / 2.f;
  near_far[0] = dmin - half;
  near_far[1] = dmax + half;
}

__global__ void __launch_bounds__(kLbsThreads) lbs_forward_kernel(LbsArgs A) {
  extern __shared__ float s_rt[];  // [B][12]: R (9, row-major) then T (3)
  const int n = blockIdx.y;
  const int v = blockIdx.x * kLbsThreads + threadIdx.x;
  stage_bones(A, n, s_rt, kLbsThreads);
  __syncthreads();
  float z = 0.f;
  const bool active = v < A.V;
  if (active) {
    const float* vp = A.verts + ((long long)(A.verts_shared ? 0 : n) * A.V + v) * 3;
    const float x0 = vp[0], y0 = vp[1], z0 = vp[2];
    float sx = x0, sy = y0, sz = z0;
    if (A.B > 1) {
      sx = sy = sz = 0.f;
      const float* sk = A.skin + (long long)(A.skin_shared ? 0 : n) * (A.B - 1) * A.V + v;
      for (int k = 1; k < A.B; ++k) {
        const float* r = s_rt + k * 12;
        const float w = sk[(long long)(k - 1) * A.V];
        const float px_ = x0 * r[0] + y0 * r[3] + z0 * r[6] + r[9];
        const float py_ = x0 * r[1] + y0 * r[4] + z0 * r[7] + r[10];
        const float pz_ = x0 * r[2] + y0 * r[5] + z0 * r[8] + r[11];
        sx += w * px_; sy += w * py_; sz += w * pz_;
      }
    }
    float cx = sx, cy = sy, cz = sz;
    if (A.tocam) {
      const float* r = s_rt;
      cx = sx * r[0] + sy * r[3] + sz * r[6] + r[9];
      cy = sx * r[1] + sy * r[4] + sz * r[7] + r[10];
      cz = sx * r[2] + sy * r[5] + sz * r[8] + r[11];
    }
    float* co = A.cam + ((long long)n * A.V + v) * 3;
    co[0] = cx; co[1] = cy; co[2] = cz;
    z = cz;
    if (A.px) {
      const float f = A.fl[n];   // same operation order as the reference: pp + (x * f) / z  (geom_utils.py:32-33)
      const float ux = A.pp[2 * n] + __fdiv_rn(cx * f, cz), uy = A.pp[2 * n + 1] + __fdiv_rn(cy * f, cz);
      reinterpret_cast<float4*>(A.px)[(long long)n * A.V + v] = make_float4(ux, uy, cz, 1.f);
      if (A.view) {
        float* vo = A.view + ((long long)n * A.V + v) * 3;
        vo[0] = ux; vo[1] = -uy; vo[2] = __fsub_rn(__fadd_rn(cz, A.eye_z), A.eye_z);
      }
    }
  }
  if (A.zminmax) {
    float zmin = active ? z : FLT_MAX, zmax = active ? z : -FLT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
      zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(A.zminmax, enc(zmin));      // exact: min/max are order independent
      atomicMax(A.zminmax + 1, enc(zmax));
    }
  }
}

struct LbsBwdArgs {
  LbsArgs f;
  const float* grad_cam;  // [N,V,3] or null
  const float* grad_px;   // [N,V,4] or null
  const float* grad_view; // [N,V,3] or null: gradient of the render-space output (f.view)
  float* grad_verts;      // [N,V,3]  (per frame; the caller sums frames for a shared mesh)
  float* grad_skin;       // [N,B-1,V] (per frame) or null
  float* partial;         // [N][nblk][B*12 + 3]  block partials of (R,T) per bone and (pp_x, pp_y, f)
  int nblk;
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two phases.  (A) one thread per vertex: skinned position, gradient of the skinned position gs, per-bone skin-weight
// gradients (written), vertex gradient (written); the vertex, gs, the skinned position and the camera-space gradient are
// left in shared memory.  (B) one WARP per bone: g_(R_k, T_k) = sum_v w_kv (v, 1)^T gs_v is a 12-value reduction over the
// block's 256 vertices -- lanes stride over the vertices, one butterfly per bone.  Round 1 block-reduced bone after bone
// with two __syncthreads and a skin-weight load in front of each (27 dependent rounds: 26 us for 8 frames of 642
// vertices, as long as a third of the rasterizer's backward at LASR's batch size).  Fixed summation order: deterministic.
__global__ void __launch_bounds__(kLbsThreads) lbs_backward_kernel(LbsBwdArgs G) {
  extern __shared__ float s_rt[];  // [B][12]
  __shared__ float s_st[15][kLbsThreads];   // 0-2 vertex, 3-5 gs, 6-8 skinned position, 9-11 g_cam, 12-14 (g_pp.x, g_pp.y, g_f)
  const LbsArgs& A = G.f;
  const int n = blockIdx.y;
  const int v = blockIdx.x * kLbsThreads + threadIdx.x;
  stage_bones(A, n, s_rt, kLbsThreads);
  __syncthreads();
  const bool active = v < A.V;
  float x0 = 0.f, y0 = 0.f, z0 = 0.f;
  float sx = 0.f, sy = 0.f, sz = 0.f;        // skinned (pre-body) position
  float gcx = 0.f, gcy = 0.f, gcz = 0.f;     // gradient w.r.t. the camera-space position
  float gppx = 0.f, gppy = 0.f, gfl = 0.f;
  float gsx = 0.f, gsy = 0.f, gsz = 0.f;
  const float* skb = A.B > 1 ? A.skin + (long long)(A.skin_shared ? 0 : n) * (A.B - 1) * A.V : nullptr;
  if (active) {
    const float* vp = A.verts + ((long long)(A.verts_shared ? 0 : n) * A.V + v) * 3;
    x0 = vp[0]; y0 = vp[1]; z0 = vp[2];
    sx = x0; sy = y0; sz = z0;
    const float* sk = skb ? skb + v : nullptr;
    if (A.B > 1) {
      sx = sy = sz = 0.f;
      for (int k = 1; k < A.B; ++k) {
        const float* r = s_rt + k * 12;
        const float w = sk[(long long)(k - 1) * A.V];
        sx += w * (x0 * r[0] + y0 * r[3] + z0 * r[6] + r[9]);
        sy += w * (x0 * r[1] + y0 * r[4] + z0 * r[7] + r[10]);
        sz += w * (x0 * r[2] + y0 * r[5] + z0 * r[8] + r[11]);
      }
    }
    if (G.grad_cam) {
      const float* g = G.grad_cam + ((long long)n * A.V + v) * 3;
      gcx = g[0]; gcy = g[1]; gcz = g[2];
    }
    if (G.grad_px || G.grad_view) {
      const float* r = s_rt;
      float cx = sx, cy = sy, cz = sz;
      if (A.tocam) {
        cx = sx * r[0] + sy * r[3] + sz * r[6] + r[9];
        cy = sx * r[1] + sy * r[4] + sz * r[7] + r[10];
        cz = sx * r[2] + sy * r[5] + sz * r[8] + r[11];
      }
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (G.grad_px) g = reinterpret_cast<const float4*>(G.grad_px)[(long long)n * A.V + v];
      if (G.grad_view) {                                   // view = (x, -y, (z + e) - e)
        const float* gv = G.grad_view + ((long long)n * A.V + v) * 3;
        g.x += gv[0]; g.y -= gv[1]; g.z += gv[2];
      }
      const float f = A.fl[n];
      const float iz = 1.f / cz;
      const float s = (g.x * cx + g.y * cy) * iz;   // (g_x X + g_y Y) / Z
      gcx += g.x * f * iz;
      gcy += g.y * f * iz;
      gcz += g.z - s * f * iz;
      gfl = s;
      gppx = g.x; gppy = g.y;
    }
    // body transform: cam = vs R0 + T0  ->  g_vs = g_cam R0^T
    gsx = gcx; gsy = gcy; gsz = gcz;
    if (A.tocam) {
      const float* r = s_rt;
      gsx = gcx * r[0] + gcy * r[1] + gcz * r[2];
      gsy = gcx * r[3] + gcy * r[4] + gcz * r[5];
      gsz = gcx * r[6] + gcy * r[7] + gcz * r[8];
    }
    float gvx = gsx, gvy = gsy, gvz = gsz;   // B == 1: vs = v
    if (A.B > 1) {
      gvx = gvy = gvz = 0.f;
      for (int k = 1; k < A.B; ++k) {
        const float* r = s_rt + k * 12;
        const float w = sk[(long long)(k - 1) * A.V];
        if (G.grad_skin) {
          const float pxk = x0 * r[0] + y0 * r[3] + z0 * r[6] + r[9];
          const float pyk = x0 * r[1] + y0 * r[4] + z0 * r[7] + r[10];
          const float pzk = x0 * r[2] + y0 * r[5] + z0 * r[8] + r[11];
          G.grad_skin[((long long)n * (A.B - 1) + (k - 1)) * A.V + v] = pxk * gsx + pyk * gsy + pzk * gsz;
        }
        const float gp0 = w * gsx, gp1 = w * gsy, gp2 = w * gsz;   // g_p_k
        gvx += gp0 * r[0] + gp1 * r[1] + gp2 * r[2];               // g_v += g_p_k R_k^T
        gvy += gp0 * r[3] + gp1 * r[4] + gp2 * r[5];
        gvz += gp0 * r[6] + gp1 * r[7] + gp2 * r[8];
      }
    }
    float* go = G.grad_verts + ((long long)n * A.V + v) * 3;
    go[0] = gvx; go[1] = gvy; go[2] = gvz;
  }
  {
    const int t = threadIdx.x;     // inactive threads stage zeros: they add nothing below
    s_st[0][t] = x0; s_st[1][t] = y0; s_st[2][t] = z0;
    s_st[3][t] = gsx; s_st[4][t] = gsy; s_st[5][t] = gsz;
    s_st[6][t] = sx; s_st[7][t] = sy; s_st[8][t] = sz;
    s_st[9][t] = gcx; s_st[10][t] = gcy; s_st[11][t] = gcz;
    s_st[12][t] = gppx; s_st[13][t] = gppy; s_st[14][t] = gfl;
  }
  __syncthreads();
  // (B) item 0: body transform (g_R0 = vs^T g_cam, g_T0 = sum g_cam); items 1..B-1: bones; item B: (g_pp, g_f)
  float* part = G.partial + ((long long)n * G.nblk + blockIdx.x) * (A.B * 12 + 3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int v_blk = blockIdx.x * kLbsThreads;
  const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  for (int item = warp; item <= A.B; item += kLbsThreads / 32) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (item == A.B) {
#pragma unroll
      for (int i = 0; i < kLbsThreads / 32; ++i) {
        const int t = lane + 32 * i;
        acc[0] += s_st[12][t]; acc[1] += s_st[13][t]; acc[2] += s_st[14][t];
      }
    } else if (item == 0) {
      if (A.tocam) {
#pragma unroll
        for (int i = 0; i < kLbsThreads / 32; ++i) {
          const int t = lane + 32 * i;
          const float a[3] = {s_st[6][t], s_st[7][t], s_st[8][t]}, g[3] = {s_st[9][t], s_st[10][t], s_st[11][t]};
#pragma unroll
          for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[3 * p + q] += a[p] * g[q];
#pragma unroll
          for (int q = 0; q < 3; ++q) acc[9 + q] += g[q];
        }
      }
    } else {
      const float* wk = skb + (long long)(item - 1) * A.V + v_blk;
      float w[kLbsThreads / 32];
#pragma unroll
      for (int i = 0; i < kLbsThreads / 32; ++i) w[i] = (v_blk + lane + 32 * i < A.V) ? __ldg(wk + lane + 32 * i) : 0.f;
#pragma unroll
      for (int i = 0; i < kLbsThreads / 32; ++i) {
        const int t = lane + 32 * i;
        const float a[3] = {s_st[0][t], s_st[1][t], s_st[2][t]};
        const float g[3] = {w[i] * s_st[3][t], w[i] * s_st[4][t], w[i] * s_st[5][t]};   // g_p_k
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) acc[3 * p + q] += a[p] * g[q];                    // g_R_k = v^T g_p_k
#pragma unroll
        for (int q = 0; q < 3; ++q) acc[9 + q] += g[q];                                  // g_T_k
      }
    }
    const int nval = item == A.B ? 3 : 12;
    warp_sum16(acc, lane);     // recursive halving: every even lane ends with the warp total of value vidx
    if ((lane & 1) == 0 && vidx < nval) part[item * 12 + vidx] = acc[0];
  }
}

// sum the block partials in block order -> grad_R [N,B,9], grad_T [N,B,3], grad_pp [N,2], grad_fl [N].
// With composed part transforms (rest != null) the sums are gradients of (R', T') and are carried to the raw inputs here:
//   g_R_k[b][a] = g_R'_k[a][b] - g_T'_k[b] rest_k[a],   g_T_k = g_T'_k,   g_rest_k[c] = g_T'_k[c] - sum_i R_k[i][c] g_T'_k[i]
// (g_rest per frame, [N,B-1,3]; the caller sums the frames of a hypothesis).
__global__ void lbs_reduce_partials_kernel(const float* partial, int nblk, int N, int B, float* grad_R, float* grad_T,
                                           float* grad_pp, float* grad_fl, const float* Rmat, const float* rest,
                                           int n_hypo, float* grad_rest) {
  const int per = B * 12 + 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * per) return;
  const int n = (int)(i / per), j = (int)(i % per);
  auto total = [&](int elem) {
    float t = 0.f;
    for (int b = 0; b < nblk; ++b) t += partial[((long long)n * nblk + b) * per + elem];
    return t;
  };
  const float t = total(j);
  if (j < B * 12) {
    const int k = j / 12, e = j % 12;
    if (rest == nullptr || k == 0) {
      if (e < 9) grad_R[((long long)n * B + k) * 9 + e] = t;
      else grad_T[((long long)n * B + k) * 3 + (e - 9)] = t;
    } else {
      const float* q = rest + ((long long)(n % n_hypo) * (B - 1) + (k - 1)) * 3;
      if (e < 9) {
        const int a = e / 3, b = e % 3;
        grad_R[((long long)n * B + k) * 9 + 3 * b + a] = t - total(k * 12 + 9 + b) * q[a];
      } else {
        const int c = e - 9;
        grad_T[((long long)n * B + k) * 3 + c] = t;
        const float* R = Rmat + ((long long)n * B + k) * 9;
        const float g0 = total(k * 12 + 9), g1 = total(k * 12 + 10), g2 = total(k * 12 + 11);
        grad_rest[((long long)n * (B - 1) + (k - 1)) * 3 + c] = t - (R[c] * g0 + R[3 + c] * g1 + R[6 + c] * g2);
      }
    }
  } else if (j - B * 12 < 2) {
    if (grad_pp) grad_pp[2 * n + (j - B * 12)] = t;
  } else if (grad_fl) {
    grad_fl[n] = t;
  }
}

static int fill_lbs(LbsArgs& A, const float* verts, const float* Rmat, const float* Tmat, const float* skin,
                    const float* pp, const float* fl, int N, int V, int B, int verts_shared, int skin_shared,
                    int tocam) {
  if (N <= 0 || V <= 0 || B <= 0) return lasr_fail("lbs: N, V, B must be positive");
  if (!verts || !Rmat || !Tmat) return lasr_fail("lbs: null tensor pointer");
  if (B > 1 && !skin) return lasr_fail("lbs: skin weights required when B > 1");
  if ((pp == nullptr) != (fl == nullptr)) return lasr_fail("lbs: pp and fl must be given together");
  if (B * 12 * sizeof(float) > 40000) return lasr_fail("lbs: too many bones for shared memory");
  A.verts = verts; A.Rmat = Rmat; A.Tmat = Tmat; A.skin = skin; A.pp = pp; A.fl = fl;
  A.N = N; A.V = V; A.B = B;
  A.verts_shared = verts_shared; A.skin_shared = skin_shared; A.tocam = tocam;
  A.rest = nullptr; A.n_hypo = 1; A.view = nullptr; A.eye_z = 0.f;
  return 0;
}

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_lbs_workspace_bytes(int N, int V, int B) {
  if (N <= 0 || V <= 0 || B <= 0) return 0;
  const size_t nblk = (size_t)(V + kLbsThreads - 1) / kLbsThreads;
  return 256 + (size_t)N * nblk * ((size_t)B * 12 + 3) * sizeof(float);
}

static int lbs_forward_impl(const float* verts, const float* Rmat, const float* Tmat, const float* rest, int n_hypo,
                            const float* skin, const float* pp, const float* fl, float* cam, float* px, float* view,
                            float eye_z, float* near_far_dev, void* workspace, size_t workspace_bytes, int N, int V,
                            int B, int verts_shared, int skin_shared, int tocam, void* stream) {
  LbsArgs A{};
  if (fill_lbs(A, verts, Rmat, Tmat, skin, pp, fl, N, V, B, verts_shared, skin_shared, tocam)) return 1;
  if (rest) {
    if (n_hypo <= 0 || N % n_hypo != 0) return lasr_fail("lbs_pose: n_hypo must divide the frame count");
    A.rest = rest; A.n_hypo = n_hypo;
  }
  if (!cam) return lasr_fail("lbs_project_forward: cam output required");
  if (pp && !px) return lasr_fail("lbs_project_forward: px output required when projecting");
  A.cam = cam; A.px = pp ? px : nullptr;
  if (view) {
    if (!pp) return lasr_fail("lbs_pose_forward: the render-space output needs a projection");
    A.view = view; A.eye_z = eye_z;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (near_far_dev) {
    if (!workspace || workspace_bytes < 256) return lasr_fail("lbs_project_forward: workspace too small");
    A.zminmax = reinterpret_cast<unsigned*>(workspace);
    lbs_init_minmax_kernel<<<1, 1, 0, st>>>(A.zminmax);
    if (lasr_check_launch("lbs_init_minmax_kernel")) return 1;
  }
  dim3 grid((V + kLbsThreads - 1) / kLbsThreads, N);
  lbs_forward_kernel<<<grid, kLbsThreads, (size_t)B * 12 * sizeof(float), st>>>(A);
  if (lasr_check_launch("lbs_forward_kernel")) return 1;
  if (near_far_dev) {
    lbs_depth_window_kernel<<<1, 1, 0, st>>>(A.zminmax, near_far_dev);
    if (lasr_check_launch("lbs_depth_window_kernel")) return 1;
  }
  return 0;
}

extern "C" int lasr_b200_lbs_project_forward(const float* verts, const float* Rmat, const float* Tmat,
                                             const float* skin, const float* pp, const float* fl, float* cam,
                                             float* px, float* near_far_dev, void* workspace,
                                             size_t workspace_bytes, int N, int V, int B, int verts_shared,
                                             int skin_shared, int tocam, void* stream) {
  return lbs_forward_impl(verts, Rmat, Tmat, nullptr, 1, skin, pp, fl, cam, px, nullptr, 0.f, near_far_dev, workspace,
                          workspace_bytes, N, V, B, verts_shared, skin_shared, tocam, stream);
}

extern "C" int lasr_b200_lbs_pose_forward(const float* verts, const float* Rmat, const float* Tmat, const float* rest_ts,
                                          const float* skin, const float* pp, const float* fl, float* cam, float* px,
                                          float* view, float eye_z, float* near_far_dev, void* workspace,
                                          size_t workspace_bytes, int N, int V, int B, int n_hypo, int verts_shared,
                                          int skin_shared, int tocam, void* stream) {
  if (rest_ts && B < 2) return lasr_fail("lbs_pose_forward: rest_ts needs at least one part");
  return lbs_forward_impl(verts, Rmat, Tmat, rest_ts, n_hypo, skin, pp, fl, cam, px, view, eye_z, near_far_dev, workspace,
                          workspace_bytes, N, V, B, verts_shared, skin_shared, tocam, stream);
}

static int lbs_backward_impl(const float* verts, const float* Rmat, const float* Tmat, const float* rest, int n_hypo,
                             const float* skin, const float* pp, const float* fl, const float* grad_cam,
                             const float* grad_px, const float* grad_view, float* grad_verts, float* grad_R,
                             float* grad_T, float* grad_rest,
                             float* grad_skin, float* grad_pp, float* grad_fl, void* workspace, size_t workspace_bytes,
                             int N, int V, int B, int verts_shared, int skin_shared, int tocam, void* stream) {
  LbsBwdArgs G{};
  if (fill_lbs(G.f, verts, Rmat, Tmat, skin, pp, fl, N, V, B, verts_shared, skin_shared, tocam)) return 1;
  if (rest) {
    if (n_hypo <= 0 || N % n_hypo != 0) return lasr_fail("lbs_pose: n_hypo must divide the frame count");
    if (!grad_rest) return lasr_fail("lbs_pose_backward: grad_rest output required");
    G.f.rest = rest; G.f.n_hypo = n_hypo;
  }
  if (!grad_verts || !grad_R || !grad_T) return lasr_fail("lbs_project_backward: null gradient output");
  if (!grad_cam && !grad_px && !grad_view) return lasr_fail("lbs_project_backward: no upstream gradient");
  if ((grad_px || grad_view) && !pp) return lasr_fail("lbs_project_backward: grad_px without a projection");
  if (!workspace || workspace_bytes < lasr_b200_lbs_workspace_bytes(N, V, B))
    return lasr_fail("lbs_project_backward: workspace too small");
  G.grad_cam = grad_cam; G.grad_px = grad_px; G.grad_view = grad_view; G.grad_verts = grad_verts;
  G.grad_skin = B > 1 ? grad_skin : nullptr;
  G.nblk = (V + kLbsThreads - 1) / kLbsThreads;
  G.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + 256);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(G.nblk, N);
  lbs_backward_kernel<<<grid, kLbsThreads, (size_t)B * 12 * sizeof(float), st>>>(G);
  if (lasr_check_launch("lbs_backward_kernel")) return 1;
  const long long tot = (long long)N * (B * 12 + 3);
  lbs_reduce_partials_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(G.partial, G.nblk, N, B, grad_R, grad_T,
                                                                            grad_pp, grad_fl, Rmat, rest, n_hypo, grad_rest);
  return lasr_check_launch("lbs_reduce_partials_kernel");
}

namespace lasr {
// Sums over frames of the per-frame gradients of SHARED parameters, one launch for the three of them (they were three
// PyTorch reductions per step): out_v [V,3] = sum_n gv[n], out_s [B-1,V] = sum_n gs[n], and -- the frames of a hypothesis,
// frames laid out [.., n_hypo] -- out_rest [n_hypo,B-1,3] = sum_m grest[m * n_hypo + h].  One thread per output element,
// frames added in order: deterministic, coalesced across threads.
__global__ void __launch_bounds__(256) lbs_sum_frames_kernel(const float* __restrict__ gv, const float* __restrict__ gs,
                                                             const float* __restrict__ grest, float* out_v, float* out_s,
                                                             float* out_rest, int N, int V, int B, int n_hypo) {
  const long long nv = gv ? 3LL * V : 0, ns = gs ? (long long)(B - 1) * V : 0;
  const long long nr = grest ? (long long)n_hypo * (B - 1) * 3 : 0;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float* src;
  float* dst;
  long long stride;
  int count;
  if (t < nv) { src = gv + t; dst = out_v + t; stride = nv; count = N; }
  else if (t < nv + ns) { src = gs + (t - nv); dst = out_s + (t - nv); stride = ns; count = N; }
  else if (t < nv + ns + nr) { src = grest + (t - nv - ns); dst = out_rest + (t - nv - ns); stride = nr; count = N / n_hypo; }
  else return;
  float acc = 0.f;
#pragma unroll 8
  for (int n = 0; n < count; ++n) acc += __ldg(src + (long long)n * stride);
  *dst = acc;
}

}  // namespace lasr

extern "C" int lasr_b200_lbs_sum_frames(const float* grad_verts, const float* grad_skin, const float* grad_rest,
                                        float* sum_verts, float* sum_skin, float* sum_rest, int N, int V, int B,
                                        int n_hypo, void* stream) {
  if (N <= 0 || V <= 0 || B < 1 || n_hypo < 1 || N % n_hypo != 0) return lasr_fail("lbs_sum_frames: bad sizes");
  if ((grad_verts == nullptr) != (sum_verts == nullptr) || (grad_skin == nullptr) != (sum_skin == nullptr) ||
      (grad_rest == nullptr) != (sum_rest == nullptr))
    return lasr_fail("lbs_sum_frames: every input needs its output and vice versa");
  if ((grad_skin || grad_rest) && B < 2) return lasr_fail("lbs_sum_frames: skin / rest gradients need at least one part");
  const long long tot = (grad_verts ? 3LL * V : 0) + (grad_skin ? (long long)(B - 1) * V : 0) +
                        (grad_rest ? (long long)n_hypo * (B - 1) * 3 : 0);
  if (tot == 0) return lasr_fail("lbs_sum_frames: nothing to sum");
  lbs_sum_frames_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      grad_verts, grad_skin, grad_rest, sum_verts, sum_skin, sum_rest, N, V, B, n_hypo);
  return lasr_check_launch("lbs_sum_frames_kernel");
}

extern "C" int lasr_b200_lbs_project_backward(const float* verts, const float* Rmat, const float* Tmat,
                                              const float* skin, const float* pp, const float* fl,
                                              const float* grad_cam, const float* grad_px, float* grad_verts,
                                              float* grad_R, float* grad_T, float* grad_skin, float* grad_pp,
                                              float* grad_fl, void* workspace, size_t workspace_bytes, int N,
                                              int V, int B, int verts_shared, int skin_shared, int tocam,
                                              void* stream) {
  return lbs_backward_impl(verts, Rmat, Tmat, nullptr, 1, skin, pp, fl, grad_cam, grad_px, nullptr, grad_verts, grad_R,
                           grad_T, nullptr, grad_skin, grad_pp, grad_fl, workspace, workspace_bytes, N, V, B, verts_shared,
                           skin_shared, tocam, stream);
}

extern "C" int lasr_b200_lbs_pose_backward(const float* verts, const float* Rmat, const float* Tmat,
                                           const float* rest_ts, const float* skin, const float* pp, const float* fl,
                                           const float* grad_cam, const float* grad_px, const float* grad_view,
                                           float* grad_verts, float* grad_R, float* grad_T, float* grad_rest,
                                           float* grad_skin, float* grad_pp, float* grad_fl, void* workspace,
                                           size_t workspace_bytes, int N, int V, int B, int n_hypo, int verts_shared,
                                           int skin_shared, int tocam, void* stream) {
  if (rest_ts && B < 2) return lasr_fail("lbs_pose_backward: rest_ts needs at least one part");
  return lbs_backward_impl(verts, Rmat, Tmat, rest_ts, n_hypo, skin, pp, fl, grad_cam, grad_px, grad_view, grad_verts,
                           grad_R, grad_T, grad_rest, grad_skin, grad_pp, grad_fl, workspace, workspace_bytes, N, V, B,
                           verts_shared, skin_shared, tocam, stream);
}
