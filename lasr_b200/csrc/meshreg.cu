// meshreg.cu -- mesh regularisers on the sparse topology, sm_100a.
//
// The reference evaluates them through dense V x V matrices:
//   LaplacianLoss  third_party/ext_nnutils/loss_utils.py:34-65   `L @ x` with a dense normalised Laplacian
//   ARAPLoss       nnutils/loss_utils.py:29-64                    six [N,V,V] tensors per call (2.2 GB each at
//                                                                 N=64, V=2930) to compare |x_b-x_a|^2 per edge
//   FlattenLoss    third_party/ext_nnutils/loss_utils.py:67-152  ~40 elementwise PyTorch kernels over [N,E,3]
// Here each is O(N * edges) over a CSR vertex adjacency (unique neighbours, ascending) or the edge list:
// one thread per (frame, vertex | edge), gradients in gather form (every output element has one owner,
// fixed summation order) -- no atomics, bit-reproducible.  HBM-bound integer/float gather work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "host_common.h"
#include "lasr_b200.h"

namespace lasr {

constexpr int kRegThreads = 256;

// block-wide sum of one float per thread, fixed order; thread 0 writes partial[blockIdx.y][blockIdx.x]
__device__ __forceinline__ void block_partial(float v, float* partial) {
  __shared__ float s_w[kRegThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRegThreads / 32; ++w) t += s_w[w];
    partial[(long long)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// loss[n] = scale * sum_b partial[n][b]  (block order)
__global__ void reg_reduce_kernel(const float* partial, int nblk, int N, float scale, float* loss) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(long long)n * nblk + b];
  loss[n] = t * scale;
}

__device__ __forceinline__ float3 ld3(const float* p) { return make_float3(p[0], p[1], p[2]); }

// ---------------------------------------------------------------------------------------------
// Laplacian: y_v = x_v - mean_{u in N(v)} x_u (0 for an isolated vertex), loss_n = sum |y|^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRegThreads) laplacian_forward_kernel(const float* __restrict__ x,
                                                                        const int* __restrict__ off,
                                                                        const int* __restrict__ nbr, float* lap,
                                                                        float* partial, int V) {
  const int n = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  float sq = 0.f;
  if (v < V) {
    const float* xb = x + (long long)n * V * 3;
    const int e0 = off[v], e1 = off[v + 1];
    float3 y = make_float3(0.f, 0.f, 0.f);
    if (e1 > e0) {
      float3 s = make_float3(0.f, 0.f, 0.f);
      for (int e = e0; e < e1; ++e) {
        const float3 u = ld3(xb + 3 * nbr[e]);
        s.x += u.x; s.y += u.y; s.z += u.z;
      }
      const float inv = 1.f / (float)(e1 - e0);
      const float3 c = ld3(xb + 3 * v);
      y = make_float3(c.x - s.x * inv, c.y - s.y * inv, c.z - s.z * inv);
    }
    float* o = lap + ((long long)n * V + v) * 3;
    o[0] = y.x; o[1] = y.y; o[2] = y.z;
    sq = y.x * y.x + y.y * y.y + y.z * y.z;
  }
  block_partial(sq, partial);
}

// grad_x = 2 g_n L^T y:  (L^T y)_v = [deg_v > 0] y_v - sum_{u in N(v)} y_u / deg_u   (adjacency symmetric)
__global__ void __launch_bounds__(kRegThreads) laplacian_backward_kernel(const float* __restrict__ lap,
                                                                         const float* __restrict__ gloss,
                                                                         const int* __restrict__ off,
                                                                         const int* __restrict__ nbr, float* gx,
                                                                         int V) {
  const int n = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* yb = lap + (long long)n * V * 3;
  const int e0 = off[v], e1 = off[v + 1];
  float3 a = e1 > e0 ? ld3(yb + 3 * v) : make_float3(0.f, 0.f, 0.f);
  for (int e = e0; e < e1; ++e) {
    const int u = nbr[e];
    const float inv = 1.f / (float)(off[u + 1] - off[u]);
    const float3 yu = ld3(yb + 3 * u);
    a.x -= yu.x * inv; a.y -= yu.y * inv; a.z -= yu.z * inv;
  }
  const float g = 2.f * gloss[n];
  float* o = gx + ((long long)n * V + v) * 3;
  o[0] = g * a.x; o[1] = g * a.y; o[2] = g * a.z;
}

// ---------------------------------------------------------------------------------------------
// ARAP: loss_n = mean over directed edges (a,b) of | |x_b - x_a|^2 - |dx_b - dx_a|^2 |
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float edge_q(const float3& xa, const float3& da, const float* xb, const float* db, int b) {
  const float3 xv = ld3(xb + 3 * b), dv = ld3(db + 3 * b);
  const float ex = xv.x - xa.x, ey = xv.y - xa.y, ez = xv.z - xa.z;
  const float fx = dv.x - da.x, fy = dv.y - da.y, fz = dv.z - da.z;
  return (ex * ex + ey * ey + ez * ez) - (fx * fx + fy * fy + fz * fz);
}

__global__ void __launch_bounds__(kRegThreads) arap_forward_kernel(const float* __restrict__ dx,
                                                                   const float* __restrict__ x,
                                                                   const int* __restrict__ off,
                                                                   const int* __restrict__ nbr, float* partial, int V) {
  const int n = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (a < V) {
    const float* xb = x + (long long)n * V * 3;
    const float* db = dx + (long long)n * V * 3;
    const float3 xa = ld3(xb + 3 * a), da = ld3(db + 3 * a);
    for (int e = off[a]; e < off[a + 1]; ++e) acc += fabsf(edge_q(xa, da, xb, db, nbr[e]));
  }
  block_partial(acc, partial);
}

// d/dx_a = (4 g_n / E2) sum_b sign(q_ab) (x_a - x_b); d/ddx_a = -(4 g_n / E2) sum_b sign(q_ab) (dx_a - dx_b)
// (each undirected edge is listed in both directions and q_ab = q_ba, hence the factor 4; sign(0) = 0)
__global__ void __launch_bounds__(kRegThreads) arap_backward_kernel(const float* __restrict__ dx,
                                                                    const float* __restrict__ x,
                                                                    const float* __restrict__ gloss,
                                                                    const int* __restrict__ off,
                                                                    const int* __restrict__ nbr, float* gdx, float* gx,
                                                                    int V, float inv_edges) {
  const int n = blockIdx.y;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= V) return;
  const float* xb = x + (long long)n * V * 3;
  const float* db = dx + (long long)n * V * 3;
  const float3 xa = ld3(xb + 3 * a), da = ld3(db + 3 * a);
  float3 sx = make_float3(0.f, 0.f, 0.f), sd = make_float3(0.f, 0.f, 0.f);
  for (int e = off[a]; e < off[a + 1]; ++e) {
    const int b = nbr[e];
    const float q = edge_q(xa, da, xb, db, b);
    const float s = q > 0.f ? 1.f : (q < 0.f ? -1.f : 0.f);
    const float3 xv = ld3(xb + 3 * b), dv = ld3(db + 3 * b);
    sx.x += s * (xa.x - xv.x); sx.y += s * (xa.y - xv.y); sx.z += s * (xa.z - xv.z);
    sd.x += s * (da.x - dv.x); sd.y += s * (da.y - dv.y); sd.z += s * (da.z - dv.z);
  }
  const float c = 4.f * gloss[n] * inv_edges;
  float* ox = gx + ((long long)n * V + a) * 3;
  float* od = gdx + ((long long)n * V + a) * 3;
  ox[0] = c * sx.x; ox[1] = c * sx.y; ox[2] = c * sx.z;
  od[0] = -c * sd.x; od[1] = -c * sd.y; od[2] = -c * sd.z;
}

// ---------------------------------------------------------------------------------------------
// Flatten: per edge (v0,v1) with the opposite vertices v2, v3 of its two faces:
//   a = p1-p0, b = p{2,3}-p0, cb = b - a (a.b)/(|a|^2+eps), |cb| ~ sqrt(|b|^2+eps) * sqrt(1-cos^2+eps)
//   cos = cb1.cb2 / (|cb1| |cb2| + eps),  loss_n = sum_e (cos + 1)^2      (ext loss_utils.py:110-152)
// ---------------------------------------------------------------------------------------------
struct FlatSide { float3 b, cb; float bl1, sin_, cos_, ab, ds, k; };

__device__ __forceinline__ float dot3f(const float3& u, const float3& v) { return u.x * v.x + u.y * v.y + u.z * v.z; }

__device__ __forceinline__ FlatSide flat_side(const float3& a, float al2, float al1, const float3& b, float eps) {
  FlatSide s;
  s.b = b;
  const float bl2 = dot3f(b, b);
  s.bl1 = sqrtf(bl2 + eps);
  s.ab = dot3f(a, b);
  s.ds = al1 * s.bl1 + eps;
  s.cos_ = s.ab / s.ds;
  s.sin_ = sqrtf(1.f - s.cos_ * s.cos_ + eps);
  s.k = s.ab / (al2 + eps);
  s.cb = make_float3(b.x - a.x * s.k, b.y - a.y * s.k, b.z - a.z * s.k);
  return s;
}

template <bool kBackward>
__global__ void __launch_bounds__(kRegThreads) flatten_kernel(const float* __restrict__ x, const int* __restrict__ edges,
                                                              const float* __restrict__ gloss, float* partial,
                                                              float* stage, int V, int E, float eps) {
  const int n = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (e < E) {
    const float* xb = x + (long long)n * V * 3;
    const int4 id = *reinterpret_cast<const int4*>(edges + 4 * e);
    const float3 p0 = ld3(xb + 3 * id.x), p1 = ld3(xb + 3 * id.y), p2 = ld3(xb + 3 * id.z), p3 = ld3(xb + 3 * id.w);
    const float3 a = make_float3(p1.x - p0.x, p1.y - p0.y, p1.z - p0.z);
    const float al2 = dot3f(a, a), al1 = sqrtf(al2 + eps);
    const FlatSide s1 = flat_side(a, al2, al1, make_float3(p2.x - p0.x, p2.y - p0.y, p2.z - p0.z), eps);
    const FlatSide s2 = flat_side(a, al2, al1, make_float3(p3.x - p0.x, p3.y - p0.y, p3.z - p0.z), eps);
    const float l1 = s1.bl1 * s1.sin_, l2 = s2.bl1 * s2.sin_;
    const float num = dot3f(s1.cb, s2.cb), D = l1 * l2 + eps;
    const float cosv = num / D;
    acc = (cosv + 1.f) * (cosv + 1.f);
    if (kBackward) {
      // reverse sweep of the expression above
      const float g_cos = 2.f * (cosv + 1.f) * gloss[n];
      const float g_num = g_cos / D, g_D = -g_cos * num / (D * D);
      float3 ga = make_float3(0.f, 0.f, 0.f);
      float3 gb[2];
      float g_al1 = 0.f, g_al2 = 0.f;
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const FlatSide& s = side == 0 ? s1 : s2;
        const FlatSide& t = side == 0 ? s2 : s1;
        const float l_other = side == 0 ? l2 : l1;
        const float3 g_cb = make_float3(g_num * t.cb.x, g_num * t.cb.y, g_num * t.cb.z);
        const float g_l = g_D * l_other;                 // d / d(|cb| of this side)
        float g_bl1 = g_l * s.sin_;
        const float g_sin = g_l * s.bl1;
        const float g_coss = g_sin * (-s.cos_ / s.sin_);
        float g_ab = g_coss / s.ds;
        const float g_ds = -g_coss * s.ab / (s.ds * s.ds);
        g_al1 += g_ds * s.bl1;
        g_bl1 += g_ds * al1;
        // cb = b - a k
        float3 g_b = g_cb;
        const float g_k = -(g_cb.x * a.x + g_cb.y * a.y + g_cb.z * a.z);
        ga.x -= g_cb.x * s.k; ga.y -= g_cb.y * s.k; ga.z -= g_cb.z * s.k;
        g_ab += g_k / (al2 + eps);
        g_al2 += -g_k * s.ab / ((al2 + eps) * (al2 + eps));
        const float g_bl2 = g_bl1 / (2.f * s.bl1);
        ga.x += g_ab * s.b.x; ga.y += g_ab * s.b.y; ga.z += g_ab * s.b.z;
        g_b.x += g_ab * a.x + 2.f * g_bl2 * s.b.x;
        g_b.y += g_ab * a.y + 2.f * g_bl2 * s.b.y;
        g_b.z += g_ab * a.z + 2.f * g_bl2 * s.b.z;
        gb[side] = g_b;
      }
      g_al2 += g_al1 / (2.f * al1);
      ga.x += 2.f * g_al2 * a.x; ga.y += 2.f * g_al2 * a.y; ga.z += 2.f * g_al2 * a.z;
      float* o = stage + ((long long)n * E + e) * 12;     // slots: p0 p1 p2 p3
      o[0] = -(ga.x + gb[0].x + gb[1].x); o[1] = -(ga.y + gb[0].y + gb[1].y); o[2] = -(ga.z + gb[0].z + gb[1].z);
      o[3] = ga.x; o[4] = ga.y; o[5] = ga.z;
      o[6] = gb[0].x; o[7] = gb[0].y; o[8] = gb[0].z;
      o[9] = gb[1].x; o[10] = gb[1].y; o[11] = gb[1].z;
    }
  }
  if (!kBackward) block_partial(acc, partial);
}

// per-edge corner gradients -> per-vertex gradients through the vertex -> (edge, slot) incidence CSR
__global__ void __launch_bounds__(kRegThreads) flatten_gather_kernel(const float* __restrict__ stage,
                                                                     const int* __restrict__ inc_off,
                                                                     const int* __restrict__ inc, float* gx, int V,
                                                                     int E) {
  const int n = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float* sb = stage + (long long)n * E * 12;
  float3 a = make_float3(0.f, 0.f, 0.f);
  for (int k = inc_off[v]; k < inc_off[v + 1]; ++k) {
    const float* g = sb + 3 * (long long)inc[k];        // entry = edge*4 + slot -> 3 floats at 3*entry
    a.x += g[0]; a.y += g[1]; a.z += g[2];
  }
  float* o = gx + ((long long)n * V + v) * 3;
  o[0] = a.x; o[1] = a.y; o[2] = a.z;
}

static int reg_args(const void* a, const void* b, const void* c, int N, int count, const char* what) {
  if (N <= 0 || count <= 0) return lasr_fail(std::string(what) + ": sizes must be positive");
  if (N > 65535) return lasr_fail(std::string(what) + ": more than 65535 frames per call unsupported");
  if (!a || !b || !c) return lasr_fail(std::string(what) + ": null tensor pointer");
  return 0;
}

static int nblk_of(int count) { return (count + kRegThreads - 1) / kRegThreads; }

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_meshreg_workspace_bytes(int N, int V, int E) {
  if (N <= 0 || V <= 0 || E < 0) return 0;
  const size_t cnt = (size_t)(V > E ? V : E);
  return (size_t)N * ((cnt + kRegThreads - 1) / kRegThreads) * sizeof(float) + (size_t)N * E * 12 * sizeof(float) + 256;
}

extern "C" int lasr_b200_laplacian_forward(const float* x, const int32_t* nbr_offsets, const int32_t* nbr,
                                           float* lap, float* loss, void* workspace, size_t workspace_bytes, int N,
                                           int V, void* stream) {
  if (reg_args(x, nbr_offsets, nbr, N, V, "laplacian_forward") || !lap || !loss)
    return (lap && loss) ? 1 : lasr_fail("laplacian_forward: null output");
  if (!workspace || workspace_bytes < lasr_b200_meshreg_workspace_bytes(N, V, 0))
    return lasr_fail("laplacian_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  const int nblk = nblk_of(V);
  laplacian_forward_kernel<<<dim3(nblk, N), kRegThreads, 0, st>>>(x, nbr_offsets, nbr, lap, partial, V);
  if (lasr_check_launch("laplacian_forward_kernel")) return 1;
  reg_reduce_kernel<<<(N + 127) / 128, 128, 0, st>>>(partial, nblk, N, 1.f, loss);
  return lasr_check_launch("reg_reduce_kernel");
}

extern "C" int lasr_b200_laplacian_backward(const float* lap, const float* grad_loss, const int32_t* nbr_offsets,
                                            const int32_t* nbr, float* grad_x, int N, int V, void* stream) {
  if (reg_args(lap, nbr_offsets, nbr, N, V, "laplacian_backward") || !grad_loss || !grad_x)
    return (grad_loss && grad_x) ? 1 : lasr_fail("laplacian_backward: null pointer");
  laplacian_backward_kernel<<<dim3(nblk_of(V), N), kRegThreads, 0, (cudaStream_t)stream>>>(lap, grad_loss, nbr_offsets,
                                                                                            nbr, grad_x, V);
  return lasr_check_launch("laplacian_backward_kernel");
}

extern "C" int lasr_b200_arap_forward(const float* dx, const float* x, const int32_t* nbr_offsets, const int32_t* nbr,
                                      float* loss, void* workspace, size_t workspace_bytes, int N, int V,
                                      int num_directed_edges, void* stream) {
  if (reg_args(dx, x, nbr_offsets, N, V, "arap_forward") || !nbr || !loss)
    return (nbr && loss) ? 1 : lasr_fail("arap_forward: null pointer");
  if (num_directed_edges <= 0) return lasr_fail("arap_forward: the mesh has no edges");
  if (!workspace || workspace_bytes < lasr_b200_meshreg_workspace_bytes(N, V, 0))
    return lasr_fail("arap_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  const int nblk = nblk_of(V);
  arap_forward_kernel<<<dim3(nblk, N), kRegThreads, 0, st>>>(dx, x, nbr_offsets, nbr, partial, V);
  if (lasr_check_launch("arap_forward_kernel")) return 1;
  reg_reduce_kernel<<<(N + 127) / 128, 128, 0, st>>>(partial, nblk, N, 1.f / (float)num_directed_edges, loss);
  return lasr_check_launch("reg_reduce_kernel");
}

extern "C" int lasr_b200_arap_backward(const float* dx, const float* x, const float* grad_loss,
                                       const int32_t* nbr_offsets, const int32_t* nbr, float* grad_dx, float* grad_x,
                                       int N, int V, int num_directed_edges, void* stream) {
  if (reg_args(dx, x, nbr_offsets, N, V, "arap_backward") || !nbr || !grad_loss || !grad_dx || !grad_x)
    return (nbr && grad_loss && grad_dx && grad_x) ? 1 : lasr_fail("arap_backward: null pointer");
  if (num_directed_edges <= 0) return lasr_fail("arap_backward: the mesh has no edges");
  arap_backward_kernel<<<dim3(nblk_of(V), N), kRegThreads, 0, (cudaStream_t)stream>>>(
      dx, x, grad_loss, nbr_offsets, nbr, grad_dx, grad_x, V, 1.f / (float)num_directed_edges);
  return lasr_check_launch("arap_backward_kernel");
}

extern "C" int lasr_b200_flatten_forward(const float* x, const int32_t* edges, float* loss, void* workspace,
                                         size_t workspace_bytes, int N, int V, int E, float eps, void* stream) {
  if (reg_args(x, edges, loss, N, E, "flatten_forward") || V <= 0) return 1;
  if (!workspace || workspace_bytes < lasr_b200_meshreg_workspace_bytes(N, V, E))
    return lasr_fail("flatten_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>(workspace);
  const int nblk = nblk_of(E);
  flatten_kernel<false><<<dim3(nblk, N), kRegThreads, 0, st>>>(x, edges, nullptr, partial, nullptr, V, E, eps);
  if (lasr_check_launch("flatten_kernel")) return 1;
  reg_reduce_kernel<<<(N + 127) / 128, 128, 0, st>>>(partial, nblk, N, 1.f, loss);
  return lasr_check_launch("reg_reduce_kernel");
}

extern "C" int lasr_b200_flatten_backward(const float* x, const int32_t* edges, const float* grad_loss,
                                          const int32_t* inc_offsets, const int32_t* inc, float* grad_x,
                                          void* workspace, size_t workspace_bytes, int N, int V, int E, float eps,
                                          void* stream) {
  if (reg_args(x, edges, grad_loss, N, E, "flatten_backward") || V <= 0) return 1;
  if (!inc_offsets || !inc || !grad_x) return lasr_fail("flatten_backward: null pointer");
  if (!workspace || workspace_bytes < lasr_b200_meshreg_workspace_bytes(N, V, E))
    return lasr_fail("flatten_backward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cnt = (size_t)(V > E ? V : E);
  float* stage = reinterpret_cast<float*>(workspace) + (size_t)N * ((cnt + kRegThreads - 1) / kRegThreads) + 32;
  flatten_kernel<true><<<dim3(nblk_of(E), N), kRegThreads, 0, st>>>(x, edges, grad_loss, nullptr, stage, V, E, eps);
  if (lasr_check_launch("flatten_kernel<backward>")) return 1;
  flatten_gather_kernel<<<dim3(nblk_of(V), N), kRegThreads, 0, st>>>(stage, inc_offsets, inc, grad_x, V, E);
  return lasr_check_launch("flatten_gather_kernel");
}
