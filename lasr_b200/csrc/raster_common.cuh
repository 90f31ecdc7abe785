// raster_common.cuh -- per-face records and per-(pixel,face) evaluation shared by the forward
// (tile-binned, pixel-parallel) and backward (face-parallel) soft-rasterizer kernels.
//
// What is computed is the reference's SoftRas model
// (third_party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu:24-216, 245-305); how it
// is computed is ours.  Two numerical design rules:
//
//  (1) The geometric predicates (face_inv / barycentrics / inside-outside / edge selection) are
//      ill-conditioned for edge-on sliver faces: a 1-ulp change in `det` flips a pixel between
//      "inside" (D ~ 1) and "outside" (D ~ 0).  To agree with the reference *as compiled by
//      nvcc* (FMA contraction on), those expressions use explicit round-to-nearest intrinsics in
//      exactly the association nvcc gives the reference's source
//      (a*b + c*d -> fma(a, b, c*d); a*b + c*d + e*f -> fma(e, f, fma(a, b, c*d));
//       a*b - c*d -> fma(a, b, -(c*d)); checked against `cuobjdump -sass` of oracle/_ref), and
//      IEEE division for face_inv.
//  (2) Everything downstream (closest-point parameter, sigmoid, clipped-barycentric depth, depth
//      softmax, colour blend and their derivatives) is well conditioned, so it uses reciprocals
//      (MUFU.RCP) and ex2.approx instead of the reference's IEEE divisions / expf; measured
//      deviation from the reference kernels is ~1e-6 (tests/golden/report.json), two orders
//      below the 1e-4 / 1e-3 tolerances.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lasr {

#ifndef LASR_NOISE_C1
#define LASR_NOISE_C1 8.f
#endif
#ifndef LASR_NOISE_C2
#define LASR_NOISE_C2 8.f
#endif
#ifndef LASR_NOISE_K
#define LASR_NOISE_K 4.f
#endif

constexpr int kMaxSets = 3;    // colour sets interpolated in one geometry pass (indexed path)
// floats per face record: 32 of geometry, then 9 per colour set (+ z for the reference-faithful
// generic path, which is single-set), rounded to float4s: 44 / 52 / 60 for 1 / 2 / 3 sets
__host__ __device__ constexpr int rec_words(int K) { return K == 1 ? 44 : ((32 + 9 * K + 3) / 4) * 4; }
constexpr int kCompactWords = 20;   // inverse 9 | 1/z 3 | 1/edge^2 3 | flags | kappa | 3 spare  (80 bytes, 16-byte aligned)

struct RasterArgs {
  const float* faces;        // [N,F,9]
  const float* textures;     // [N,F,T,3]
  float* aggrs_info;         // [N,2,S,S]
  float* soft_colors;        // [N,4,S,S]
  const float* grad_soft_colors;  // [N,4,S,S]   (backward)
  const float4* packed;           // [N][1 + ceil(3*live/4)][S*S] packed per-pixel gradients (backward of the fused region)
  float4* pack_area;              // workspace room for the single-set backward's own packing pass (raster_pack_kernel)
  float* grad_faces;         // [N,F,9]          (backward)
  float* grad_textures;      // [N,F,T,3]        (backward)
  // workspace carve-up
  uint2* face_window;        // [N*F] packed pixel windows
  float4* face_border;       // [N*F] dilated bounding boxes (lox, hix, loy, hiy)
  float* face_record;        // [N*F][rec_words(K)]
  float* compact_record;     // [N*F][kCompactWords] or null: the division results of the record only -- what the backward
                             //   keeps of a face when the caller saves 80 bytes per face instead of the whole workspace
  float* coords;             // [S] pixel-centre coordinates
  int records_valid;         // backward: workspace already holds the forward's setup of the same inputs
  int* big_list;             // [1 + cap]: count, then flat (frame*F + face) ids of faces deferred to the CTA-per-face pass
  int* frame_window;         // [N][4] (x_lo, -x_hi, y_lo, -y_hi) of the union of face windows (atomicMin)
  int* tile_count;           // [N][tiles] faces whose window reaches each forward tile (work estimate)
  int* tile_order;           // [N*tiles] (frame * tiles + tile), heaviest first: the forward's dispatch order
  const float* near_far_dev; // optional {near, far} on device
  // indexed ("fused") input mode: faces/textures are gathered on the fly from per-vertex arrays
  const float* verts;        // [N,V,3] clip-space vertices   (null: direct mode, use `faces`)
  const float* set_ptr[kMaxSets];  // colour set s: [N,V,3] vertex colours; frame n reads frame (n + set_shift[s]) mod N of it
  int set_shift[kMaxSets];         //   (one stacked [K][N,V,3] tensor is the special case ptr + s*N*V*3, shift 0)
  int K;                     // number of colour sets (1 on the reference-signature paths); images have 3K+1 planes
  int first_live_set;        // backward: colour sets below this index are detached (no gradient read / written)
  int bg_in_args;            // forward: background colours come from bg[] instead of a pre-filled soft_colors
  float bg[3 * kMaxSets];
  const int* tri;            // [F,3] shared topology
  int V;
  int N, F, S, T, R;
  float near_, far_, eps, sigma, dist_eps, gamma;
  float inv_sigma, inv_gamma;
  int dist, rgb, alpha, sample, double_side;
};

// Fetch the 9 vertex floats (and, for 'vertex' textures, the 9 corner colours) of face f of frame
// n, either from the reference's gathered tensors or straight from per-vertex arrays + topology.
// (tex9 receives 9 floats per colour set, A.K sets, when want_tex.)
__device__ __forceinline__ void fetch_face(const RasterArgs& A, int n, int f, float* fv, float* tex9, bool want_tex) {
  if (A.tri) {
    const int i0 = A.tri[3 * f], i1 = A.tri[3 * f + 1], i2 = A.tri[3 * f + 2];
    const float* vb = A.verts + (long long)n * A.V * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) { fv[k] = vb[3 * i0 + k]; fv[3 + k] = vb[3 * i1 + k]; fv[6 + k] = vb[3 * i2 + k]; }
    if (want_tex) {
#pragma unroll
      for (int s = 0; s < kMaxSets; ++s) {
        if (s < A.K) {
          int m = n + A.set_shift[s];
          if (m >= A.N) m -= A.N;
          const float* cb = A.set_ptr[s] + (long long)m * A.V * 3;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            tex9[9 * s + k] = cb[3 * i0 + k]; tex9[9 * s + 3 + k] = cb[3 * i1 + k]; tex9[9 * s + 6 + k] = cb[3 * i2 + k];
          }
        }
      }
    }
  } else {
    const long long i = (long long)n * A.F + f;
#pragma unroll
    for (int k = 0; k < 9; ++k) fv[k] = A.faces[i * 9 + k];
    if (want_tex) {
#pragma unroll
      for (int k = 0; k < 9; ++k) tex9[k] = A.textures[i * 9 + k];
    }
  }
}

// ---- reference-shaped sums of products (rule 1) ----------------------------------------------
__device__ __forceinline__ float dot2(float a, float b, float c, float d) {
  return __fmaf_rn(a, b, __fmul_rn(c, d));
}
__device__ __forceinline__ float dot3(float a, float b, float c, float d, float e, float f) {
  return __fmaf_rn(e, f, __fmaf_rn(a, b, __fmul_rn(c, d)));
}
__device__ __forceinline__ float cross2(float a, float b, float c, float d) {
  return __fmaf_rn(a, b, -__fmul_rn(c, d));
}
__device__ __forceinline__ float min3(float a, float b, float c) { return fminf(fminf(a, b), c); }
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
// MUFU.RCP / MUFU.EX2 without the denormal-range scaling nvcc wraps around __fdividef / __expf when it
// compiles without -ftz (FSETP + FMUL 2^24 + FSEL ... : five instructions for one reciprocal).  Every
// operand on the kFast path is far from the denormal range (1 + e^x, sums of weights >= 1e-5, softmax sums
// >= e^{eps/gamma}), and an underflowing e^x flushing to 0 instead of 1e-40 is below anything the images
// or gradients resolve.  Non-volatile asm: the compiler may schedule, hoist and CSE it.
//
// The arithmetic mode of the per-pair functions is a template int: 0 = generic (the reference's IEEE
// divisions and expf), 1 = fast (reciprocals / ex2 through __fdividef / __expf), 2 = fast with the bare
// FTZ MUFU forms.  Measured on B200 (profiles/RESULTS.md): mode 2 is 4-7 % faster than mode 1 in the forward
// kernel and 13 % SLOWER in the backward (same instruction savings, worse schedule at 4 warps per
// scheduler), so each kernel picks its own.
template <int kMode> __device__ __forceinline__ float fast_rcp(float x) {
#if defined(__CUDA_ARCH__)
  if (kMode == 2) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
#endif
  return __fdividef(1.f, x);
}
template <int kMode> __device__ __forceinline__ float fast_exp(float x) {
#if defined(__CUDA_ARCH__)
  if (kMode == 2) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
  }
#endif
  return __expf(x);
}

// Rule (2) applies to LASR's training configuration (kFast: euclidean / softmax / prod / vertex
// colours), which has no hard decision downstream of the distance stage.  The other operator
// modes do ('surface' texel selection sits exactly on `frac sum <= 1` along clipped edges, the
// 'hard' z-test and 'hard' alpha compare rounded values), so the generic instantiation keeps
// the reference's IEEE divisions and expf.
template <int kFast> __device__ __forceinline__ float div_(float a, float b) {
  return kFast ? a * fast_rcp<kFast>(b) : __fdiv_rn(a, b);
}
template <int kFast> __device__ __forceinline__ float exp_(float x) { return kFast ? fast_exp<kFast>(x) : expf(x); }

// Correctly rounded num/den from a correctly rounded reciprocal rd = RN(1/den): one Newton step on
// the quotient (Markstein).  Three FP32 instructions instead of the ~10-instruction IEEE division
// sequence, same result.  The closest-point parameter feeds the squared distance that is compared
// against the hard cut-off (`dis >= threshold`, .cu:402), so it must round like the reference.
__device__ __forceinline__ float div_by_rcp(float num, float den, float rd) {
  const float q0 = __fmul_rn(num, rd);
  const float r = __fmaf_rn(-den, q0, num);
  const float q1 = __fmaf_rn(r, rd, q0);
  // foreshortened (edge-on) edges have den == 0 after rounding -> rd = inf, r = NaN; huge rd can
  // overflow q0.  Those rare cases take the IEEE division so inf/NaN propagate as in the reference.
  return (fabsf(rd) < 1e30f && fabsf(q0) < 1e30f) ? q1 : __fdiv_rn(num, den);
}

// Pixel-centre coordinate, (2*i + 1 - S) / S evaluated in double and rounded once (.cu:345-346).
// Only the setup kernel calls this (S values per launch); everyone else reads the table.
__device__ __forceinline__ float pixel_coord(int i, int S) {
  return (float)((2.0 * (double)i + 1.0 - (double)S) / (double)S);
}

// ---- per-face derived quantities (.cu:245-305 + the border box of :32-38) -------------------
struct FaceGeom {
  float inv[9];       // inverse of [[x0 x1 x2],[y0 y1 y2],[1 1 1]] (rows give barycentrics)
  float x[3], y[3], z[3];
  float iz[3];        // 1 / z
  float a[3][3];      // a[e][j] = sym[e][j] - sym[(e+1)%3][j],  sym = P P^T + 1
  float rden[3];      // 1 / (a[e][e] - a[e][(e+1)%3])  (= 1 / squared edge length)
  float lox, hix, loy, hiy;  // bounding box dilated by sqrt(threshold)
  float det;          // clamped determinant the rows of inv were divided by
  float kappa;        // edge-prediction trust threshold (see euclid_distance): (s_mid - s_lo)^2 > kappa * s_mid
  int obt;            // index of the (first) obtuse corner, -1 if none
  bool front;
};

// a[e][j] from the vertices: sym = F F^T with z -> 1 (.cu:288-294); only row differences are ever used (:81-83,132-134)
__device__ __forceinline__ void face_edge_rows(FaceGeom& g) {
  float sym[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) sym[j][k] = __fadd_rn(dot2(g.x[j], g.x[k], g.y[j], g.y[k]), 1.0f);
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.a[e][j] = __fsub_rn(sym[e][j], sym[(e + 1) % 3][j]);
}

// bounding box dilated by sqrt(threshold) (.cu:32-38)
__device__ __forceinline__ void face_box(FaceGeom& g, float sthr) {
  g.hix = __fadd_rn(max3(g.x[0], g.x[1], g.x[2]), sthr);
  g.lox = __fsub_rn(min3(g.x[0], g.x[1], g.x[2]), sthr);
  g.hiy = __fadd_rn(max3(g.y[0], g.y[1], g.y[2]), sthr);
  g.loy = __fsub_rn(min3(g.y[0], g.y[1], g.y[2]), sthr);
}

__device__ __forceinline__ void face_setup(const float* __restrict__ fv, float sthr, FaceGeom& g) {
  const float x0 = fv[0], y0 = fv[1], x1 = fv[3], y1 = fv[4], x2 = fv[6], y2 = fv[7];
  g.x[0] = x0; g.x[1] = x1; g.x[2] = x2;
  g.y[0] = y0; g.y[1] = y1; g.y[2] = y2;
  g.z[0] = fv[2]; g.z[1] = fv[5]; g.z[2] = fv[8];
  g.iz[0] = __fdiv_rn(1.f, fv[2]); g.iz[1] = __fdiv_rn(1.f, fv[5]); g.iz[2] = __fdiv_rn(1.f, fv[8]);
  // determinant: x2*(y0-y1) + x0*(y1-y2) + x1*(y2-y0)   (.cu:278-281)
  float det = __fmaf_rn(x1, __fsub_rn(y2, y0), __fmaf_rn(x2, __fsub_rn(y0, y1), __fmul_rn(x0, __fsub_rn(y1, y2))));
  det = det > 0.f ? fmaxf(det, 1e-10f) : fminf(det, -1e-10f);  // :282
  g.det = det;
  const float adj[9] = {__fsub_rn(y1, y2), __fsub_rn(x2, x1), cross2(x1, y2, x2, y1),
                        __fsub_rn(y2, y0), __fsub_rn(x0, x2), cross2(x2, y0, x0, y2),
                        __fsub_rn(y0, y1), __fsub_rn(x1, x0), cross2(x0, y1, x1, y0)};
#pragma unroll
  for (int k = 0; k < 9; ++k) g.inv[k] = __fdiv_rn(adj[k], det);  // :284-286 (IEEE division)
  face_edge_rows(g);
#pragma unroll
  for (int e = 0; e < 3; ++e) g.rden[e] = __fdiv_rn(1.f, __fsub_rn(g.a[e][e], g.a[e][(e + 1) % 3]));
  // first obtuse corner (:296-304)
  g.obt = -1;
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    const float d = dot2(__fsub_rn(g.x[k1], g.x[k]), __fsub_rn(g.x[k2], g.x[k]),
                         __fsub_rn(g.y[k1], g.y[k]), __fsub_rn(g.y[k2], g.y[k]));
    if (d < 0.f) g.obt = k;  // descending k so the smallest obtuse k wins, as `break` does
  }
  g.front = (y2 - y0) * (x1 - x0) < (y1 - y0) * (x2 - x0);  // :41-44
  {
    // Rounding-noise amplitude (in NDC) of the distances the reference's float pipeline computes for
    // this face: barycentric noise 2^-24 * |inv row| carried onto the vertices, plus the along-edge
    // wobble of the foot point, ~2^-24 * |sym| / |shortest edge|.  The edge *prediction* for inside
    // pixels is trusted only when the two nearest edges are separated by kNoiseK such amplitudes:
    //   d_mid - d_lo > K eta   <=   (s_mid - s_lo)^2 * det^2 > 4 K^2 eta^2 * s_mid      (d_e^2 = s_e det^2)
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) m = fmaxf(m, fabsf(g.inv[3 * k]) + fabsf(g.inv[3 * k + 1]) + fabsf(g.inv[3 * k + 2]));
    const float xm = fmaxf(1.f, fmaxf(max3(fabsf(x0), fabsf(x1), fabsf(x2)), max3(fabsf(y0), fabsf(y1), fabsf(y2))));
    const float emin2 = min3(__fsub_rn(g.a[0][0], g.a[0][1]), __fsub_rn(g.a[1][1], g.a[1][2]), __fsub_rn(g.a[2][2], g.a[2][0]));
    const float eta = 5.96e-8f * (LASR_NOISE_C1 * m * xm + LASR_NOISE_C2 * xm * xm * rsqrtf(fmaxf(emin2, 1e-30f)));
    const float ke = 2.f * LASR_NOISE_K * eta / det;
    g.kappa = ke * ke;   // inf / NaN for degenerate faces: the prediction is then never trusted
  }
  face_box(g, sthr);
}

// The face as the backward needs it from the vertices plus the COMPACT record (store_compact): everything that costs a
// division comes from the record, the rest -- the edge rows and the box: sums, differences, min / max -- is re-derived
// by the same expressions face_setup uses, so every field has the bits the forward worked with.
// (the record is loaded by its own call so that the kernel can issue these loads BEFORE it fetches the vertices: placed
// after them, the compiler scheduled the box computation -- a wait for the vertices -- ahead of the record's loads, two
// memory round trips in a row per face: trip 23's ncu capture)
__device__ __forceinline__ void load_compact(const float* __restrict__ r, float (&v)[kCompactWords]) {
  const float4* r4 = reinterpret_cast<const float4*>(r);
#pragma unroll
  for (int k = 0; k < kCompactWords / 4; ++k) {
    const float4 q = __ldg(r4 + k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
}

__device__ __forceinline__ void face_from_compact(const float* __restrict__ fv, float sthr, const float (&v)[kCompactWords],
                                                  FaceGeom& g) {
  g.x[0] = fv[0]; g.x[1] = fv[3]; g.x[2] = fv[6];
  g.y[0] = fv[1]; g.y[1] = fv[4]; g.y[2] = fv[7];
  g.z[0] = fv[2]; g.z[1] = fv[5]; g.z[2] = fv[8];
#pragma unroll
  for (int k = 0; k < 9; ++k) g.inv[k] = v[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) { g.iz[k] = v[9 + k]; g.rden[k] = v[12 + k]; }
  const int fl = __float_as_int(v[15]);
  g.obt = (fl & 3) - 1;
  g.front = (fl & 4) != 0;
  g.kappa = v[16];
  g.det = 0.f;       // used by face_setup only
  face_edge_rows(g);
  face_box(g, sthr);
}

__device__ __forceinline__ void store_compact(float* __restrict__ r, const FaceGeom& g) {
  float v[kCompactWords];
#pragma unroll
  for (int k = 0; k < kCompactWords; ++k) v[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) v[k] = g.inv[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) { v[9 + k] = g.iz[k]; v[12 + k] = g.rden[k]; }
  v[15] = __int_as_float((g.obt + 1) | (g.front ? 4 : 0));
  v[16] = g.kappa;
  float4* r4 = reinterpret_cast<float4*>(r);
#pragma unroll
  for (int k = 0; k < kCompactWords / 4; ++k) r4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// record layout (float index):
//  0..8 inv | 9..11 x | 12..14 y | 15 flags | 16..18 iz | 19..21 rden | 22..30 a[e][j] | 31 kappa
//  32..40 vertex-colour texture | 41..43 z
template <int K>
__device__ __forceinline__ void store_record(float* __restrict__ r, const FaceGeom& g, const float* tex9) {
  constexpr int kRecWords = rec_words(K);
  float v[kRecWords];
#pragma unroll
  for (int k = 0; k < kRecWords; ++k) v[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) v[k] = g.inv[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    v[9 + k] = g.x[k]; v[12 + k] = g.y[k]; v[16 + k] = g.iz[k]; v[19 + k] = g.rden[k];
    if (K == 1) v[41 + k] = g.z[k];
  }
  v[15] = __int_as_float((g.obt + 1) | (g.front ? 4 : 0));
  v[31] = g.kappa;
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[22 + 3 * e + j] = g.a[e][j];
  if (tex9) {
#pragma unroll
    for (int k = 0; k < 9 * K; ++k) v[32 + k] = tex9[k];
  }
  float4* r4 = reinterpret_cast<float4*>(r);
#pragma unroll
  for (int k = 0; k < kRecWords / 4; ++k) r4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// Uniform (warp-broadcast) 128-bit loads of a record: every lane reads the same address.
template <int K>
__device__ __forceinline__ void load_record(const float* __restrict__ r, FaceGeom& g, float* tex9) {
  constexpr int kRecWords = rec_words(K);
  const float4* r4 = reinterpret_cast<const float4*>(r);
  float v[kRecWords];
#pragma unroll
  for (int k = 0; k < kRecWords / 4; ++k) {
    const float4 q = __ldg(r4 + k);
    v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) g.inv[k] = v[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g.x[k] = v[9 + k]; g.y[k] = v[12 + k]; g.iz[k] = v[16 + k]; g.rden[k] = v[19 + k];
    g.z[k] = K == 1 ? v[41 + k] : 0.f;
  }
  g.kappa = v[31];
  const int fl = __float_as_int(v[15]);
  g.obt = (fl & 3) - 1;
  g.front = (fl & 4) != 0;
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.a[e][j] = v[22 + 3 * e + j];
#pragma unroll
  for (int k = 0; k < 9 * K; ++k) tex9[k] = v[32 + k];
}

// The same from a record staged in shared memory (the bulk-copy forward, raster.cu): 32-bit shared-space address, explicit
// ld.shared.v4 -- every lane reads the same address, a broadcast, one wavefront per load.
#ifdef __CUDACC__
template <int K>
__device__ __forceinline__ void load_record_shared(uint32_t saddr, FaceGeom& g, float* tex9) {
  constexpr int kRecWords = rec_words(K);
  float v[kRecWords];
#pragma unroll
  for (int k = 0; k < kRecWords / 4; ++k)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * k]), "=f"(v[4 * k + 1]), "=f"(v[4 * k + 2]), "=f"(v[4 * k + 3]) : "r"(saddr + 16u * k));
#pragma unroll
  for (int k = 0; k < 9; ++k) g.inv[k] = v[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g.x[k] = v[9 + k]; g.y[k] = v[12 + k]; g.iz[k] = v[16 + k]; g.rden[k] = v[19 + k];
    g.z[k] = K == 1 ? v[41 + k] : 0.f;
  }
  g.kappa = v[31];
  const int fl = __float_as_int(v[15]);
  g.obt = (fl & 3) - 1;
  g.front = (fl & 4) != 0;
#pragma unroll
  for (int e = 0; e < 3; ++e)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.a[e][j] = v[22 + 3 * e + j];
#pragma unroll
  for (int k = 0; k < 9 * K; ++k) tex9[k] = v[32 + k];
}
#endif

// Inclusive pixel-index window of the pixel centres inside [lo, hi] (centre i sits at (2i + 1 - S) / S, so coordinate c
// has index (c*S + S - 1) / 2).  The exact predicate is still applied per pixel; this only has to be a superset, and it
// is a tight one: the bounds are widened by 1/64 pixel, which covers the float rounding of this conversion and of the
// coordinate table for every S <= 8192 (<= 1e-3 pixel; tests/test_kernel_math_host.py sweeps it).  Round 1 rounded
// outwards by a whole pixel on both sides -- a 12.8-pixel box became a ~14.8-pixel window, and a quarter of the
// backward's sweep slots and of the forward's tile-list entries fell outside the box (LASR_LOOSE_WINDOWS=1 restores it).
// Empty windows come back lo > hi.  Boxes reaching beyond the screen -- including +-inf and the huge coordinates a
// pinhole projection produces for z ~ 0 -- are clamped to it, as the reference's per-pixel test would accept them; only
// NaN bounds are culled (the reference's behaviour for NaN vertices is undefined: out-of-bounds stack writes in
// euclidean_p2f_distance when every comparison is false).
#ifndef LASR_LOOSE_WINDOWS
#define LASR_LOOSE_WINDOWS 0
#endif
__device__ __forceinline__ void coord_window(float lo, float hi, int S, int& ilo, int& ihi) {
  float a = (lo * (float)S + (float)(S - 1)) * 0.5f;
  float b = (hi * (float)S + (float)(S - 1)) * 0.5f;
  if (!(a <= b + 1.f)) { ilo = 1; ihi = 0; return; }          // NaN (or inverted) bounds
  a = fminf(fmaxf(a, -2.f), (float)S + 1.f);                   // beyond the screen: clamp before the int conversion
  b = fminf(fmaxf(b, -2.f), (float)S + 1.f);
#if LASR_LOOSE_WINDOWS
  ilo = max(0, (int)floorf(a - 0.0625f));
  ihi = min(S - 1, (int)ceilf(b + 0.0625f));
#else
  ilo = max(0, (int)ceilf(a - 0.015625f));
  ihi = min(S - 1, (int)floorf(b + 0.015625f));
#endif
}

// ---- the soft fragment of one (pixel, face) pair --------------------------------------------
struct Fragment {
  float w[3];        // barycentrics (unclipped)
  float t[3];        // euclidean: closest point minus w;  barycentric: w
  float sign, dx, dy, dis;
  float frag;        // D in the paper
};

__device__ __forceinline__ bool pixel_inside(const float* w) {  // .cu:47-50
  return w[0] <= 1.f && w[0] >= 0.f && w[1] <= 1.f && w[1] >= 0.f && w[2] <= 1.f && w[2] >= 0.f;
}

// .cu:61-151
//
// The reference evaluates all three edges for a pixel strictly inside the face and keeps the
// nearest (first minimum), and one edge -- chosen from the signs of w -- for a pixel outside.  A
// warp that holds both kinds of pixel would execute both code paths, the three-edge one with two or
// three live lanes.  Here (kFast, LASR_FAST_INSIDE) an inside pixel *predicts* the winning edge
// from  dist(p, line e)^2 = w_opp(e)^2 * det^2 / |edge e|^2  (det^2 is common to the three, so
// w_opp^2 * rden[e] ranks them) and then runs the same single-edge evaluation as an outside pixel,
// with the reference's own rounding; when the two best candidates are within 1e-3 of each other
// (or anything is NaN) the lane falls back to the three-edge loop, so the edge -- and with it every
// bit of (t, dx, dy, dis) -- is the one the reference's loop selects.
#ifndef LASR_FAST_INSIDE
#define LASR_FAST_INSIDE 1
#endif
template <int kFast>
__device__ __forceinline__ void euclid_distance(const FaceGeom& g, float xp, float yp, Fragment& fr) {
  const float w0 = fr.w[0], w1 = fr.w[1], w2 = fr.w[2];
  const bool inside = w0 > 0.f && w1 > 0.f && w2 > 0.f && w0 < 1.f && w1 < 1.f && w2 < 1.f;
  int e = 0;
  bool all_edges = inside;
  if (kFast != 0 && LASR_FAST_INSIDE) {
    const float s0 = w2 * w2 * g.rden[0], s1 = w0 * w0 * g.rden[1], s2 = w1 * w1 * g.rden[2];
    const float lo = fminf(fminf(s0, s1), s2);
    const float mid = fmaxf(fminf(s0, s1), fminf(fmaxf(s0, s1), s2));   // second smallest
    e = s0 == lo ? 0 : (s1 == lo ? 1 : 2);
    const float gap = mid - lo;
    // near tie relative to the face's rounding noise, or NaN/inf from a degenerate edge -> all edges
    all_edges = inside && !(gap > 1e-3f * mid && gap * gap > g.kappa * mid);
  }
#ifdef LASR_HOST_STATS   // tests/host_harness only
  if (inside) ++lasr_host_inside_pairs;
  if (all_edges) ++lasr_host_all_edges_pairs;
#endif
  if (all_edges) {
    float best = 100000000.f, bx = 0.f, by = 0.f;
#pragma unroll
    for (int e3 = 0; e3 < 3; ++e3) {
      const int b = (e3 + 1) % 3, c = (e3 + 2) % 3;
      const float num = __fsub_rn(dot3(w0, g.a[e3][0], w1, g.a[e3][1], w2, g.a[e3][2]), g.a[e3][b]);
      const float den = __fsub_rn(g.a[e3][e3], g.a[e3][b]);
      const float q = kFast ? div_by_rcp(num, den, g.rden[e3]) : __fdiv_rn(num, den);
      float t[3];
      t[e3] = q; t[b] = 1.f - q; t[c] = 0.f;
      t[0] -= w0; t[1] -= w1; t[2] -= w2;
      const float dx = dot3(t[0], g.x[0], t[1], g.x[1], t[2], g.x[2]);
      const float dy = dot3(t[0], g.y[0], t[1], g.y[1], t[2], g.y[2]);
      const float dd = dot2(dx, dx, dy, dy);
      if (dd < best) { best = dd; bx = dx; by = dy; fr.t[0] = t[0]; fr.t[1] = t[1]; fr.t[2] = t[2]; }
    }
    fr.dx = bx; fr.dy = by; fr.sign = 1.f;
    return;
  }
  if (!inside) {
    if (w1 <= 0.f && w2 <= 0.f) {
      e = 0;
      if (g.obt == 0 && dot2(xp - g.x[0], g.x[2] - g.x[0], yp - g.y[0], g.y[2] - g.y[0]) > 0.f) e = 2;
    } else if (w2 <= 0.f && w0 <= 0.f) {
      e = 1;
      if (g.obt == 1 && dot2(xp - g.x[1], g.x[0] - g.x[1], yp - g.y[1], g.y[0] - g.y[1]) > 0.f) e = 0;
    } else if (w0 <= 0.f && w1 <= 0.f) {
      e = 2;
      if (g.obt == 2 && dot2(xp - g.x[2], g.x[1] - g.x[2], yp - g.y[2], g.y[1] - g.y[2]) > 0.f) e = 1;
    } else if (w0 <= 0.f) {
      e = 1;
    } else if (w1 <= 0.f) {
      e = 2;
    } else {
      e = 0;
    }
  }
  // pick the row of `a` for edge e without dynamic register indexing
  const float a0 = e == 0 ? g.a[0][0] : (e == 1 ? g.a[1][0] : g.a[2][0]);
  const float a1 = e == 0 ? g.a[0][1] : (e == 1 ? g.a[1][1] : g.a[2][1]);
  const float a2 = e == 0 ? g.a[0][2] : (e == 1 ? g.a[1][2] : g.a[2][2]);
  const float ae = e == 0 ? a0 : (e == 1 ? a1 : a2);
  const float ab = e == 0 ? a1 : (e == 1 ? a2 : a0);
  const float rd = e == 0 ? g.rden[0] : (e == 1 ? g.rden[1] : g.rden[2]);
  const float num = __fsub_rn(dot3(w0, a0, w1, a1, w2, a2), ab);
  const float den = __fsub_rn(ae, ab);
  const float q = kFast ? div_by_rcp(num, den, rd) : __fdiv_rn(num, den);
  // inside: the foot of the perpendicular on the edge's line, unclamped (.cu:81-87); outside: clamped
  // to the segment (.cu:136-139)
  const float qe = inside ? q : fminf(fmaxf(q, 0.f), 1.f);
  const float qb = inside ? 1.f - q : fminf(fmaxf(1.f - q, 0.f), 1.f);
  // t[e] = qe, t[(e+1)%3] = qb, t[(e+2)%3] = 0, then minus w
  const float t0 = (e == 0 ? qe : (e == 2 ? qb : 0.f)) - w0;
  const float t1 = (e == 1 ? qe : (e == 0 ? qb : 0.f)) - w1;
  const float t2 = (e == 2 ? qe : (e == 1 ? qb : 0.f)) - w2;
  fr.t[0] = t0; fr.t[1] = t1; fr.t[2] = t2;
  fr.dx = dot3(t0, g.x[0], t1, g.x[1], t2, g.x[2]);
  fr.dy = dot3(t0, g.y[0], t1, g.y[1], t2, g.y[2]);
  fr.sign = inside ? 1.f : -1.f;
}

// Returns false when the pair is rejected by the distance stage (.cu:389-404 / :556-572).
template <int kFast>
__device__ __forceinline__ bool soft_fragment(const RasterArgs& A, const FaceGeom& g, float xp, float yp,
                                              float thr, Fragment& fr) {
  fr.w[0] = __fadd_rn(__fmaf_rn(xp, g.inv[0], __fmul_rn(yp, g.inv[1])), g.inv[2]);  // :24-29
  fr.w[1] = __fadd_rn(__fmaf_rn(xp, g.inv[3], __fmul_rn(yp, g.inv[4])), g.inv[5]);
  fr.w[2] = __fadd_rn(__fmaf_rn(xp, g.inv[6], __fmul_rn(yp, g.inv[7])), g.inv[8]);
  const int dist = kFast ? 2 : A.dist;
  fr.sign = 0.f; fr.dx = 0.f; fr.dy = 0.f; fr.dis = 0.f;
  fr.t[0] = fr.w[0]; fr.t[1] = fr.w[1]; fr.t[2] = fr.w[2];
  if (dist == 0) {
    fr.frag = pixel_inside(fr.w) ? 1.f : 0.f;
    return fr.frag != 0.f;
  } else if (dist == 1) {  // :154-158
    const float w0 = fr.w[0], w1 = fr.w[1], w2 = fr.w[2];
    float d = w0 > w1 ? (w1 > w2 ? w2 : w1) : (w0 > w2 ? w2 : w0);
    d = d > 0.f ? d * d : -(d * d);
    fr.dis = d;
    if (-d >= thr) return false;
    fr.frag = kFast ? fast_rcp<kFast>(1.f + fast_exp<kFast>(-d * A.inv_sigma)) : 1.f / (1.f + expf(-d / A.sigma));
    return true;
  } else {
    euclid_distance<kFast>(g, xp, yp, fr);
    fr.dis = dot2(fr.dx, fr.dx, fr.dy, fr.dy);
    if (fr.sign < 0.f && fr.dis >= thr) return false;
    fr.frag = kFast ? fast_rcp<kFast>(1.f + fast_exp<kFast>(-fr.sign * fr.dis * A.inv_sigma))
                    : 1.f / (1.f + expf(-fr.sign * fr.dis / A.sigma));
    return true;
  }
}

// Clipped barycentrics (.cu:53-58) and the depth they interpolate (:423).
//   wc_k = c_k / s,  c = clamp(w, 0, 1),  s = max(sum c, 1e-5);   zp = 1 / sum_k wc_k / z_k = s / sum_k c_k / z_k
template <int kFast>
__device__ __forceinline__ float clip_and_depth(const FaceGeom& g, const float* w, float* wc) {
  const float c0 = fmaxf(fminf(w[0], 1.f), 0.f);
  const float c1 = fmaxf(fminf(w[1], 1.f), 0.f);
  const float c2 = fmaxf(fminf(w[2], 1.f), 0.f);
  const float s = fmaxf(c0 + c1 + c2, 1e-5f);
  if (!kFast) {  // the reference's own operation sequence (see div_/exp_ above)
    wc[0] = __fdiv_rn(c0, s); wc[1] = __fdiv_rn(c1, s); wc[2] = __fdiv_rn(c2, s);
    return __fdiv_rn(1.f, __fdiv_rn(wc[0], g.z[0]) + __fdiv_rn(wc[1], g.z[1]) + __fdiv_rn(wc[2], g.z[2]));
  }
  const float rs = fast_rcp<kFast>(s);
  wc[0] = c0 * rs; wc[1] = c1 * rs; wc[2] = c2 * rs;
  return s * fast_rcp<kFast>(c0 * g.iz[0] + c1 * g.iz[1] + c2 * g.iz[2]);
}

// 'surface' texel index of .cu:178-194 (may equal R*R: the reference reads the next face's block)
__device__ __forceinline__ int surface_texel(const float* wc, int R) {
  const int wx = (int)(wc[0] * R);
  const int wy = (int)(wc[1] * R);
  return ((wc[0] + wc[1]) * R - wx - wy <= 1.f) ? (wy * R + wx) : ((R - 1 - wy) * R + (R - 1 - wx));
}

// ---- forward: per-pixel state and the evaluation of one (pixel, face) pair ---------------------------------
// (in this header rather than in raster.cu so that the host harness can run the forward's own code on the CPU:
// tests/host_harness, tests/test_kernel_math_host.py)
#ifndef LASR_FTZ_MUFU_FWD           // bare FTZ MUFU reciprocal / ex2 in the forward (faster) ...
#define LASR_FTZ_MUFU_FWD 1
#endif
template <int K>
struct PixelState {
  float alpha;             // running product / sum / flag
  float ssum, smax;        // online softmax state
  float col[3 * K];        // K colour sets share the geometry work of one pass
  float depth_min;         // hard rgb
  int face_min;
};

struct FwdConsts {
  float thr, near_, far_, inv_range;
};

// rec: the face's record in global memory, or -- rec == nullptr -- staged in shared memory at shared-space address srec
template <bool kFast, int K>
__device__ __forceinline__ void forward_pair(const RasterArgs& A, const FwdConsts& C, const float* __restrict__ rec,
                                             float4 bd, int face, int n, float xp, float yp, PixelState<K>& st,
                                             uint32_t srec = 0) {
  constexpr int kMode = kFast ? (LASR_FTZ_MUFU_FWD ? 2 : 1) : 0;   // arithmetic mode, see raster_common.cuh
  // exact border predicate of the reference (.cu:32-38,375)
  if (xp > bd.y || xp < bd.x || yp > bd.w || yp < bd.z) return;
  FaceGeom g;
  float tex9[9 * K];
#ifdef __CUDACC__
  if (rec == nullptr) load_record_shared<K>(srec, g, tex9);
  else
#endif
  load_record<K>(rec, g, tex9);
  Fragment fr;
  if (!soft_fragment<kMode>(A, g, xp, yp, C.thr, fr)) return;
  const int alpha_f = kFast ? 2 : A.alpha;
  const int rgb_f = kFast ? 1 : A.rgb;
  const int sample = kFast ? 1 : A.sample;
  // alpha aggregation (.cu:408-417)
  if (alpha_f == 0) { if (fr.frag > 0.5f) st.alpha = 1.f; }
  else if (alpha_f == 1) st.alpha += fr.frag;
  else st.alpha *= 1.f - fr.frag;
  // clipped-barycentric depth (.cu:421-424)
  float wc[3];
  const float zp = clip_and_depth<kMode>(g, fr.w, wc);
  if (zp < C.near_ || zp > C.far_) return;
  const bool visible = g.front || (kFast ? true : A.double_side != 0);
  auto sample_tex = [&](int k) -> float {   // k = 3*set + channel
    if (sample == 1) {
      const float* t = tex9 + 9 * (k / 3);
      return wc[0] * t[k % 3] + wc[1] * t[3 + k % 3] + wc[2] * t[6 + k % 3];
    }
    const long long base = ((long long)n * A.F + face) * A.T * 3;
    const long long idx = base + (long long)surface_texel(wc, A.R) * 3 + k;
    return idx < (long long)A.N * A.F * A.T * 3 ? A.textures[idx] : 0.f;
  };
  if (rgb_f == 0) {  // hard z-buffer (.cu:428-435)
    if (zp < st.depth_min && pixel_inside(fr.w) && visible) {
      st.depth_min = zp;
      st.face_min = face;
#pragma unroll
      for (int k = 0; k < 3; ++k) st.col[k] = sample_tex(k);
    }
  } else if (visible) {  // depth softmax, online (.cu:437-451)
    const float zn = kFast ? (C.far_ - zp) * C.inv_range : (C.far_ - zp) / (C.far_ - C.near_);
    float rescale = 1.f;
    if (zn > st.smax) {
      rescale = kFast ? fast_exp<kMode>((st.smax - zn) * A.inv_gamma) : expf((st.smax - zn) / A.gamma);
      st.smax = zn;
    }
    const float ez = (kFast ? fast_exp<kMode>((zn - st.smax) * A.inv_gamma) : expf((zn - st.smax) / A.gamma)) * fr.frag;
    st.ssum = rescale * st.ssum + ez;
#pragma unroll
    for (int k = 0; k < 3 * K; ++k) st.col[k] = rescale * st.col[k] + ez * sample_tex(k);
  }
}

// Sum 16 per-lane values over the warp with 16 shuffles instead of 80 (recursive halving: at each
// stage a lane keeps half of its values and ships the other half to its partner).  On return lane L
// holds in v[0] the warp total of value  idx = 8*bit4(L) + 4*bit3(L) + 2*bit2(L) + bit1(L).
// Fixed combination order => deterministic.
__device__ __forceinline__ void warp_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int stage = 0; stage < 4; ++stage) {
    const int o = 16 >> stage;         // 16, 8, 4, 2
    const int half = 8 >> stage;       // values kept after this stage: 8, 4, 2, 1
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float a = v[i], b = v[i + half];
      const float keep = upper ? b : a;
      const float send = upper ? a : b;
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// warp butterfly sum (fixed order => deterministic)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lasr
