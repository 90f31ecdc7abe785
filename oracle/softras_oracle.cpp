// =============================================================================
// oracle/softras_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference SoftRas soft-rasterizer (forward + backward)
// used by google/lasr.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl cpu leg may call into this file; the product path
// (lasr_b200/csrc) never does.
//
// Reference followed (all paths relative to /root/reference):
//   third_party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu
//     :24-58    barycentric / border / frontside / inside / clip helpers
//     :61-151   euclidean point-to-face distance (incl. obtuse-corner logic)
//     :154-175  barycentric distance fwd/bwd
//     :178-216  texture sampling fwd/bwd ('surface' R x R texels, 'vertex')
//     :245-305  per-face precompute (face_inv, face_sym, face_obt)
//     :308-483  per-pixel forward
//     :486-668  per-pixel backward (atomicAdd -> ordered double accumulation)
//   third_party/softras/soft_renderer/functional/soft_rasterize.py:22-53,88-91
//     enum maps, dist_eps -> log(1/eps - 1), buffer initialisation (done by the
//     Python wrapper oracle/oracle.py, exactly as the reference's Python does).
//
// FMA contraction: nvcc compiles the reference with -fmad=true, so sums of products in the
// geometric predicates become fused multiply-adds.  For edge-on sliver faces those predicates
// are ill-conditioned (a 1-ulp change of `det` flips a pixel between inside and outside), so the
// association matters.  With `contract` on (the default; lasr_oracle_set_contract) the handful of
// sums of products that feed hard decisions are evaluated with std::fma in exactly the
// association read from `cuobjdump -sass` of the reference build for sm_100a
// (a*b + c*d -> fma(a,b,c*d);  a*b + c*d + e*f -> fma(e,f,fma(a,b,c*d));  a*b - c*d ->
// fma(a,b,-(c*d))); with it off every operation is individually rounded (plain C++).
//
// Numerics: the reference kernels are templated on scalar_t but keep `float`
// scalar parameters and `double` literals, so the float instantiation promotes a
// few sub-expressions to fp64.  This file is templated
// the same way (T = float | double) and writes those promotions out explicitly
// (every `double(...)` / `1.0` below is a promotion the reference's expression
// text implies under the usual arithmetic conversions).  Compiled with
// -ffp-contract=off: no FMA contraction, unlike nvcc's default -fmad=true, so
// this is NOT bit-identical to the reference on a GPU; the measured gap is
// recorded in tests/golden/README.md.
//
// min / max: the reference writes `min(max(t, 0.), 1.)` etc. in device code, where min / max on doubles are fmin /
// fmax -- a NaN operand is dropped.  That matters for exactly degenerate faces (a zero-length edge makes the
// closest-point parameter 0/0, which the clamp turns into 0 on the GPU); std::min / std::max would propagate the NaN
// into every pixel of the face's box.  std::fmin / std::fmax are used wherever the reference clamps.
//
// Parity status: pinned against outputs of the reference kernels themselves
// (oracle/_ref, built from the unmodified reference sources) run on a B200 --
// see tests/golden/ and tests/test_oracle_vs_reference_golden.py.
//
// Loop structure differs from the reference on purpose (this is a restatement,
// not a copy): forward is pixel-major with faces visited in ascending index
// (the order the reference's per-pixel loop uses, which matters for the online
// softmax and the hard z-test); backward is face-major so each face's gradient
// is a private sum (deterministic, accumulated in double, rounded once).
// `accel != 0` restricts pixel/face candidates with a conservative bounding
// box before applying the reference's exact per-pixel border predicate, which
// is value-identical to the brute-force O(N*S^2*F) scan (accel == 0).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_contract = 1;

template <class T> inline T dot2(T a, T b, T c, T d) { return g_contract ? std::fma(a, b, c * d) : a * b + c * d; }
template <class T> inline T dot3(T a, T b, T c, T d, T e, T f) {
  return g_contract ? std::fma(e, f, std::fma(a, b, c * d)) : a * b + c * d + e * f;
}
template <class T> inline T cross2(T a, T b, T c, T d) { return g_contract ? std::fma(a, b, -(c * d)) : a * b - c * d; }

struct RasterParams {
  int N, F, S, texture_size, texture_res;
  float near_, far_, eps, sigma_val, dist_eps, gamma_val;  // `float` in the reference signature
  int dist_func;    // 0 hard, 1 barycentric, 2 euclidean
  int rgb_func;     // 0 hard, 1 softmax
  int alpha_func;   // 0 hard, 1 sum, 2 prod
  int sample_type;  // 0 surface, 1 vertex
  bool double_side;
};

// ---- helpers (.cu:24-58) -----------------------------------------------------
template <class T>
inline void bary_coords(T* w, T x, T y, const T* inv) {  // :24-29
  w[0] = dot2(inv[0], x, inv[1], y) + inv[2];
  w[1] = dot2(inv[3], x, inv[4], y) + inv[5];
  w[2] = dot2(inv[6], x, inv[7], y) + inv[8];
}

template <class T>
inline bool outside_border(T x, T y, const T* fv, T thr) {  // :32-38 (true => skip)
  const T xmax = std::max(std::max(fv[0], fv[3]), fv[6]);
  const T xmin = std::min(std::min(fv[0], fv[3]), fv[6]);
  const T ymax = std::max(std::max(fv[1], fv[4]), fv[7]);
  const T ymin = std::min(std::min(fv[1], fv[4]), fv[7]);
  return (x > xmax + thr || x < xmin - thr || y > ymax + thr || y < ymin - thr);
}

template <class T>
inline bool frontside(const T* fv) {  // :41-44
  return (fv[7] - fv[1]) * (fv[3] - fv[0]) < (fv[4] - fv[1]) * (fv[6] - fv[0]);
}

template <class T>
inline bool pixel_inside(const T* w) {  // :47-50
  return w[0] <= 1 && w[0] >= 0 && w[1] <= 1 && w[1] >= 0 && w[2] <= 1 && w[2] >= 0;
}

template <class T>
inline void bary_clip(T* w) {  // :53-58  (min/max against double literals)
  // CUDA's min/max on doubles are fmin/fmax: a NaN operand is dropped, not propagated (std::min/std::max would keep it)
  for (int k = 0; k < 3; ++k) w[k] = T(std::fmax(std::fmin(double(w[k]), 1.0), 0.0));
  const T wsum = T(std::fmax(double(w[0] + w[1] + w[2]), 1e-5));
  for (int k = 0; k < 3; ++k) w[k] /= wsum;
}

// ---- euclidean distance (.cu:61-151) ------------------------------------------
template <class T>
inline void euclid_p2f(T& sign, T& dx, T& dy, const T* w, T* t, const T* fv, const T* info, T xp,
                       T yp) {
  const T* sym = info + 9;
  const T* obt = info + 18;
  const bool strictly_inside =
      w[0] > 0 && w[1] > 0 && w[2] > 0 && w[0] < 1 && w[1] < 1 && w[2] < 1;
  if (strictly_inside) {  // :69-109 closest of the three (unclamped) edge projections
    T best = 100000000;
    T bx = 0, by = 0;
    for (int e = 0; e < 3; ++e) {
      const int a = e, b = (e + 1) % 3, c = (e + 2) % 3;
      T d[3] = {sym[3 * a + 0] - sym[3 * b + 0], sym[3 * a + 1] - sym[3 * b + 1],
                sym[3 * a + 2] - sym[3 * b + 2]};
      T q[3];
      q[a] = (dot3(w[0], d[0], w[1], d[1], w[2], d[2]) - d[b]) / (d[a] - d[b]);
      q[b] = 1 - q[a];
      q[c] = 0;
      q[0] -= w[0];
      q[1] -= w[1];
      q[2] -= w[2];
      dx = dot3(q[0], fv[0], q[1], fv[3], q[2], fv[6]);
      dy = dot3(q[0], fv[1], q[1], fv[4], q[2], fv[7]);
      const T dd = dot2(dx, dx, dy, dy);
      if (dd < best) {
        best = dd;
        bx = dx;
        by = dy;
        t[0] = q[0];
        t[1] = q[1];
        t[2] = q[2];
      }
    }
    dx = bx;
    dy = by;
    sign = 1;
  } else {  // :110-150 outside: pick the edge (or corner) region, clamp onto the segment
    int a = -1;
    if (w[1] <= 0 && w[2] <= 0) {
      a = 0;
      if (obt[0] == 1 && dot2(xp - fv[0], fv[6] - fv[0], yp - fv[1], fv[7] - fv[1]) > 0) a = 2;
    } else if (w[2] <= 0 && w[0] <= 0) {
      a = 1;
      if (obt[1] == 1 && dot2(xp - fv[3], fv[0] - fv[3], yp - fv[4], fv[1] - fv[4]) > 0) a = 0;
    } else if (w[0] <= 0 && w[1] <= 0) {
      a = 2;
      if (obt[2] == 1 && dot2(xp - fv[6], fv[3] - fv[6], yp - fv[7], fv[4] - fv[7]) > 0) a = 1;
    } else if (w[0] <= 0) {
      a = 1;
    } else if (w[1] <= 0) {
      a = 2;
    } else if (w[2] <= 0) {
      a = 0;
    }
    // NOTE: a stays -1 only if some w[k] >= 1 while all are > 0, which cannot happen for
    // finite barycentrics summing to ~1; the reference would index out of bounds there too.
    if (a < 0) a = 0;
    const int b = (a + 1) % 3, c = (a + 2) % 3;
    T d[3] = {sym[3 * a + 0] - sym[3 * b + 0], sym[3 * a + 1] - sym[3 * b + 1],
              sym[3 * a + 2] - sym[3 * b + 2]};
    t[a] = (dot3(w[0], d[0], w[1], d[1], w[2], d[2]) - d[b]) / (d[a] - d[b]);
    t[b] = 1 - t[a];
    t[c] = 0;
    for (int k = 0; k < 3; ++k) {
      t[k] = T(std::fmin(std::fmax(double(t[k]), 0.0), 1.0));   // fmax(NaN, 0) = 0: a zero-length edge (0/0) clamps to 0, as on the GPU
      t[k] -= w[k];
    }
    dx = dot3(t[0], fv[0], t[1], fv[3], t[2], fv[6]);
    dy = dot3(t[0], fv[1], t[1], fv[4], t[2], fv[7]);
    sign = -1;
  }
}

// ---- barycentric distance (.cu:154-175) ---------------------------------------
template <class T>
inline T bary_dist(const T* w) {
  T d = w[0] > w[1] ? (w[1] > w[2] ? w[2] : w[1]) : (w[0] > w[2] ? w[2] : w[0]);
  return d > 0 ? d * d : -(d * d);  // pow(d, 2) with an int exponent == d*d
}

template <class T>
inline void bary_dist_grad(T g[3][3], const T* w, const T* inv, T xp, T yp, T dis, T C) {
  const int p = w[0] > w[1] ? (w[1] > w[2] ? 2 : 1) : (w[0] > w[2] ? 2 : 0);
  for (int l = 0; l < 2; ++l)
    for (int k = 0; k < 3; ++k) {
      T acc = 0;
      for (int q = 0; q < 3; ++q)
        acc += -inv[3 * p + l] * inv[3 * k + q] * (q == 0 ? xp : (q == 1 ? yp : T(1)));
      g[k][l] = acc * C;
      const double s = dis > 0 ? (2.0 * double(std::sqrt(dis))) : (2.0 * double(std::sqrt(-dis)));
      g[k][l] = T(double(g[k][l]) * s);
    }
}

// ---- texture sampling (.cu:178-216) -------------------------------------------
// `tex_end` guards the one place where the reference's 'surface' index formula can
// step past the face's own R*R texel block (w[0]*R == R when a clipped w[0] == 1):
// inside the tensor we return what the reference would read (the neighbouring
// block); past the end of the tensor the reference reads out of bounds (undefined),
// we return 0.
template <class T>
inline T sample_tex(const T* tex, const T* tex_end, const T* w, int R, int k, int type) {
  if (type == 0) {
    const int wx = int(w[0] * R);
    const int wy = int(w[1] * R);
    const T* p = ((w[0] + w[1]) * R - wx - wy <= 1) ? tex + (wy * R + wx) * 3 + k
                                                    : tex + ((R - 1 - wy) * R + (R - 1 - wx)) * 3 + k;
    return p < tex_end ? *p : T(0);
  }
  return w[0] * tex[k] + w[1] * tex[3 + k] + w[2] * tex[6 + k];
}

template <class T>
inline T sample_tex_grad(T g, const T* w, int R, int j, int type) {
  if (type == 0) {
    const int wx = int(w[0] * R);
    const int wy = int(w[1] * R);
    if ((w[0] + w[1]) * R - wx - wy <= 1) return j == wy * R + wx ? g : T(0);
    return j == (R - 1 - wy) * R + (R - 1 - wx) ? g : T(0);
  }
  return w[j] * g;  // reference indexes w[k] with the texel index (valid for j < 3)
}

// ---- per-face precompute (.cu:245-305) ----------------------------------------
template <class T>
inline void face_precompute(const T* fv, T* info) {
  T* inv = info;
  T* sym = info + 9;
  T* obt = info + 18;
  const T px[3] = {fv[0], fv[3], fv[6]};
  const T py[3] = {fv[1], fv[4], fv[7]};
  const T adj[9] = {py[1] - py[2], px[2] - px[1], cross2(px[1], py[2], px[2], py[1]),
                    py[2] - py[0], px[0] - px[2], cross2(px[2], py[0], px[0], py[2]),
                    py[0] - py[1], px[1] - px[0], cross2(px[0], py[1], px[1], py[0])};
  T det = dot3(px[2], py[0] - py[1], px[0], py[1] - py[2], px[1], py[2] - py[0]);
  det = det > 0 ? T(std::fmax(double(det), 1e-10)) : T(std::fmin(double(det), -1e-10));
  for (int k = 0; k < 9; ++k) inv[k] = adj[k] / det;
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) sym[3 * j + k] = dot2(px[j], px[k], py[j], py[k]) + 1;
  // obt is only ever SET (caller zero-fills, soft_rasterize.py:47); first obtuse corner wins
  for (int k = 0; k < 3; ++k) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    if (dot2(px[k1] - px[k], px[k2] - px[k], py[k1] - py[k], py[k2] - py[k]) < 0) {
      obt[k] = 1;
      break;
    }
  }
}

// conservative pixel-index window [lo, hi] of coordinates c with  cmin-thr <= c <= cmax+thr,
// padded by 2 px so that the exact float predicate (outside_border) stays the only decider.
inline void conservative_window(double cmin, double cmax, double thr, int S, int& lo, int& hi) {
  double a = ((cmin - thr) * S + S - 1.0) * 0.5;
  double b = ((cmax + thr) * S + S - 1.0) * 0.5;
  // boxes reaching far beyond the screen (|pixel coordinate| past the int range: out-of-range float -> int conversion is
  // undefined and used to come back as an empty window): clamp first, as the kernels' coord_window does
  a = std::min(std::max(a, -4.0), double(S) + 4.0);
  b = std::min(std::max(b, -4.0), double(S) + 4.0);
  lo = std::max(0, int(std::floor(a)) - 2);
  hi = std::min(S - 1, int(std::ceil(b)) + 2);
  if (!(a == a) || !(b == b)) {  // NaN vertices: let the exact predicate decide everywhere
    lo = 0;
    hi = S - 1;
  }
}

template <class T>
inline T pixel_coord(int idx, int S) {  // :345-346  (2. * i + 1. - is) / is  in double
  return T((2.0 * idx + 1.0 - S) / S);
}

// Evaluate the soft fragment of one (pixel, face) pair; returns false if the pair is
// skipped by the distance stage (:389-404 / :556-572).
template <class T>
inline bool soft_fragment(const RasterParams& P, const T* fv, const T* info, T xp, T yp, T thr,
                          T* w, T* t, T& sign, T& dx, T& dy, T& dis, T& frag) {
  bary_coords(w, xp, yp, info);
  sign = 0;
  dx = dy = dis = 0;
  t[0] = t[1] = t[2] = 0;
  if (P.dist_func == 0) {
    frag = pixel_inside(w) ? T(1) : T(0);
    if (frag == 0) return false;
  } else if (P.dist_func == 1) {
    dis = bary_dist(w);
    for (int k = 0; k < 3; ++k) t[k] = w[k];
    if (-dis >= thr) return false;
    frag = T(1.0 / (1.0 + double(std::exp(-dis / P.sigma_val))));
  } else {
    euclid_p2f(sign, dx, dy, w, t, fv, info, xp, yp);
    dis = dot2(dx, dx, dy, dy);
    if (sign < 0 && dis >= thr) return false;
    frag = T(1.0 / (1.0 + double(std::exp(-sign * dis / P.sigma_val))));
  }
  return true;
}

// ---- forward (.cu:308-483) ------------------------------------------------------
template <class T>
void forward_impl(const RasterParams& P, const T* faces, const T* textures, T* faces_info,
                  T* aggrs_info, T* soft_colors, int accel, long long* pair_count) {
  const int N = P.N, F = P.F, S = P.S, TS = P.texture_size;
  const long long SS = (long long)S * S;
  const T* tex_end = textures + (long long)N * F * TS * 3;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)N * F; ++i) face_precompute(faces + i * 9, faces_info + i * 27);

  const T thr = P.dist_eps * P.sigma_val;  // float * float
  const T sthr = std::sqrt(thr);
  long long pairs_total = 0;

  for (int n = 0; n < N; ++n) {
    const T* fN = faces + (long long)n * F * 9;
    const T* iN = faces_info + (long long)n * F * 27;
    const T* tN = textures + (long long)n * F * TS * 3;
    // per-face conservative windows (accel only)
    std::vector<int> x0, x1, y0, y1;
    if (accel) {
      x0.resize(F); x1.resize(F); y0.resize(F); y1.resize(F);
      for (int f = 0; f < F; ++f) {
        const T* fv = fN + f * 9;
        conservative_window(std::min(std::min(fv[0], fv[3]), fv[6]),
                            std::max(std::max(fv[0], fv[3]), fv[6]), sthr, S, x0[f], x1[f]);
        conservative_window(std::min(std::min(fv[1], fv[4]), fv[7]),
                            std::max(std::max(fv[1], fv[4]), fv[7]), sthr, S, y0[f], y1[f]);
      }
    }
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : pairs_total)
    for (int row = 0; row < S; ++row) {  // row 0 = top of the image = yi = S-1 (:343)
      const int yi = S - 1 - row;
      const T yp = pixel_coord<T>(yi, S);
      std::vector<int> cand;
      if (accel) {
        for (int f = 0; f < F; ++f)
          if (yi >= y0[f] && yi <= y1[f]) cand.push_back(f);
      }
      const int ncand = accel ? (int)cand.size() : F;
      for (int xi = 0; xi < S; ++xi) {
        const long long pn = (long long)row * S + xi;
        const T xp = pixel_coord<T>(xi, S);
        T col[4] = {1, 1, 1, 0};
        if (P.alpha_func == 2) col[3] = 1;
        T ssum = std::exp(P.eps / P.gamma_val);  // expf(float / float)
        T smax = P.eps;
        for (int k = 0; k < 3; ++k) {
          const T bg = soft_colors[((long long)n * 4 + k) * SS + pn];
          if (P.rgb_func == 0) col[k] = bg;
          else if (P.rgb_func == 1) col[k] = bg * ssum;
        }
        T depth_min = 10000000;
        int face_min = -1;
        for (int ci = 0; ci < ncand; ++ci) {
          const int f = accel ? cand[ci] : ci;
          if (accel && (xi < x0[f] || xi > x1[f])) continue;
          const T* fv = fN + f * 9;
          const T* info = iN + f * 27;
          const T* tex = tN + (long long)f * TS * 3;
          if (outside_border(xp, yp, fv, sthr)) continue;
          T w[3], t[3], sign, dx, dy, dis, frag;
          if (!soft_fragment(P, fv, info, xp, yp, thr, w, t, sign, dx, dy, dis, frag)) continue;
          ++pairs_total;
          // alpha aggregation (:408-417)
          if (P.alpha_func == 0) {
            if (frag > 0.5) col[3] = 1;
          } else if (P.alpha_func == 1) {
            col[3] += frag;
          } else {
            col[3] = T(double(col[3]) * (1.0 - double(frag)));
          }
          // depth (:421-424)
          T wc[3] = {w[0], w[1], w[2]};
          bary_clip(wc);
          const T zp = T(1.0 / double(wc[0] / fv[2] + wc[1] / fv[5] + wc[2] / fv[8]));
          if (zp < P.near_ || zp > P.far_) continue;
          // rgb aggregation (:428-452)
          if (P.rgb_func == 0) {
            if (zp < depth_min && pixel_inside(w) && (P.double_side || frontside(fv))) {
              depth_min = zp;
              face_min = f;
              for (int k = 0; k < 3; ++k) col[k] = sample_tex(tex, tex_end, wc, P.texture_res, k, P.sample_type);
            }
          } else if (P.rgb_func == 1) {
            if (frontside(fv) || P.double_side) {
              const T zn = (P.far_ - zp) / (P.far_ - P.near_);
              T rescale = 1;
              if (zn > smax) {
                rescale = std::exp((smax - zn) / P.gamma_val);
                smax = zn;
              }
              const T ez = std::exp((zn - smax) / P.gamma_val);
              ssum = rescale * ssum + ez * frag;
              for (int k = 0; k < 3; ++k) {
                const T c = sample_tex(tex, tex_end, wc, P.texture_res, k, P.sample_type);
                col[k] = rescale * col[k] + ez * frag * c;
              }
            }
          }
        }
        // finalize (:457-482)
        T* out_a = soft_colors + ((long long)n * 4 + 3) * SS + pn;
        if (P.alpha_func == 0) *out_a = col[3];
        else if (P.alpha_func == 1) *out_a = col[3] / F;
        else *out_a = T(1.0 - double(col[3]));
        if (P.rgb_func == 0) {
          if (face_min != -1)
            for (int k = 0; k < 3; ++k) soft_colors[((long long)n * 4 + k) * SS + pn] = col[k];
          aggrs_info[((long long)n * 2 + 0) * SS + pn] = depth_min;
          aggrs_info[((long long)n * 2 + 1) * SS + pn] = T(face_min);
        } else if (P.rgb_func == 1) {
          for (int k = 0; k < 3; ++k) soft_colors[((long long)n * 4 + k) * SS + pn] = col[k] / ssum;
          aggrs_info[((long long)n * 2 + 0) * SS + pn] = ssum;
          aggrs_info[((long long)n * 2 + 1) * SS + pn] = smax;
        }
      }
    }
  }
  if (pair_count) *pair_count = pairs_total;
}

// ---- backward (.cu:486-668) -----------------------------------------------------
template <class T>
void backward_impl(const RasterParams& P, const T* faces, const T* textures, const T* soft_colors,
                   const T* faces_info, const T* aggrs_info, T* grad_faces, T* grad_textures,
                   const T* grad_soft_colors, int accel) {
  const int N = P.N, F = P.F, S = P.S, TS = P.texture_size;
  const long long SS = (long long)S * S;
  const T* tex_end = textures + (long long)N * F * TS * 3;
  const T thr = P.dist_eps * P.sigma_val;
  const T sthr = std::sqrt(thr);

#pragma omp parallel for schedule(dynamic, 16)
  for (long long nf = 0; nf < (long long)N * F; ++nf) {
    const int n = int(nf / F);
    const int f = int(nf % F);
    const T* fv = faces + nf * 9;
    const T* info = faces_info + nf * 27;
    const T* tex = textures + nf * TS * 3;
    double gface[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> gtex((size_t)TS * 3, 0.0);
    int xlo = 0, xhi = S - 1, ylo = 0, yhi = S - 1;
    if (accel) {
      conservative_window(std::min(std::min(fv[0], fv[3]), fv[6]),
                          std::max(std::max(fv[0], fv[3]), fv[6]), sthr, S, xlo, xhi);
      conservative_window(std::min(std::min(fv[1], fv[4]), fv[7]),
                          std::max(std::max(fv[1], fv[4]), fv[7]), sthr, S, ylo, yhi);
    }
    // visit pixels in the reference's thread order (pn ascending = row-major from the top)
    for (int row = S - 1 - yhi; row <= S - 1 - ylo; ++row) {
      const int yi = S - 1 - row;
      const T yp = T((2.0 * yi + 1 - S) / S);  // :526
      for (int xi = xlo; xi <= xhi; ++xi) {
        const T xp = T((2.0 * xi + 1 - S) / S);
        if (outside_border(xp, yp, fv, sthr)) continue;
        const long long pn = (long long)row * S + xi;
        T w[3], t[3], sign, dx, dy, dis, frag;
        if (!soft_fragment(P, fv, info, xp, yp, thr, w, t, sign, dx, dy, dis, frag)) continue;
        const T ssum = aggrs_info[((long long)n * 2 + 0) * SS + pn];
        const T smax = aggrs_info[((long long)n * 2 + 1) * SS + pn];
        const T* gsc = grad_soft_colors + (long long)n * 4 * SS + pn;
        const T* sc = soft_colors + (long long)n * 4 * SS + pn;

        T gv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        T Cxy = 0;
        // alpha (:582-593)
        T Ca = gsc[3 * SS];
        if (P.alpha_func == 1) {
          Ca /= F;
        } else if (P.alpha_func == 2) {
          Ca = T(double(Ca) * (double(1 - sc[3 * SS]) / std::fmax(double(1 - frag), 1e-6)));
        }
        Cxy += Ca;
        // depth (:596-599)
        const T w0[3] = {w[0], w[1], w[2]};
        bary_clip(w);
        const T zp = T(1.0 / double(w[0] / fv[2] + w[1] / fv[5] + w[2] / fv[8]));
        if (zp < P.near_ || zp > P.far_) continue;

        if (P.rgb_func == 0) {  // :602-609 hard: texture gradient of the winning face only
          if (T(f) == smax) {
            for (int k = 0; k < 3; ++k)
              for (int j = 0; j < TS; ++j)
                gtex[3 * j + k] += double(sample_tex_grad(gsc[k * SS], w, P.texture_res, j, P.sample_type));
          }
        } else if (P.rgb_func == 1 && (frontside(fv) || P.double_side)) {  // :611-640
          T Crgb = 0;
          const T zn = (P.far_ - zp) / (P.far_ - P.near_);
          const T wgt = frag * std::exp((zn - smax) / P.gamma_val) / ssum;
          for (int k = 0; k < 3; ++k) {
            const T g = gsc[k * SS];
            for (int j = 0; j < TS; ++j)
              gtex[3 * j + k] += double(wgt * sample_tex_grad(g, w, P.texture_res, j, P.sample_type));
            const T c = sample_tex(tex, tex_end, w, P.texture_res, k, P.sample_type);
            Crgb += g * (c - sc[k * SS]);
          }
          Crgb *= wgt;
          Cxy += Crgb / frag;
          const T Cz = Crgb / P.gamma_val / (P.near_ - P.far_) * zp * zp;
          gv[0][2] = Cz * w[0] / fv[2] / fv[2];
          gv[1][2] = Cz * w[1] / fv[5] / fv[5];
          gv[2][2] = Cz * w[2] / fv[8] / fv[8];
        }
        // distance gradient (:644-655)
        Cxy *= frag * (1 - frag) / P.sigma_val;
        if (P.dist_func == 1) {
          bary_dist_grad(gv, t, info, xp, yp, dis, Cxy);
        } else if (P.dist_func == 2) {
          for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 2; ++l) gv[k][l] = 2 * sign * Cxy * (t[k] + w0[k]) * (l == 0 ? dx : dy);
        }
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) gface[3 * k + l] += double(gv[k][l]);
      }
    }
    for (int k = 0; k < 9; ++k) grad_faces[nf * 9 + k] += T(gface[k]);
    for (int k = 0; k < TS * 3; ++k) grad_textures[nf * TS * 3 + k] += T(gtex[k]);
  }
}

RasterParams make_params(int N, int F, int S, int texture_size, float near_, float far_, float eps,
                         float sigma_val, int dist_func, float dist_eps, float gamma_val,
                         int rgb_func, int alpha_func, int sample_type, int double_side) {
  RasterParams P;
  P.N = N; P.F = F; P.S = S;
  P.texture_size = texture_size;
  P.texture_res = int(std::sqrt(double(texture_size)));  // .cu:696
  P.near_ = near_; P.far_ = far_; P.eps = eps; P.sigma_val = sigma_val;
  P.dist_eps = dist_eps; P.gamma_val = gamma_val;
  P.dist_func = dist_func; P.rgb_func = rgb_func; P.alpha_func = alpha_func;
  P.sample_type = sample_type; P.double_side = double_side != 0;
  return P;
}

void set_threads(int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#else
  (void)nthreads;
#endif
}

}  // namespace

// C ABI -- argument order mirrors the reference pybind functions
// (soft_rasterize_cuda.cpp:59-76, 94-114) with sizes made explicit.
#define ORACLE_FWD(NAME, TYPE)                                                                     \
  extern "C" int NAME(const TYPE* faces, const TYPE* textures, TYPE* faces_info, TYPE* aggrs_info, \
                      TYPE* soft_colors, int N, int F, int S, int texture_size, float near_,       \
                      float far_, float eps, float sigma_val, int dist_func, float dist_eps,       \
                      float gamma_val, int rgb_func, int alpha_func, int sample_type,              \
                      int double_side, int accel, int nthreads, long long* pair_count) {           \
    set_threads(nthreads);                                                                         \
    RasterParams P = make_params(N, F, S, texture_size, near_, far_, eps, sigma_val, dist_func,    \
                                 dist_eps, gamma_val, rgb_func, alpha_func, sample_type,           \
                                 double_side);                                                     \
    forward_impl<TYPE>(P, faces, textures, faces_info, aggrs_info, soft_colors, accel, pair_count);\
    return 0;                                                                                      \
  }

#define ORACLE_BWD(NAME, TYPE)                                                                     \
  extern "C" int NAME(const TYPE* faces, const TYPE* textures, const TYPE* soft_colors,            \
                      const TYPE* faces_info, const TYPE* aggrs_info, TYPE* grad_faces,            \
                      TYPE* grad_textures, const TYPE* grad_soft_colors, int N, int F, int S,      \
                      int texture_size, float near_, float far_, float eps, float sigma_val,       \
                      int dist_func, float dist_eps, float gamma_val, int rgb_func,                \
                      int alpha_func, int sample_type, int double_side, int accel, int nthreads) { \
    set_threads(nthreads);                                                                         \
    RasterParams P = make_params(N, F, S, texture_size, near_, far_, eps, sigma_val, dist_func,    \
                                 dist_eps, gamma_val, rgb_func, alpha_func, sample_type,           \
                                 double_side);                                                     \
    backward_impl<TYPE>(P, faces, textures, soft_colors, faces_info, aggrs_info, grad_faces,       \
                        grad_textures, grad_soft_colors, accel);                                   \
    return 0;                                                                                      \
  }

ORACLE_FWD(lasr_oracle_forward_f32, float)
ORACLE_FWD(lasr_oracle_forward_f64, double)
ORACLE_BWD(lasr_oracle_backward_f32, float)
ORACLE_BWD(lasr_oracle_backward_f64, double)

extern "C" void lasr_oracle_set_contract(int on) { g_contract = on ? 1 : 0; }
extern "C" int lasr_oracle_get_contract(void) { return g_contract; }

extern "C" int lasr_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
