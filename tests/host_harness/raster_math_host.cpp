// TEST INFRASTRUCTURE.  Compiles lasr_b200/csrc/raster_common.cuh for the HOST (g++) so that the
// per-(pixel, face) arithmetic of the CUDA kernels -- record setup, barycentrics, the euclidean
// distance stage with its edge selection -- can be exercised on a machine without a GPU and compared
// with the C++ oracle bit for bit.  The device intrinsics are mapped to their IEEE host equivalents
// (explicit fma / mul / add, so -ffp-contract=off keeps the association the kernels spell out).
// Build: see tests/test_kernel_math_host.py.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include <cuda_runtime.h>   // vector types only; nothing is launched

static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __expf(x) expf(x)
static inline float rsqrtf(float x) { return 1.f / sqrtf(x); }
static inline float atomicAdd(float* p, float v) { const float o = *p; *p = o + v; return o; }
#define LASR_HOST_STATS 1
static long long lasr_host_inside_pairs = 0, lasr_host_all_edges_pairs = 0;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
namespace lasr {
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
}

#include "raster_common.cuh"

using namespace lasr;

// For n (face, pixel) pairs: face_setup -> store_record -> load_record (the kernels' round trip) ->
// soft_fragment<kFast=true> with LASR's distance mode.  out[n][10] = accepted, w0,w1,w2, t0,t1,t2, dx, dy,
// dis (accepted == 0: rejected by the distance stage); out_frag[n] = D.
extern "C" void lasr_host_soft_fragment(const float* faces9, const float* px, const float* py, int n, float sigma,
                                        float dist_eps, float* out, float* out_frag) {
  RasterArgs A{};
  A.dist = 2; A.sigma = sigma; A.dist_eps = dist_eps; A.inv_sigma = (float)(1.0 / (double)sigma);
  const float thr = dist_eps * sigma;
  const float sthr = sqrtf(thr);
  for (int i = 0; i < n; ++i) {
    FaceGeom g0, g;
    face_setup(faces9 + 9 * i, sthr, g0);
    alignas(16) float rec[rec_words(1)];
    float tex9[9] = {0};
    store_record<1>(rec, g0, tex9);
    load_record<1>(rec, g, tex9);
    Fragment fr;
    const bool ok = soft_fragment<1>(A, g, px[i], py[i], thr, fr);
    float* o = out + 10 * i;
    o[0] = ok ? 1.f : 0.f;
    o[1] = fr.w[0]; o[2] = fr.w[1]; o[3] = fr.w[2];
    o[4] = fr.t[0]; o[5] = fr.t[1]; o[6] = fr.t[2];
    o[7] = fr.dx; o[8] = fr.dy; o[9] = fr.dis;
    out_frag[i] = ok ? fr.frag : 0.f;
  }
}

extern "C" void lasr_host_stats(long long* inside, long long* all_edges, int reset) {
  *inside = lasr_host_inside_pairs; *all_edges = lasr_host_all_edges_pairs;
  if (reset) { lasr_host_inside_pairs = 0; lasr_host_all_edges_pairs = 0; }
}

// The forward kernel's own pair code (forward_pair<kFast = true, K = 1>: border test, record load, soft fragment, alpha
// product, clipped-barycentric depth, online depth softmax, colour blend) run on the host for ONE frame: every pixel
// visits the faces in ascending order, exactly what a warp lane does after the tile / footprint culls (which only
// skip faces the border test rejects anyway).  State initialisation and the final division mirror the kernel's
// init_state / finish.  out_rgba [4,S,S], out_aggr [2,S,S]; textures: vertex colours [F,9].
extern "C" void lasr_host_forward(const float* faces9, const float* tex9, int F, int S, float near_, float far_, float sigma,
                                  float gamma, float eps, float dist_eps, const float* bg3, float* out_rgba, float* out_aggr) {
  RasterArgs A{};
  A.N = 1; A.F = F; A.S = S; A.T = 3; A.K = 1; A.R = 1;
  A.dist = 2; A.rgb = 1; A.alpha = 2; A.sample = 1; A.double_side = 1;
  A.near_ = near_; A.far_ = far_; A.eps = eps; A.sigma = sigma; A.dist_eps = dist_eps; A.gamma = gamma;
  A.inv_sigma = (float)(1.0 / (double)sigma); A.inv_gamma = (float)(1.0 / (double)gamma);
  FwdConsts C;
  C.near_ = near_; C.far_ = far_; C.thr = dist_eps * sigma; C.inv_range = 1.f / (far_ - near_);
  const float sthr = sqrtf(dist_eps * sigma);
  float* recs = new float[(size_t)F * rec_words(1)];
  float4* bds = new float4[F];
  for (int f = 0; f < F; ++f) {
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    store_record<1>(recs + (size_t)f * rec_words(1), g, tex9 + 9 * f);
    bds[f] = make_float4(g.lox, g.hix, g.loy, g.hiy);
  }
  const size_t SS = (size_t)S * S;
  for (int py = 0; py < S; ++py)
    for (int px = 0; px < S; ++px) {
      const float xp = pixel_coord(px, S), yp = pixel_coord(S - 1 - py, S);      // row 0 is the top
      PixelState<1> st;
      st.alpha = 1.f; st.ssum = expf(eps / gamma); st.smax = eps; st.depth_min = 10000000.f; st.face_min = -1;
      for (int k = 0; k < 3; ++k) st.col[k] = bg3[k] * st.ssum;
      for (int f = 0; f < F; ++f)
        forward_pair<true, 1>(A, C, recs + (size_t)f * rec_words(1), bds[f], f, 0, xp, yp, st);
      const size_t pn = (size_t)py * S + px;
      for (int k = 0; k < 3; ++k) out_rgba[k * SS + pn] = st.col[k] / st.ssum;
      out_rgba[3 * SS + pn] = 1.f - st.alpha;
      out_aggr[pn] = st.ssum; out_aggr[SS + pn] = st.smax;
    }
  delete[] recs; delete[] bds;
}

// The backward kernel's own pixel-loop body (raster_backward_pair.inc, included textually exactly as the kernel includes
// it) run on the host for ONE frame: every face sweeps its pixel window, as a warp does, and sums the pairs'
// contributions (sequentially here; the kernel sums per lane and reduces in a fixed order).  LASR configuration
// (kFast), one colour set.  grad_faces [F,9], grad_tex [F,9].
extern "C" void lasr_host_backward(const float* faces9, const float* textures9, int F, int S, float near_, float far_,
                                   float sigma, float gamma, float eps, float dist_eps, const float* soft_colors,
                                   const float* aggrs, const float* grad_soft_colors, float* grad_faces, float* grad_tex) {
  constexpr bool kFast = true;
  constexpr int kMode = 1, K = 1, KF = 0, P = 3 * K + 1;
  constexpr bool kPacked = false; const float4* pk = nullptr; (void)pk;
  RasterArgs A{};
  A.N = 1; A.F = F; A.S = S; A.T = 3; A.K = 1; A.R = 1;
  A.dist = 2; A.rgb = 1; A.alpha = 2; A.sample = 1; A.double_side = 1;
  A.near_ = near_; A.far_ = far_; A.eps = eps; A.sigma = sigma; A.dist_eps = dist_eps; A.gamma = gamma;
  A.inv_sigma = (float)(1.0 / (double)sigma); A.inv_gamma = (float)(1.0 / (double)gamma);
  const float thr = dist_eps * sigma, sthr = sqrtf(thr);
  const float inv_range = 1.f / (far_ - near_), cz_scale = A.inv_gamma / (near_ - far_);
  const int dist_f = 2, rgb_f = 1, alpha_f = 2, sample = 1;
  (void)dist_f; (void)rgb_f; (void)alpha_f; (void)sample;
  const long long SS = (long long)S * S;
  const float *sc = soft_colors, *ag = aggrs, *gs = grad_soft_colors;
  for (int f = 0; f < F; ++f) {
    const long long wg = f;
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    const float* tex9 = textures9 + 9 * f;
    float* gtex_out = grad_tex + 9 * f;
    float gf[9] = {0}, gt[9] = {0};
    int xlo, xhi, ylo, yhi;
    coord_window(g.lox, g.hix, S, xlo, xhi);
    coord_window(g.loy, g.hiy, S, ylo, yhi);
    for (int yi = yhi; yi >= ylo && xlo <= xhi; --yi) {
      const int xa = xlo, xb = xhi;
      for (int xi = xa; xi <= xb; ++xi) {
        const float xp = pixel_coord(xi, S), yp = pixel_coord(yi, S);
        if (xp > g.hix || xp < g.lox || yp > g.hiy || yp < g.loy) continue;
        const long long pn = (long long)(S - 1 - yi) * S + xi;
#define PLANE(k) ((k) * SS + pn)
#include "raster_backward_pair.inc"
#undef PLANE
      }
    }
    for (int k = 0; k < 9; ++k) { grad_faces[9 * f + k] = gf[k]; grad_tex[9 * f + k] = gt[k]; }
  }
}

// ---- the generic instantiation (kFast = false: every operator mode, the reference's IEEE arithmetic) on the host ------
// One frame; textures [F,T,3] ('vertex': T = 3 = the three corner colours; 'surface': T = R*R texels).
static void fill_mode_args(RasterArgs& A, int F, int S, int T, float near_, float far_, float sigma, float gamma, float eps,
                           float dist_eps, int dist, int rgb, int alpha, int sample, int double_side, const float* textures) {
  A.N = 1; A.F = F; A.S = S; A.T = T; A.K = 1; A.R = (int)sqrt((double)T);
  A.dist = dist; A.rgb = rgb; A.alpha = alpha; A.sample = sample; A.double_side = double_side;
  A.near_ = near_; A.far_ = far_; A.eps = eps; A.sigma = sigma; A.dist_eps = dist_eps; A.gamma = gamma;
  A.inv_sigma = (float)(1.0 / (double)sigma); A.inv_gamma = (float)(1.0 / (double)gamma);
  A.textures = textures;
}

extern "C" void lasr_host_forward_modes(const float* faces9, const float* textures, int F, int S, int T, float near_,
                                        float far_, float sigma, float gamma, float eps, float dist_eps, int dist, int rgb,
                                        int alpha, int sample, int double_side, const float* bg3, float* out_rgba,
                                        float* out_aggr) {
  RasterArgs A{};
  fill_mode_args(A, F, S, T, near_, far_, sigma, gamma, eps, dist_eps, dist, rgb, alpha, sample, double_side, textures);
  FwdConsts C;
  C.near_ = near_; C.far_ = far_; C.thr = dist_eps * sigma; C.inv_range = 1.f / (far_ - near_);
  const float sthr = sqrtf(dist_eps * sigma);
  float* recs = new float[(size_t)F * rec_words(1)];
  float4* bds = new float4[F];
  for (int f = 0; f < F; ++f) {
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    store_record<1>(recs + (size_t)f * rec_words(1), g, sample == 1 ? textures + 9 * f : nullptr);
    bds[f] = make_float4(g.lox, g.hix, g.loy, g.hiy);
  }
  const size_t SS = (size_t)S * S;
  for (int py = 0; py < S; ++py)
    for (int px = 0; px < S; ++px) {
      const float xp = pixel_coord(px, S), yp = pixel_coord(S - 1 - py, S);
      PixelState<1> st;                                        // raster.cu: init_state
      st.alpha = alpha == 2 ? 1.f : 0.f; st.ssum = expf(eps / gamma); st.smax = eps; st.depth_min = 10000000.f; st.face_min = -1;
      for (int k = 0; k < 3; ++k) st.col[k] = rgb == 1 ? bg3[k] * st.ssum : bg3[k];
      for (int f = 0; f < F; ++f)
        forward_pair<false, 1>(A, C, recs + (size_t)f * rec_words(1), bds[f], f, 0, xp, yp, st);
      const size_t pn = (size_t)py * S + px;                   // raster.cu: finish
      out_rgba[3 * SS + pn] = alpha == 0 ? st.alpha : (alpha == 1 ? st.alpha / F : 1.f - st.alpha);
      if (rgb == 0) {
        for (int k = 0; k < 3; ++k) out_rgba[k * SS + pn] = st.face_min != -1 ? st.col[k] : bg3[k];
        out_aggr[pn] = st.depth_min; out_aggr[SS + pn] = (float)st.face_min;
      } else {
        for (int k = 0; k < 3; ++k) out_rgba[k * SS + pn] = st.col[k] / st.ssum;
        out_aggr[pn] = st.ssum; out_aggr[SS + pn] = st.smax;
      }
    }
  delete[] recs; delete[] bds;
}

#define atomicAdd(p, v) (*(p) += (v))
extern "C" void lasr_host_backward_modes(const float* faces9, const float* textures, int F, int S, int T, float near_,
                                         float far_, float sigma, float gamma, float eps, float dist_eps, int dist, int rgb,
                                         int alpha, int sample_, int double_side, const float* soft_colors,
                                         const float* aggrs, const float* grad_soft_colors, float* grad_faces,
                                         float* grad_tex) {
  constexpr bool kFast = false;
  constexpr int kMode = 0, K = 1, KF = 0, P = 3 * K + 1;
  constexpr bool kPacked = false; const float4* pk = nullptr; (void)pk;
  RasterArgs A{};
  fill_mode_args(A, F, S, T, near_, far_, sigma, gamma, eps, dist_eps, dist, rgb, alpha, sample_, double_side, textures);
  const float thr = dist_eps * sigma, sthr = sqrtf(thr);
  const float inv_range = 1.f / (far_ - near_), cz_scale = A.inv_gamma / (near_ - far_);
  (void)inv_range; (void)cz_scale;
  const int dist_f = dist, rgb_f = rgb, alpha_f = alpha, sample = sample_;
  const long long SS = (long long)S * S;
  const float *sc = soft_colors, *ag = aggrs, *gs = grad_soft_colors;
  for (int f = 0; f < F; ++f) {
    const long long wg = f;
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    float tex9[9] = {0};
    if (sample == 1) for (int k = 0; k < 9; ++k) tex9[k] = textures[9 * f + k];
    float* gtex_out = grad_tex + (size_t)f * T * 3;
    float gf[9] = {0}, gt[9] = {0};
    int xlo, xhi, ylo, yhi;
    coord_window(g.lox, g.hix, S, xlo, xhi);
    coord_window(g.loy, g.hiy, S, ylo, yhi);
    for (int yi = yhi; yi >= ylo && xlo <= xhi; --yi)
      for (int xi = xlo; xi <= xhi; ++xi) {
        const float xp = pixel_coord(xi, S), yp = pixel_coord(yi, S);
        if (xp > g.hix || xp < g.lox || yp > g.hiy || yp < g.loy) continue;
        const long long pn = (long long)(S - 1 - yi) * S + xi;
#define PLANE(k) ((k) * SS + pn)
#include "raster_backward_pair.inc"
#undef PLANE
      }
    for (int k = 0; k < 9; ++k) grad_faces[9 * f + k] = gf[k];
    if (sample == 1) for (int k = 0; k < 9; ++k) gtex_out[k] = gt[k];     // 'surface': scattered by the loop body itself
  }
}
#undef atomicAdd

// ---- several colour sets per geometry pass (K = 2, 3; LASR's configuration only) on the host -------------------------
// tex [F][9K]: set-major, nine corner-colour floats per set (what the setup kernel gathers into the record).
template <int K>
static void host_forward_sets(const float* faces9, const float* tex, int F, int S, float near_, float far_, float sigma,
                              float gamma, float eps, float dist_eps, const float* bg, float* out, float* out_aggr) {
  RasterArgs A{};
  fill_mode_args(A, F, S, 3, near_, far_, sigma, gamma, eps, dist_eps, 2, 1, 2, 1, 1, nullptr);
  A.K = K;
  FwdConsts C;
  C.near_ = near_; C.far_ = far_; C.thr = dist_eps * sigma; C.inv_range = 1.f / (far_ - near_);
  const float sthr = sqrtf(dist_eps * sigma);
  constexpr int W = rec_words(K), P = 3 * K + 1;
  float* recs = new float[(size_t)F * W];
  float4* bds = new float4[F];
  for (int f = 0; f < F; ++f) {
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    store_record<K>(recs + (size_t)f * W, g, tex + (size_t)9 * K * f);
    bds[f] = make_float4(g.lox, g.hix, g.loy, g.hiy);
  }
  const size_t SS = (size_t)S * S;
  for (int py = 0; py < S; ++py)
    for (int px = 0; px < S; ++px) {
      const float xp = pixel_coord(px, S), yp = pixel_coord(S - 1 - py, S);
      PixelState<K> st;
      st.alpha = 1.f; st.ssum = expf(eps / gamma); st.smax = eps; st.depth_min = 10000000.f; st.face_min = -1;
      for (int k = 0; k < 3 * K; ++k) st.col[k] = bg[k] * st.ssum;
      for (int f = 0; f < F; ++f) forward_pair<true, K>(A, C, recs + (size_t)f * W, bds[f], f, 0, xp, yp, st);
      const size_t pn = (size_t)py * S + px;
      for (int k = 0; k < 3 * K; ++k) out[k * SS + pn] = st.col[k] / st.ssum;
      out[(P - 1) * SS + pn] = 1.f - st.alpha;
      out_aggr[pn] = st.ssum; out_aggr[SS + pn] = st.smax;
    }
  delete[] recs; delete[] bds;
}

extern "C" int lasr_host_forward_sets(const float* faces9, const float* tex, int F, int S, int K, float near_, float far_,
                                      float sigma, float gamma, float eps, float dist_eps, const float* bg, float* out,
                                      float* out_aggr) {
  if (K == 1) host_forward_sets<1>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, bg, out, out_aggr);
  else if (K == 2) host_forward_sets<2>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, bg, out, out_aggr);
  else if (K == 3) host_forward_sets<3>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, bg, out, out_aggr);
  else return 1;
  return 0;
}

// kPacked: what a pair reads from its pixel is first packed per pixel the way raster_pack_kernel (one colour set) and
// head_backward_kernel<packed> (three sets, the first detached) lay it out on the GPU -- planes [1 + ceil(3 live / 4)][S*S]
// of float4: { g_alpha (1 - alpha), sum over live colour planes of g_k image_k, 1 / softmax sum, softmax max }, then the
// live g_k -- and the pair code runs its packed branch.
template <int K, int KF, bool kPacked = false>
static void host_backward_sets(const float* faces9, const float* tex, int F, int S, float near_, float far_, float sigma,
                               float gamma, float eps, float dist_eps, const float* soft_colors, const float* aggrs,
                               const float* grad_soft_colors, float* grad_faces, float* grad_tex) {
  constexpr bool kFast = true;
  constexpr int kMode = 1, P = 3 * K + 1;
  constexpr int kPlanes = 1 + (3 * (K - KF) + 3) / 4;
  float4* pk_buf = nullptr;
  if (kPacked) {
    const size_t SSz = (size_t)S * S;
    pk_buf = new float4[(size_t)kPlanes * SSz];
    for (size_t p = 0; p < SSz; ++p) {
      float gk[4 * (kPlanes - 1)] = {0};
      float T = 0.f;
      for (int k = 3 * KF; k < 3 * K; ++k) {
        gk[k - 3 * KF] = grad_soft_colors[k * SSz + p];
        T += gk[k - 3 * KF] * soft_colors[k * SSz + p];
      }
      pk_buf[p] = make_float4(grad_soft_colors[(P - 1) * SSz + p] * (1.f - soft_colors[(P - 1) * SSz + p]), T,
                              fast_rcp<kMode>(aggrs[p]), aggrs[SSz + p]);
      for (int j = 0; j < kPlanes - 1; ++j)
        pk_buf[(1 + j) * SSz + p] = make_float4(gk[4 * j], gk[4 * j + 1], gk[4 * j + 2], gk[4 * j + 3]);
    }
  }
  const float4* pk = pk_buf; (void)pk;
  RasterArgs A{};
  fill_mode_args(A, F, S, 3, near_, far_, sigma, gamma, eps, dist_eps, 2, 1, 2, 1, 1, nullptr);
  A.K = K;
  const float thr = dist_eps * sigma, sthr = sqrtf(thr);
  const float inv_range = 1.f / (far_ - near_), cz_scale = A.inv_gamma / (near_ - far_);
  const int dist_f = 2, rgb_f = 1, alpha_f = 2, sample = 1;
  (void)dist_f; (void)rgb_f; (void)alpha_f; (void)sample;
  const long long SS = (long long)S * S;
  const float *sc = soft_colors, *ag = aggrs, *gs = grad_soft_colors;
  for (int f = 0; f < F; ++f) {
    const long long wg = f;
    FaceGeom g;
    face_setup(faces9 + 9 * f, sthr, g);
    const float* tex9 = tex + (size_t)9 * K * f;
    float* gtex_out = grad_tex + (size_t)9 * K * f + 9 * KF;
    float gf[9] = {0}, gt[9 * (K - KF)] = {0};
    int xlo, xhi, ylo, yhi;
    coord_window(g.lox, g.hix, S, xlo, xhi);
    coord_window(g.loy, g.hiy, S, ylo, yhi);
    for (int yi = yhi; yi >= ylo && xlo <= xhi; --yi)
      for (int xi = xlo; xi <= xhi; ++xi) {
        const float xp = pixel_coord(xi, S), yp = pixel_coord(yi, S);
        if (xp > g.hix || xp < g.lox || yp > g.hiy || yp < g.loy) continue;
        const long long pn = (long long)(S - 1 - yi) * S + xi;
#define PLANE(k) ((k) * SS + pn)
#include "raster_backward_pair.inc"
#undef PLANE
      }
    for (int k = 0; k < 9; ++k) grad_faces[9 * f + k] = gf[k];
    for (int k = 0; k < 9 * (K - KF); ++k) gtex_out[k] = gt[k];
  }
  delete[] pk_buf;
}

// the packed branch of the pair code for the two instantiations the library launches it with: K = 1 (raster_pack_kernel
// feeding the single-set backward) and K = 3 with the first set detached (the fused region)
extern "C" int lasr_host_backward_packed(const float* faces9, const float* tex, int F, int S, int K, int KF, float near_,
                                         float far_, float sigma, float gamma, float eps, float dist_eps,
                                         const float* soft_colors, const float* aggrs, const float* grad_soft_colors,
                                         float* grad_faces, float* grad_tex) {
  if (K == 1 && KF == 0)
    host_backward_sets<1, 0, true>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, soft_colors, aggrs,
                                   grad_soft_colors, grad_faces, grad_tex);
  else if (K == 3 && KF == 1)
    host_backward_sets<3, 1, true>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, soft_colors, aggrs,
                                   grad_soft_colors, grad_faces, grad_tex);
  else
    return 1;
  return 0;
}

extern "C" int lasr_host_backward_sets(const float* faces9, const float* tex, int F, int S, int K, int KF, float near_,
                                       float far_, float sigma, float gamma, float eps, float dist_eps,
                                       const float* soft_colors, const float* aggrs, const float* grad_soft_colors,
                                       float* grad_faces, float* grad_tex) {
#define LASR_SETS_CASE(k, kf)                                                                                          \
  if (K == k && KF == kf) {                                                                                            \
    host_backward_sets<k, kf>(faces9, tex, F, S, near_, far_, sigma, gamma, eps, dist_eps, soft_colors, aggrs,         \
                              grad_soft_colors, grad_faces, grad_tex);                                                 \
    return 0;                                                                                                          \
  }
  LASR_SETS_CASE(1, 0) LASR_SETS_CASE(2, 0) LASR_SETS_CASE(2, 1) LASR_SETS_CASE(3, 0) LASR_SETS_CASE(3, 1)
#undef LASR_SETS_CASE
  return 1;
}

// coord_window on n intervals: the inclusive pixel-index windows both kernels cull with (setup kernel -> face_window; the
// backward sweeps them).  They must contain every pixel whose centre passes the exact float predicate lo <= c <= hi.
extern "C" void lasr_host_coord_windows(const float* lo, const float* hi, int n, int S, int* out_lo_hi) {
  for (int i = 0; i < n; ++i) coord_window(lo[i], hi[i], S, out_lo_hi[2 * i], out_lo_hi[2 * i + 1]);
}
extern "C" void lasr_host_pixel_coords(int S, float* out) {
  for (int i = 0; i < S; ++i) out[i] = pixel_coord(i, S);
}

extern "C" int lasr_host_fast_inside(void) { return LASR_FAST_INSIDE; }

// Compact record round trip: face_setup -> store_compact -> face_from_compact must give back the FaceGeom the forward
// worked with, field for field and bit for bit (det excepted: only face_setup uses it).  Returns the number of faces
// with any differing field.
extern "C" int lasr_host_compact_roundtrip(const float* faces9, int F, float sthr) {
  int bad = 0;
  for (int f = 0; f < F; ++f) {
    FaceGeom a, b;
    face_setup(faces9 + 9 * f, sthr, a);
    alignas(16) float rec[kCompactWords];
    store_compact(rec, a);
    float cr[kCompactWords];
    load_compact(rec, cr);
    face_from_compact(faces9 + 9 * f, sthr, cr, b);
    bool same = a.obt == b.obt && a.front == b.front;
    same = same && memcmp(a.inv, b.inv, sizeof a.inv) == 0 && memcmp(a.x, b.x, sizeof a.x) == 0 &&
           memcmp(a.y, b.y, sizeof a.y) == 0 && memcmp(a.z, b.z, sizeof a.z) == 0 && memcmp(a.iz, b.iz, sizeof a.iz) == 0 &&
           memcmp(a.a, b.a, sizeof a.a) == 0 && memcmp(a.rden, b.rden, sizeof a.rden) == 0 &&
           memcmp(&a.lox, &b.lox, sizeof(float)) == 0 && memcmp(&a.hix, &b.hix, sizeof(float)) == 0 &&
           memcmp(&a.loy, &b.loy, sizeof(float)) == 0 && memcmp(&a.hiy, &b.hiy, sizeof(float)) == 0 &&
           memcmp(&a.kappa, &b.kappa, sizeof(float)) == 0;
    bad += same ? 0 : 1;
  }
  return bad;
}
