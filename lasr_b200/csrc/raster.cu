// raster.cu -- B200 (sm_100a) soft rasterizer: forward (tile-binned, pixel-parallel) and backward
// (face-parallel, atomic-free) kernels plus their C-ABI entry points.
//
// Replaces third_party/softras/soft_renderer/cuda/soft_rasterize_cuda_kernel.cu:245-813 of the
// reference.  The reference launches one thread per pixel and walks ALL F faces per pixel
// (N*S^2*F iterations, .cu:370,538) and scatters gradients with 18 float atomics per surviving
// (pixel,face) pair (.cu:620-622,657-666).  Here:
//
//   setup     one thread per face: derived per-face constants (a 176-byte record), the dilated
//             bounding box and its packed pixel window (8 B), written once to the workspace -- the records as ONE bulk
//             store per warp (shared -> global through the TMA unit); for large launches also a per-tile face count,
//             from which a one-CTA counting sort lists the (frame, tile) items heaviest first -- the order the
//             forward's CTAs start in;
//   forward   one CTA per 16x16 pixel tile: (1) the CTA scans the packed windows and compacts the
//             indices of overlapping faces, in ascending order, into shared memory; (2) the list is consumed in rounds
//             of 64 faces through a two-deep shared-memory ring filled by cp.async.bulk copies of the faces' records and
//             boxes that complete on mbarriers (LASR configuration); (3) each warp (an 8x4 pixel footprint) walks the
//             staged round on its own, culls 32 faces at a time against its footprint with one ballot and evaluates
//             only the survivors from LDS broadcasts.  Faces are visited in ascending index, the order the reference's
//             per-pixel loop uses, so the online depth-softmax and the hard z-test see the same sequence;
//   backward  one warp per face, a CTA owning a run of consecutive faces: the warp sweeps the face's own (tight) pixel
//             window, every lane owns a pixel, the 18 partial derivatives are reduced with a fixed-order warp
//             butterfly and written once.  No atomics, deterministic.  What a pair needs from its pixel arrives either
//             as ten plane gathers or, packed per pixel by the fused head / raster_pack_kernel, as two or three
//             coalesced 128-bit loads.
//
// Nothing here is GEMM-shaped: no tensor cores.  The kernels are FP32-issue/MUFU bound after
// binning; HBM traffic is the compulsory 72F+24S^2 / 144F+40S^2 bytes per frame
// plus the per-face scratch, which stays in L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <string>

#include "lasr_b200.h"
#include "host_common.h"
#include "raster_common.cuh"

namespace lasr {

#ifndef LASR_TILE_H
#define LASR_TILE_H 16
#endif
constexpr int kTileW = 16;                      // tile width in pixels (two 8x4 warp footprints)
constexpr int kTileH = LASR_TILE_H;             // tile height in pixels (multiple of 4)
constexpr int kFwdWarps = 2 * (kTileH / 4);     // each warp owns an 8x4 pixel footprint
#ifndef LASR_FWD_LPT                // dispatch forward tiles heaviest-first (longest-processing-time order)
#define LASR_FWD_LPT 1
#endif
constexpr int kFwdThreads = 32 * kFwdWarps;
#ifndef LASR_SCAN_UNROLL            // window loads kept in flight per thread in the tile scan
#define LASR_SCAN_UNROLL 2
#endif
constexpr int kScanUnroll = LASR_SCAN_UNROLL;
constexpr int kMaxScanChunk = 12032;  // faces scanned per round (47 KB of indices, under the 48 KB default limit with the static smem); one round if F fits
#ifndef LASR_FWD_MIN_BLOCKS
#define LASR_FWD_MIN_BLOCKS 4
#endif
#ifndef LASR_BWD_MIN_BLOCKS          // resident CTAs of LASR_BWD_THREADS threads the register budget must allow
#define LASR_BWD_MIN_BLOCKS 4
#endif
#ifndef LASR_FWD_MIN_BLOCKS_MULTI   // several colour sets per pass carry more live state per pixel
#define LASR_FWD_MIN_BLOCKS_MULTI 3
#endif
#ifndef LASR_BWD_MIN_BLOCKS_MULTI   // three live colour sets (36 accumulators): 168 registers instead of spilling at 128
#define LASR_BWD_MIN_BLOCKS_MULTI 3  // (K = 3 backward 2.05 -> 1.95 ms)
#endif
#ifndef LASR_BWD_MIN_BLOCKS_LIVE2   // two live sets (LASR's hot path: K = 3 with the frame's own positions detached):
#define LASR_BWD_MIN_BLOCKS_LIVE2 4 // 128 registers and a few spilled words beat 168 registers at 12 warps per SM
#endif
#ifndef LASR_PREFETCH_REC           // warm L1 with a face's record when it passes the warp cull
#define LASR_PREFETCH_REC 1
#endif
#ifndef LASR_BWD_THREADS
#define LASR_BWD_THREADS 128
#endif
#ifndef LASR_FTZ_MUFU_BWD           // ... but not in the backward (measured slower)
#define LASR_FTZ_MUFU_BWD 0
#endif
#ifndef LASR_FWD_BULK                // forward: stage the tile's face records + boxes in shared memory with cp.async.bulk + mbarrier
#define LASR_FWD_BULK 1
#endif
#ifndef LASR_BWD_PACK_K1             // single-set backward (LASR configuration): pack the per-pixel gradients first (raster_pack_kernel)
#define LASR_BWD_PACK_K1 1
#endif
#ifndef LASR_BWD_FACES_PER_WARP     // consecutive faces a backward warp works through, single colour set (see raster_backward_kernel)
#define LASR_BWD_FACES_PER_WARP 4
#endif
__host__ __device__ constexpr int bwd_faces_per_warp(int K) { return K == 1 ? LASR_BWD_FACES_PER_WARP : 1; }
#ifndef LASR_EXP_PLANES_L1
#define LASR_EXP_PLANES_L1 0
#endif
#ifndef LASR_BWD_INDEX32_MULTI      // ... and in the multi-set backward
#define LASR_BWD_INDEX32_MULTI 0
#endif
#ifndef LASR_BWD_INDEX32            // 32-bit per-frame plane offsets in the backward's pixel loop
#define LASR_BWD_INDEX32 1
#endif

// Blackwell / Hopper async-proxy primitives for the bulk-staged forward (LASR_FWD_BULK): mbarrier objects in shared
// memory, 1-D bulk copies global -> shared that complete on them (the TMA unit moves the bytes, no registers, no
// per-lane load instructions), parity waits.
constexpr int kStageSlots = 64;                 // face records per staging buffer (two cull steps of a warp)
constexpr int kStageBufs = 2;
__host__ __device__ constexpr int stage_bytes(int K) { return kStageSlots * (rec_words(K) * 4 + 16); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const uint32_t a = smem_u32(bar);
  unsigned done = 0, spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();     // a protocol bug must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk store of `bytes` (16-byte granular) by ONE thread; the caller orders its generic-proxy writes to
// the source with fence.proxy.async first and must not leave before the read of shared memory has completed
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// setup: one thread per face -> packed pixel window, dilated border box, full face record;
// the first S threads also fill the pixel-centre coordinate table.
// ---------------------------------------------------------------------------------------------
#ifndef LASR_SETUP_BULK_STORE        // records leave the setup kernel as one bulk store per warp (else 11-15 strided 16-byte stores per thread)
#define LASR_SETUP_BULK_STORE 1
#endif
__global__ void __launch_bounds__(128) raster_setup_kernel(RasterArgs A, int with_records) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nf = (long long)A.N * A.F;
  if (i < A.S) A.coords[i] = pixel_coord((int)i, A.S);
  // A thread's record is 176 - 240 contiguous bytes, so a warp's 32 records are one contiguous 5.5 - 7.5 KB run of the
  // workspace: they are laid down in shared memory and ONE lane hands the run to the TMA unit (cp.async.bulk
  // shared -> global).  Written per thread, every 16-byte store of a warp instruction lands in a different 176-byte
  // stride (32 half-used sectors per instruction).  Lanes past the last face of the last warp stay for the warp-level
  // steps but write nothing.
  const bool staged = LASR_SETUP_BULK_STORE != 0 && with_records != 0;
  __shared__ __align__(16) float s_rec[LASR_SETUP_BULK_STORE ? 4 : 1][LASR_SETUP_BULK_STORE ? 32 * rec_words(kMaxSets) : 4];
  const long long warp_first = i - (threadIdx.x & 31);
  if (staged ? warp_first >= nf : i >= nf) return;
  const bool live_thread = i < nf;
  const long long fi = live_thread ? i : nf - 1;
  const bool vertex_tex = A.sample == 1;
  float fv[9], tex9[9 * kMaxSets];
  fetch_face(A, (int)(fi / A.F), (int)(fi % A.F), fv, tex9, vertex_tex && with_records);
  const float sthr = sqrtf(A.dist_eps * A.sigma);
  FaceGeom g;
  face_setup(fv, sthr, g);
  int xlo, xhi, ylo, yhi;
  coord_window(g.lox, g.hix, A.S, xlo, xhi);
  coord_window(g.loy, g.hiy, A.S, ylo, yhi);
  if (xlo > xhi || ylo > yhi || !live_thread) { xlo = 1; xhi = 0; ylo = 1; yhi = 0; }
  if (live_thread) A.face_window[i] = make_uint2((uint32_t)xlo | ((uint32_t)xhi << 16), (uint32_t)ylo | ((uint32_t)yhi << 16));
  if (!with_records) return;
  {  // union of the frame's face windows (lets empty tiles skip their scan): one atomic per warp, not per face
    const int frame = (int)(fi / A.F);
    const bool live = xlo <= xhi;
    const unsigned act = __activemask();
    const int f0 = __shfl_sync(act, frame, __ffs(act) - 1);
    const bool uniform = __all_sync(act, frame == f0);
    int a = live ? xlo : 0x7f7f7f7f, b = live ? -xhi : 0x7f7f7f7f, c = live ? ylo : 0x7f7f7f7f, d = live ? -yhi : 0x7f7f7f7f;
    int* fw = A.frame_window + 4 * frame;
    if (uniform) {
      a = __reduce_min_sync(act, a); b = __reduce_min_sync(act, b);
      c = __reduce_min_sync(act, c); d = __reduce_min_sync(act, d);
      if ((threadIdx.x & 31) == __ffs(act) - 1 && a != 0x7f7f7f7f) {
        atomicMin(fw + 0, a); atomicMin(fw + 1, b); atomicMin(fw + 2, c); atomicMin(fw + 3, d);
      }
    } else if (live) {
      atomicMin(fw + 0, a); atomicMin(fw + 1, b); atomicMin(fw + 2, c); atomicMin(fw + 3, d);
    }
  }
  if (A.tile_count && xlo <= xhi) {   // work estimate per forward tile: faces whose window reaches it
    const int tiles_x = (A.S + kTileW - 1) / kTileW, tiles_y = (A.S + kTileH - 1) / kTileH;
    int* tc = A.tile_count + (long long)(fi / A.F) * tiles_x * tiles_y;
    const int tx0 = xlo / kTileW, tx1 = xhi / kTileW;
    const int ty0 = (A.S - 1 - yhi) / kTileH, ty1 = (A.S - 1 - ylo) / kTileH;   // rows from the top
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(tc + ty * tiles_x + tx, 1);
  }
  if (live_thread) A.face_border[i] = make_float4(g.lox, g.hix, g.loy, g.hiy);
  if (A.compact_record && live_thread) store_compact(A.compact_record + i * kCompactWords, g);   // what the backward keeps
  const int words = rec_words(A.K);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* dst = staged ? s_rec[warp] + lane * words : A.face_record + fi * words;
  if (A.K == 1) store_record<1>(dst, g, vertex_tex ? tex9 : nullptr);
  else if (A.K == 2) store_record<2>(dst, g, tex9);
  else store_record<3>(dst, g, tex9);
  if (staged) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
      const long long left = nf - warp_first;
      const unsigned cnt = (unsigned)(left < 32 ? left : 32);
      bulk_s2g(A.face_record + warp_first * words, smem_u32(s_rec[warp]), cnt * (unsigned)words * 4u);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dispatch order of the forward's CTAs: (frame, tile) items sorted by their face count, heaviest first
// (counting sort, one CTA).  CTAs are started in block-index order and a tile on the object lives for tens of
// microseconds, so with frame-major order the SMs idled through a long tail (ncu: active 81 % of the launch);
// longest-first dispatch is the classic remedy.  Only the ORDER of independent tiles changes -- ties between
// equal counts are broken by atomics, which cannot affect any output.
// ---------------------------------------------------------------------------------------------
constexpr int kOrderBuckets = 1024;
__global__ void __launch_bounds__(1024) raster_order_kernel(const int* __restrict__ count, int* __restrict__ order,
                                                            int items) {
  __shared__ int s_hist[kOrderBuckets];
  __shared__ int s_zero;                        // background tiles (count 0, half of all items): one ballot-aggregated
                                                // atomic per warp instead of thousands on the same bucket
  const int lane = threadIdx.x & 31;
  for (int b = threadIdx.x; b < kOrderBuckets; b += blockDim.x) s_hist[b] = 0;
  if (threadIdx.x == 0) s_zero = 0;
  __syncthreads();
  const int padded = (items + 31) & ~31;        // whole warps stay in the loop together (ballot)
  for (int i = threadIdx.x; i < padded; i += blockDim.x) {
    const int c = i < items ? count[i] : -1;
    const unsigned z = __ballot_sync(0xffffffffu, c == 0);
    if (c > 0) atomicAdd(&s_hist[min(c, kOrderBuckets - 1)], 1);
    else if (z && lane == __ffs(z) - 1) atomicAdd(&s_zero, __popc(z));
  }
  __syncthreads();
  if (threadIdx.x < 32) {   // exclusive prefix over buckets in DESCENDING count order, one warp
    int carry = 0;
    for (int base = kOrderBuckets - 32; base >= 0; base -= 32) {
      const int b = base + 31 - (int)threadIdx.x;           // lane 0 takes the largest count of this group
      const int h = s_hist[b];
      int incl = h;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)threadIdx.x >= o) incl += v;
      }
      s_hist[b] = carry + incl - h;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) s_zero = carry;       // the empty tiles follow all the others
  }
  __syncthreads();
  for (int i = threadIdx.x; i < padded; i += blockDim.x) {
    const int c = i < items ? count[i] : -1;
    const unsigned z = __ballot_sync(0xffffffffu, c == 0);
    if (c > 0) {
      order[atomicAdd(&s_hist[min(c, kOrderBuckets - 1)], 1)] = i;
    } else if (c == 0) {                        // exactly the lanes of z
      const int leader = __ffs(z) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&s_zero, __popc(z));
      base = __shfl_sync(z, base, leader);
      order[base + __popc(z & ((1u << lane) - 1))] = i;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
#ifndef LASR_FWD_CENTRE_OUT
#define LASR_FWD_CENTRE_OUT 1
#endif
// r-th element of the sequence c, c+1, c-1, c+2, ... over [0, n), c = (n-1)/2: the middle first, the ends last
__device__ __forceinline__ int centre_out(int r, int n) {
  const int c = (n - 1) >> 1, k = (r + 1) >> 1;
  return (r & 1) ? c + k : c - k;
}

// One CTA per 16 x kTileH tile, warp w owns footprint w for the whole kernel; the face list may be built in several
// rounds (F > kMaxScanChunk), the per-pixel state living in registers across them.
template <bool kFast, int K>
__global__ void __launch_bounds__(kFwdThreads, K == 1 ? LASR_FWD_MIN_BLOCKS : LASR_FWD_MIN_BLOCKS_MULTI) raster_forward_kernel(RasterArgs A, int chunk) {
  constexpr int kMode = kFast ? (LASR_FTZ_MUFU_FWD ? 2 : 1) : 0;
  constexpr int kRecWords = rec_words(K);
  constexpr int P = 3 * K + 1;      // image planes per frame: K colour sets, then alpha
  constexpr int kTW = kTileW, kTH = kTileH;
  constexpr int kFpX = kTW / 8;
  extern __shared__ __align__(16) int s_list[];   // [chunk] face indices, 8 per-warp segments of chunk/8 (+ staging buffers)
  __shared__ int s_cnt[kFwdWarps];
  const int wseg = chunk / kFwdWarps;
  constexpr bool kBulk = kFast && LASR_FWD_BULK != 0;
  __shared__ uint64_t s_full[kStageBufs], s_empty[kStageBufs];
  if (kBulk) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int b = 0; b < kStageBufs; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], kFwdWarps); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }

  const int S = A.S;
  const int tiles_x = (S + kTW - 1) / kTW, tiles_y = (S + kTH - 1) / kTH;
  int n, tile_x, tile_y;
  if (A.tile_order) {   // heaviest tiles first (raster_order_kernel)
    const int item = A.tile_order[blockIdx.x];
    n = item / (tiles_x * tiles_y);
    tile_x = (item % (tiles_x * tiles_y)) % tiles_x;
    tile_y = (item % (tiles_x * tiles_y)) / tiles_x;
  } else {
#if LASR_FWD_CENTRE_OUT
  // CTAs are dispatched in block-index order and live long (tens of microseconds for a tile on the object), so
  // whatever is dispatched last decides how long the other SMs idle at the end of the launch (ncu: SMs active
  // 81 % of the kernel with frame-major order).  Order: frame fastest, tiles from the image centre outwards --
  // the last CTAs of the grid are then the border tiles of all frames, which are mostly background.
  n = blockIdx.x % A.N;
  const int rank = blockIdx.x / A.N;
  tile_x = centre_out(rank % tiles_x, tiles_x);
  tile_y = centre_out(rank / tiles_x, tiles_y);
#else
  n = blockIdx.x / (tiles_x * tiles_y);
  tile_x = (blockIdx.x % (tiles_x * tiles_y)) % tiles_x;
  tile_y = (blockIdx.x % (tiles_x * tiles_y)) / tiles_x;
#endif
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long SS = (long long)S * S;

  FwdConsts C;
  C.near_ = A.near_; C.far_ = A.far_;
  if (A.near_far_dev) { C.near_ = A.near_far_dev[0]; C.far_ = A.near_far_dev[1]; }
  C.thr = A.dist_eps * A.sigma;
  C.inv_range = 1.f / (C.far_ - C.near_);
  const int rgb_f = kFast ? 1 : A.rgb;
  const int alpha_f = kFast ? 2 : A.alpha;

  // tile window in pixel indices (x index, y index = S-1-row)
  const int tx0 = tile_x * kTW, tx1 = tx0 + kTW - 1;
  const int ty1 = S - 1 - tile_y * kTH, ty0 = ty1 - (kTH - 1);

  const uint2* win = A.face_window + (long long)n * A.F;
  const float4* border = A.face_border + (long long)n * A.F;
  const float* records = A.face_record + (long long)n * A.F * kRecWords;

  // tiles that no face window of this frame reaches skip the scan and go straight to the epilogue
  const int4 fwin = *reinterpret_cast<const int4*>(A.frame_window + 4 * n);
  const bool tile_live = tx1 >= fwin.x && tx0 <= -fwin.y && ty1 >= fwin.z && ty0 <= -fwin.w;

  // (1) scan: warp w owns faces [base + w*wseg, base + (w+1)*wseg) and compacts hits in order
  auto scan = [&](int base) {
    int cnt = 0;
    const int seg0 = base + warp * wseg;
#pragma unroll kScanUnroll
    for (int it = 0; it < wseg; it += 32) {
      const int f = seg0 + it + lane;
      bool hit = false;
      if (f < A.F) {
        const uint2 wv = win[f];
        const int xlo = wv.x & 0xffff, xhi = wv.x >> 16, ylo = wv.y & 0xffff, yhi = wv.y >> 16;
        hit = xlo <= tx1 && xhi >= tx0 && ylo <= ty1 && yhi >= ty0 && xlo <= xhi;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) s_list[warp * wseg + cnt + __popc(m & ((1u << lane) - 1))] = f;
      cnt += __popc(m);
    }
    if (lane == 0) s_cnt[warp] = cnt;
  };

  // one 8x4 footprint: lane = pixel
  struct Footprint {
    int px, py;
    bool valid;
    float xp, yp, fx0, fx1, fy0, fy1;
    long long pn;
  };
  auto place = [&](int fp) {
    Footprint q;
    const int wx0 = tile_x * kTW + (fp % kFpX) * 8, wy0 = tile_y * kTH + (fp / kFpX) * 4;
    q.px = wx0 + (lane & 7);
    q.py = wy0 + (lane >> 3);  // row from the top
    q.valid = q.px < S && q.py < S;
    q.xp = A.coords[min(q.px, S - 1)];
    q.yp = A.coords[max(S - 1 - q.py, 0)];
    q.pn = (long long)q.py * S + q.px;
    // footprint in pixel-centre coordinates (for the per-warp cull)
    q.fx0 = A.coords[min(wx0, S - 1)];
    q.fx1 = A.coords[min(wx0 + 7, S - 1)];
    q.fy1 = A.coords[max(S - 1 - wy0, 0)];
    q.fy0 = A.coords[max(S - 1 - (wy0 + 3), 0)];
    return q;
  };
  auto init_state = [&](const Footprint& q, PixelState<K>& st) {
    st.alpha = alpha_f == 2 ? 1.f : 0.f;
    st.ssum = expf(A.eps / A.gamma);
    st.smax = A.eps;
    st.depth_min = 10000000.f;
    st.face_min = -1;
#pragma unroll
    for (int k = 0; k < 3 * K; ++k) {
      const float bg = A.bg_in_args ? A.bg[k] : (q.valid ? A.soft_colors[((long long)n * P + k) * SS + q.pn] : 0.f);
      st.col[k] = rgb_f == 1 ? bg * st.ssum : bg;
    }
  };
  // (2) a warp walks the whole ordered list on its own: cull 32 faces per ballot against its 8x4 footprint,
  //     evaluate survivors in ascending face order.  No barriers.
  auto walk = [&](const Footprint& q, PixelState<K>& st) {
#pragma unroll 1
    for (int seg = 0; seg < kFwdWarps; ++seg) {
      const int nseg = s_cnt[seg];
      const int* lst = s_list + seg * wseg;
#pragma unroll 1
      for (int c = 0; c < nseg; c += 32) {
        const int j = c + lane;
        int f = 0;
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        bool pass = false;
        if (j < nseg) {
          f = lst[j];
          bd = __ldg(border + f);
          pass = !(q.fx0 > bd.y || q.fx1 < bd.x || q.fy0 > bd.w || q.fy1 < bd.z);
#if LASR_PREFETCH_REC
          if (pass) {
            const char* rp = reinterpret_cast<const char*>(records + (long long)f * kRecWords);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rp));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + 128));
            if (kRecWords * 4 > 208) asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + kRecWords * 4 - 16));
          }
#endif
        }
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          const int fb = __shfl_sync(0xffffffffu, f, b);
          float4 bb;
          bb.x = __shfl_sync(0xffffffffu, bd.x, b);
          bb.y = __shfl_sync(0xffffffffu, bd.y, b);
          bb.z = __shfl_sync(0xffffffffu, bd.z, b);
          bb.w = __shfl_sync(0xffffffffu, bd.w, b);
          if (q.valid) forward_pair<kFast, K>(A, C, records + (long long)fb * kRecWords, bb, fb, n, q.xp, q.yp, st);
        }
      }
    }
  };
  // (2') LASR_FWD_BULK: the same walk over records and boxes STAGED IN SHARED MEMORY by the TMA unit.  The tile's list is
  //      consumed in rounds of kStageSlots faces through a two-deep ring: warp 0 issues, per face, one bulk copy of the
  //      record and one of the box (cp.async.bulk ... mbarrier::complete_tx) into the free buffer and arms its `full`
  //      barrier with the byte count; every warp waits on `full`, culls and evaluates from LDS broadcasts, and arrives on
  //      the buffer's `empty` barrier; warp 0 refills a buffer once all eight warps have released it.  No __syncthreads
  //      in the loop, no global loads in the walk, round r + 1 in flight while round r is evaluated.
  unsigned stage_it = 0;                       // staging rounds issued so far by this CTA (uniform): buffer = it & 1
  auto walk_bulk = [&](const Footprint& q, PixelState<K>& st) {
    constexpr unsigned kRecBytes = kRecWords * 4;
    const uint32_t stage0 = smem_u32(s_list) + (uint32_t)chunk * 4u;
    int total = 0;
#pragma unroll
    for (int w = 0; w < kFwdWarps; ++w) total += s_cnt[w];
    const int rounds = (total + kStageSlots - 1) / kStageSlots;
    auto produce = [&](int r) {                // warp 0: fill buffer (stage_it + r) & 1 with round r of the list
      const unsigned it = stage_it + (unsigned)r, b = it & 1u, u = it >> 1;
      if (u >= 1) mbar_wait(&s_empty[b], (u - 1) & 1u);
      const int n_r = min(kStageSlots, total - r * kStageSlots);
      if (lane == 0) mbar_expect_tx(&s_full[b], (unsigned)n_r * (kRecBytes + 16u));
      __syncwarp();
      const uint32_t sbox = stage0 + b * (uint32_t)stage_bytes(K), srec = sbox + kStageSlots * 16u;
      for (int j = lane; j < n_r; j += 32) {
        const int flat = r * kStageSlots + j;
        int seg = 0, seg_first = 0, run = 0;      // which scanning warp's segment holds flat entry `flat`
#pragma unroll
        for (int w = 0; w < kFwdWarps - 1; ++w) {
          run += s_cnt[w];
          if (flat >= run) { seg = w + 1; seg_first = run; }
        }
        const int f = s_list[seg * wseg + flat - seg_first];
        bulk_g2s(srec + (uint32_t)j * kRecBytes, records + (long long)f * kRecWords, kRecBytes, &s_full[b]);
        bulk_g2s(sbox + (uint32_t)j * 16u, border + f, 16u, &s_full[b]);
      }
    };
    if (warp == 0 && rounds > 0) produce(0);
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
      if (warp == 0 && r + 1 < rounds) produce(r + 1);
      const unsigned it = stage_it + (unsigned)r, b = it & 1u, u = it >> 1;
      mbar_wait(&s_full[b], u & 1u);
      const int n_r = min(kStageSlots, total - r * kStageSlots);
      const uint32_t sbox = stage0 + b * (uint32_t)stage_bytes(K), srec = sbox + kStageSlots * 16u;
#pragma unroll 1
      for (int c = 0; c < n_r; c += 32) {
        const int j = c + lane;
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        bool pass = false;
        if (j < n_r) {
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bd.x), "=f"(bd.y), "=f"(bd.z), "=f"(bd.w) : "r"(sbox + (uint32_t)j * 16u));
          pass = !(q.fx0 > bd.y || q.fx1 < bd.x || q.fy0 > bd.w || q.fy1 < bd.z);
        }
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
          const int bsel = __ffs(m) - 1;
          m &= m - 1;
          float4 bb;
          bb.x = __shfl_sync(0xffffffffu, bd.x, bsel);
          bb.y = __shfl_sync(0xffffffffu, bd.y, bsel);
          bb.z = __shfl_sync(0xffffffffu, bd.z, bsel);
          bb.w = __shfl_sync(0xffffffffu, bd.w, bsel);
          if (q.valid) forward_pair<kFast, K>(A, C, nullptr, bb, 0, n, q.xp, q.yp, st, srec + (uint32_t)(c + bsel) * kRecBytes);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
    }
    stage_it += (unsigned)rounds;
  };
  // finalize (.cu:457-482)
  auto finish = [&](const Footprint& q, const PixelState<K>& st) {
    if (!q.valid) return;
    float* out = A.soft_colors + (long long)n * P * SS + q.pn;
    float* agg = A.aggrs_info + (long long)n * 2 * SS + q.pn;
    if (alpha_f == 0) out[(P - 1) * SS] = st.alpha;
    else if (alpha_f == 1) out[(P - 1) * SS] = st.alpha / A.F;
    else out[(P - 1) * SS] = 1.f - st.alpha;
    if (rgb_f == 0) {
      if (st.face_min != -1 || A.bg_in_args) {   // uncovered pixels keep the caller's pre-fill, or get the background argument
#pragma unroll
        for (int k = 0; k < 3; ++k) out[k * SS] = st.col[k];
      }
      agg[0] = st.depth_min;
      agg[SS] = (float)st.face_min;
    } else {
#pragma unroll
      for (int k = 0; k < 3 * K; ++k) out[k * SS] = div_<kMode>(st.col[k], st.ssum);
      agg[0] = st.ssum;
      agg[SS] = st.smax;
    }
  };

  const Footprint q = place(warp);
  PixelState<K> st;
  init_state(q, st);
  for (int base = 0; tile_live && base < A.F; base += chunk) {
    scan(base);
    __syncthreads();
    if (kBulk) walk_bulk(q, st); else walk(q, st);
    if (base + chunk < A.F) __syncthreads();  // the list is reused by the next round
  }
  finish(q, st);
}

// ---------------------------------------------------------------------------------------------
// backward: one warp per face; faces whose pixel window is huge are deferred to a second launch of
// the same code with one CTA per face (kBig), so a few screen-filling triangles cannot serialise the
// kernel behind a single warp.  Either way a face's gradient is a private fixed-order sum.
// ---------------------------------------------------------------------------------------------
constexpr int kBigFacePixels = 4096;   // window area above which a face goes to the CTA-per-face pass
constexpr int kBigListCap = 16384;     // deferred faces the workspace can list (the rest run inline)

// KF: colour sets [0,KF) are detached -- their image-gradient planes are not read and they get no
// colour gradient (LASR's flow render uses the frame's own positions only under .detach(),
// nnutils/mesh_net.py:102); sets [KF,K) are live.
template <bool k32> struct PlaneIndex { typedef long long type; };
template <> struct PlaneIndex<true> { typedef unsigned type; };

// kCompact: the face comes from the COMPACT record + the op's inputs (compact mode); a separate instantiation, so the
// default kernels are the code they were before the mode existed (folded into one kernel behind a run-time test, the
// single-set backward measured 13 % slower in the DEFAULT mode: profiles/r2/README.md, trip 21).
template <bool kFast, bool kBig, int K, int KF, bool kPacked = false, bool kCompact = false>
__global__ void __launch_bounds__(kBig ? 256 : LASR_BWD_THREADS, kBig ? 2 : (K == 1 ? LASR_BWD_MIN_BLOCKS : (K - KF >= 3 ? LASR_BWD_MIN_BLOCKS_MULTI : LASR_BWD_MIN_BLOCKS_LIVE2))) raster_backward_kernel(RasterArgs A) {
  constexpr int kMode = kFast ? (LASR_FTZ_MUFU_BWD ? 2 : 1) : 0;   // arithmetic mode, see raster_common.cuh
  constexpr int kRecWords = rec_words(K);
  constexpr int P = 3 * K + 1;          // image planes per frame
  constexpr int KL = K - KF;            // live colour sets
  constexpr int NV = 9 + 9 * KL;        // gradient values per face: 9 position + 9 per live colour set
  __shared__ float s_part[8][NV];
  const int lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int S = A.S;
  const long long SS = (long long)S * S;

  float near_ = A.near_, far_ = A.far_;
  if (A.near_far_dev) { near_ = A.near_far_dev[0]; far_ = A.near_far_dev[1]; }
  const float thr = A.dist_eps * A.sigma;
  const float sthr = sqrtf(thr);
  const float inv_range = 1.f / (far_ - near_);
  const float cz_scale = A.inv_gamma / (near_ - far_);
  const int dist_f = kFast ? 2 : A.dist;
  const int rgb_f = kFast ? 1 : A.rgb;
  const int alpha_f = kFast ? 2 : A.alpha;
  const int sample = kFast ? 1 : A.sample;

  // One-warp-per-face pass: a CTA owns a run of nwarps * LASR_BWD_FACES_PER_WARP CONSECUTIVE faces and its warps take
  // them interleaved.  Neighbouring faces of a mesh overlap in their pixel windows (a ~5-pixel face has a ~13-pixel
  // window: every pixel is inside ~7 windows), so keeping them on one SM turns most of the ten plane loads per pixel
  // into L1 hits; with one face per warp consecutive faces landed on 148 different SMs (CTAs are dealt round-robin).
  // Measured (profiles/r2/README.md, trip 3): L1 hit rate 69 -> 80 %, single-set backward -3 % at 4 or 8 faces per
  // warp; the multi-set kernels, whose windows carry 16-22 planes, lose 2 % and keep one face per warp.
  constexpr int kFPW = bwd_faces_per_warp(K);
  const int n_items = kBig ? min(A.big_list[0], kBigListCap) : kFPW;
#pragma unroll 1
  for (int item = kBig ? blockIdx.x : 0; item < n_items; item += kBig ? gridDim.x : 1) {
  const long long wg = kBig ? (long long)A.big_list[1 + item]
                            : ((long long)blockIdx.x * kFPW + item) * nwarps + (threadIdx.x >> 5);
  if (!kBig && wg >= (long long)A.N * A.F) return;
  const int n = (int)(wg / A.F);
  const int f = (int)(wg % A.F);
  const int gtid = kBig ? threadIdx.x : lane;        // index within the group sweeping this face
  const int gsize = kBig ? blockDim.x : 32;

  float tex9[9 * K];
  FaceGeom g;
  if (kCompact) {                // the forward's COMPACT record (80 bytes per face) + the op's own inputs
    float cr[kCompactWords];
    load_compact(A.compact_record + wg * kCompactWords, cr);     // issued first: one memory round trip with the vertices
    float fv[9];
    fetch_face(A, n, f, fv, tex9, sample == 1);
    face_from_compact(fv, sthr, cr, g);
  } else if (A.records_valid) {   // the forward's per-face record (the reference hands faces_info to its backward too)
    load_record<K>(A.face_record + wg * kRecWords, g, tex9);
    const float4 bd = __ldg(A.face_border + wg);
    g.lox = bd.x; g.hix = bd.y; g.loy = bd.z; g.hiy = bd.w;
  } else {
    float fv[9];
    fetch_face(A, n, f, fv, tex9, sample == 1);
    face_setup(fv, sthr, g);
  }
  float* gtex_out = A.grad_textures + wg * (K == 1 ? A.T * 3 : 9 * K) + 9 * KF;

  int xlo, xhi, ylo, yhi;
  coord_window(g.lox, g.hix, S, xlo, xhi);
  coord_window(g.loy, g.hiy, S, ylo, yhi);
  if (!kBig && xlo <= xhi && ylo <= yhi && (xhi - xlo + 1) * (yhi - ylo + 1) > kBigFacePixels) {
    int slot = 0;
    if (lane == 0) slot = atomicAdd(A.big_list, 1);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < kBigListCap) {           // deferred: the CTA-per-face pass writes this face's outputs
      if (lane == 0) A.big_list[1 + slot] = (int)wg;
      continue;
    }
  }
  float gf[9], gt[9 * KL];
#pragma unroll
  for (int k = 0; k < 9; ++k) gf[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 9 * KL; ++k) gt[k] = 0.f;

  if (xlo <= xhi && ylo <= yhi) {
    const int wdt = xhi - xlo + 1;
    const int total = wdt * (yhi - ylo + 1);
    const float inv_wdt = 1.f / (float)wdt;
    const float* sc = A.soft_colors + (long long)n * P * SS;
    const float* ag = A.aggrs_info + (long long)n * 2 * SS;
    const float* gs = A.grad_soft_colors + (long long)n * P * SS;
    const float4* pk = kPacked ? A.packed + (long long)n * (1 + (3 * KL + 3) / 4) * SS : nullptr;
    // Single colour set: per-frame plane offsets as 32-bit values (fill_args bounds P*S*S) on opaque 64-bit
    // bases -> one IMAD.WIDE per load instead of an IADD3/IADD3.X/LEA/LEA.HI.X chain (the ten plane loads of
    // a pixel were ~40 integer instructions per sweep; backward 1.208 -> 1.168 ms on config A).  With several
    // colour sets the 64-bit form schedules better (measured: 2.06 vs 2.12 ms at K = 3), so it is kept there.
    constexpr bool kIdx32 = K == 1 ? LASR_BWD_INDEX32 != 0 : LASR_BWD_INDEX32_MULTI != 0;
    typedef typename PlaneIndex<kIdx32>::type index_t;
    const index_t SSx = (index_t)S * (index_t)S;
    if (kIdx32) asm volatile("" : "+l"(sc), "+l"(ag), "+l"(gs), "+l"(pk));
#if LASR_EXP_PLANES_L1   // TIMING-ONLY bound experiment (wrong gradients): every plane load hits one L1-resident line
#define PLANE(k) ((index_t)(k) * 32u + (index_t)lane)
#else
#define PLANE(k) ((index_t)(k) * SSx + pn)
#endif
    for (int idx = gtid; idx < total; idx += gsize) {
      int xi, yi;
      {
        int r = (int)(((float)idx + 0.5f) * inv_wdt);  // idx / wdt without an integer divide
        int col = idx - r * wdt;
        if (col < 0) { --r; col += wdt; } else if (col >= wdt) { ++r; col -= wdt; }
        xi = xlo + col;
        yi = yhi - r;
      }
      const float xp = A.coords[xi];
      const float yp = A.coords[yi];
      if (xp > g.hix || xp < g.lox || yp > g.hiy || yp < g.loy) continue;
      const index_t pn = (index_t)(S - 1 - yi) * (index_t)S + (index_t)xi;
#include "raster_backward_pair.inc"   // one (pixel, face) pair -> gf / gt
    }
  }
  // fixed-order reduction (16 values by recursive halving + 2 butterflies), one write per value
  float* gface_out = A.grad_faces + wg * 9;
  float v16[16];
#pragma unroll
  for (int k = 0; k < 9; ++k) v16[k] = gf[k];
#pragma unroll
  for (int k = 0; k < 7; ++k) v16[9 + k] = gt[k];
  warp_sum16(v16, lane);
  const float t7 = warp_sum(gt[7]), t8 = warp_sum(gt[8]);
  const int vidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  if (K > 1) {   // several colour sets: the same recursive-halving sum, 16 values at a time
    constexpr int G = (NV + 15) / 16;
    float tot[G];
#pragma unroll
    for (int b = 0; b < G; ++b) {
      float w16[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int q = 16 * b + k;
        w16[k] = q < 9 ? gf[q < 9 ? q : 0] : (q < NV ? gt[q < NV && q >= 9 ? q - 9 : 0] : 0.f);
      }
      warp_sum16(w16, lane);
      tot[b] = w16[0];
    }
    if (!kBig) {
      if ((lane & 1) == 0) {
#pragma unroll
        for (int b = 0; b < G; ++b) {
          const int q = 16 * b + vidx;
          if (q < 9) gface_out[q] = tot[b];
          else if (q < NV) gtex_out[q - 9] = tot[b];
        }
      }
    } else {
      const int warp = threadIdx.x >> 5;
      if ((lane & 1) == 0) {
#pragma unroll
        for (int b = 0; b < G; ++b) {
          const int q = 16 * b + vidx;
          if (q < NV) s_part[warp][q] = tot[b];
        }
      }
      __syncthreads();
      if (threadIdx.x < NV) {
        float t = 0.f;
        for (int w = 0; w < nwarps; ++w) t += s_part[w][threadIdx.x];
        if (threadIdx.x < 9) gface_out[threadIdx.x] = t;
        else gtex_out[threadIdx.x - 9] = t;
      }
      __syncthreads();
    }
    continue;
  }
  if (!kBig) {
    if ((lane & 1) == 0) {
      if (vidx < 9) gface_out[vidx] = v16[0];
      else if (sample == 1) gtex_out[vidx - 9] = v16[0];
    }
    if (sample == 1 && lane == 0) { gtex_out[7] = t7; gtex_out[8] = t8; }
  } else {   // CTA per face: combine the warps' totals in warp order
    const int warp = threadIdx.x >> 5;
    if ((lane & 1) == 0) s_part[warp][vidx] = v16[0];
    if (lane == 0) { s_part[warp][16] = t7; s_part[warp][17] = t8; }
    __syncthreads();
    if (threadIdx.x < 18) {
      float t = 0.f;
      for (int w = 0; w < nwarps; ++w) t += s_part[w][threadIdx.x];
      if (threadIdx.x < 9) gface_out[threadIdx.x] = t;
      else if (sample == 1) gtex_out[threadIdx.x - 9] = t;
    }
    __syncthreads();
  }
  }  // item loop
}

// ---------------------------------------------------------------------------------------------
// Per-pixel pre-pass of the single-set backward (LASR configuration).  Everything a (pixel, face) pair reads from the
// pixel -- image (4 planes), its gradient (4), the softmax sum / max (2): ten scalar gathers per pair -- is independent of
// the face, and every pixel is inside ~25 face windows.  One coalesced pass packs it into two float4 per pixel
//   { g_alpha * (1 - alpha),  sum_k g_k * colour_k,  1 / softmax sum,  softmax max },  { g0, g1, g2, 0 }
// (planes [N][2][S*S], the layout lasr_b200_head_backward_packed writes for three sets), and the pair costs two
// coalesced 128-bit loads, no `1 - alpha`, one MUFU.RCP less and three subtractions less.  4 pixels per thread.
// ---------------------------------------------------------------------------------------------
template <int vec>
__global__ void __launch_bounds__(256) raster_pack_kernel(const float* __restrict__ sc, const float* __restrict__ gs,
                                                          const float* __restrict__ ag, float4* __restrict__ pk,
                                                          long long SS) {
  const int n = blockIdx.y;
  const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * vec;
  if (p >= SS) return;
  const float* s0 = sc + (long long)n * 4 * SS + p;
  const float* g0 = gs + (long long)n * 4 * SS + p;
  const float* a0 = ag + (long long)n * 2 * SS + p;
  float4* o = pk + (long long)n * 2 * SS + p;
  float c[4][vec], g[4][vec], a[2][vec];
  if constexpr (vec == 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(s0 + k * SS)), y = __ldg(reinterpret_cast<const float4*>(g0 + k * SS));
      c[k][0] = x.x; c[k][1] = x.y; c[k][2] = x.z; c[k][3] = x.w;
      g[k][0] = y.x; g[k][1] = y.y; g[k][2] = y.z; g[k][3] = y.w;
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a0 + k * SS));
      a[k][0] = x.x; a[k][1] = x.y; a[k][2] = x.z; a[k][3] = x.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) { c[k][0] = s0[k * SS]; g[k][0] = g0[k * SS]; }
    a[0][0] = a0[0]; a[1][0] = a0[SS];
  }
#pragma unroll
  for (int e = 0; e < vec; ++e) {
    const float T = g[0][e] * c[0][e] + g[1][e] * c[1][e] + g[2][e] * c[2][e];
    o[e] = make_float4(g[3][e] * (1.f - c[3][e]), T, fast_rcp<LASR_FTZ_MUFU_BWD ? 2 : 1>(a[0][e]), a[1][e]);
    o[SS + e] = make_float4(g[0][e], g[1][e], g[2][e], 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int fail(const std::string& msg) { return lasr_fail(msg); }
static int check_launch(const char* what) { return lasr_check_launch(what); }

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Heaviest-first dispatch costs a memset, ~3 atomics per face in the setup kernel and a one-CTA sort
// (~25 us together); it pays once the forward runs for more than a few hundred microseconds (Spot, 64 frames:
// 1.21 -> 1.05 ms; ico3, 64 frames: 0.36 -> 0.39 ms).  LASR_B200_FWD_LPT=0/1 in the environment forces it.
static bool use_lpt(int N, int F) {
  static const int forced = [] { const char* e = getenv("LASR_B200_FWD_LPT"); return e ? atoi(e) : -1; }();
  if (!LASR_FWD_LPT) return false;
  if (forced >= 0) return forced != 0;
  return (long long)N * F >= 150000;
}

static size_t num_tiles(int S) { return (size_t)((S + kTileW - 1) / kTileW) * (size_t)((S + kTileH - 1) / kTileH); }

// The packing pre-pass of the single-set backward (raster_pack_kernel) costs one pass over the images (72 B per pixel) and
// saves eight scalar gathers per (pixel, face) pair: it pays when a pixel sits inside many face windows.  Estimate of that
// number from what the host knows -- F faces of ~0.8 S^2 / F pixels each (an object over ~40 % of the image, two layers),
// dilated by the blur radius r on every side: F (w + 2r)^2 / S^2.  Measured (profiles/r2/README.md, trip 12): config D
// (estimate 28) backward -6.6 %, config A (10) +-0, ico3 (4) +5 % -- the pass is taken from 16 up.  The workspace reserves
// room for it by the same estimate at LASR's sigma = 1e-4 (its size may depend on N, F, S only).
static float est_windows_per_pixel(int F, int S, float sthr_ndc) {
  const float r = sthr_ndc * 0.5f * (float)S;
  const float w = sqrtf(0.8f * (float)S * (float)S / (float)F);
  return (float)F * (w + 2.f * r) * (w + 2.f * r) / ((float)S * (float)S);
}
static bool pack_room(int F, int S, int K) {
  static const int forced = [] { const char* e = getenv("LASR_B200_BWD_PACK"); return e ? atoi(e) : -1; }();
  if (!LASR_BWD_PACK_K1 || K != 1 || forced == 0) return false;
  return forced == 1 || est_windows_per_pixel(F, S, 0.030349f) >= 16.f;     // sqrt(ln(9999) * 1e-4)
}

static size_t raster_ws_bytes(int N, int F, int S, int K) {
  if (N <= 0 || F <= 0 || S <= 0 || K < 1 || K > kMaxSets) return 0;
  const size_t nf = (size_t)N * (size_t)F;
  return align256(nf * sizeof(uint2)) + align256((size_t)S * sizeof(float)) + align256((size_t)N * 4 * sizeof(int)) +
         align256((size_t)(kBigListCap + 1) * sizeof(int)) +
         (LASR_FWD_LPT ? 2 * align256((size_t)N * num_tiles(S) * sizeof(int)) : 0) +
         align256(nf * sizeof(float4)) +
         align256(nf * rec_words(K) * sizeof(float)) +
         (pack_room(F, S, K) ? align256((size_t)N * 2 * (size_t)S * S * sizeof(float4)) : 0);   // packed per-pixel gradients
}

static int fill_args(RasterArgs& A, int N, int F, int S, int T, float near_, float far_, const float* nf_dev,
                     float eps, float sigma, int dist, float dist_eps, float gamma, int rgb, int alpha, int sample,
                     int double_side, void* workspace, size_t workspace_bytes, int K = 1, bool carve = true) {
  if (N <= 0 || F <= 0 || S <= 0 || T <= 0) return fail("raster: batch_size, num_faces, image_size, texture_size must be positive");
  if (S > 8192) return fail("raster: image_size > 8192 unsupported (per-frame plane offsets are 32-bit)");
  if (dist < 0 || dist > 2 || rgb < 0 || rgb > 1 || alpha < 0 || alpha > 2 || sample < 0 || sample > 1)
    return fail("raster: bad func id");
  if (sample == 1 && T != 3) return fail("raster: 'vertex' textures must be [N,F,3,3]");
  if (K < 1 || K > kMaxSets) return fail("render: num_sets must be 1, 2 or 3");
  if (carve && (!workspace || workspace_bytes < raster_ws_bytes(N, F, S, K))) return fail("raster: workspace too small");
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)   // bulk copies and 128-bit loads of the records need 16-byte alignment
    return fail("raster: workspace must be 256-byte aligned (cudaMalloc / torch allocations are)");
  if (K > 1 && !(dist == 2 && rgb == 1 && alpha == 2 && sample == 1 && double_side != 0))
    return fail("render: several colour sets need euclidean / softmax / prod / vertex / fill_back (LASR's configuration)");
  A.N = N; A.F = F; A.S = S; A.T = T; A.K = K;
  A.R = (int)sqrt((double)T);  // .cu:696
  A.near_ = near_; A.far_ = far_; A.near_far_dev = nf_dev;
  A.eps = eps; A.sigma = sigma; A.dist_eps = dist_eps; A.gamma = gamma;
  A.dist = dist; A.rgb = rgb; A.alpha = alpha; A.sample = sample; A.double_side = double_side;
  A.inv_sigma = (float)(1.0 / (double)sigma);
  A.inv_gamma = (float)(1.0 / (double)gamma);
  if (!carve) return 0;        // the compact backward lays out its own (much smaller) scratch
  // workspace carve-up (all 256-byte aligned)
  char* p = reinterpret_cast<char*>(workspace);
  const size_t nf = (size_t)N * F;
  A.face_window = reinterpret_cast<uint2*>(p);   p += align256(nf * sizeof(uint2));
  A.coords = reinterpret_cast<float*>(p);        p += align256((size_t)S * sizeof(float));
  A.frame_window = reinterpret_cast<int*>(p);    p += align256((size_t)N * 4 * sizeof(int));
  A.big_list = reinterpret_cast<int*>(p);        p += align256((size_t)(kBigListCap + 1) * sizeof(int));
  if (LASR_FWD_LPT) {   // space is always reserved (the workspace size must not depend on the environment)
    int* tc = reinterpret_cast<int*>(p);         p += align256((size_t)N * num_tiles(S) * sizeof(int));
    int* to = reinterpret_cast<int*>(p);         p += align256((size_t)N * num_tiles(S) * sizeof(int));
    if (use_lpt(N, F)) { A.tile_count = tc; A.tile_order = to; }
  }
  A.face_border = reinterpret_cast<float4*>(p);  p += align256(nf * sizeof(float4));
  A.face_record = reinterpret_cast<float*>(p);   p += align256(nf * rec_words(K) * sizeof(float));
  A.pack_area = pack_room(F, S, K) ? reinterpret_cast<float4*>(p) : nullptr;
  return 0;
}

static bool is_fast(const RasterArgs& A);

// ---- compact mode: what is kept from forward to backward is a separate caller-owned buffer of 80 bytes per face (+ the
// S-entry coordinate table) instead of the whole workspace (200 - 264 bytes per face); the workspace becomes scratch the
// caller may free as soon as the forward call returns, and the backward takes a scratch of its own.
static size_t raster_saved_bytes(int N, int F, int S) {
  if (N <= 0 || F <= 0 || S <= 0) return 0;
  return align256((size_t)S * sizeof(float)) + align256((size_t)N * F * kCompactWords * sizeof(float));
}
static size_t backward_scratch_bytes(int N, int F, int S, int K) {
  if (N <= 0 || F <= 0 || S <= 0) return 0;
  return align256((size_t)(kBigListCap + 1) * sizeof(int)) +
         (pack_room(F, S, K) ? align256((size_t)N * 2 * (size_t)S * S * sizeof(float4)) : 0);
}
static int attach_saved(RasterArgs& A, void* saved, size_t saved_bytes) {
  if (!saved || saved_bytes < raster_saved_bytes(A.N, A.F, A.S)) return fail("raster: saved buffer too small");
  if (reinterpret_cast<uintptr_t>(saved) % 256 != 0) return fail("raster: saved buffer must be 256-byte aligned");
  char* p = reinterpret_cast<char*>(saved);
  A.coords = reinterpret_cast<float*>(p);  p += align256((size_t)A.S * sizeof(float));
  A.compact_record = reinterpret_cast<float*>(p);
  return 0;
}
// backward scratch in compact mode: big-face list, packing area of the single-set pre-pass, then `extra` bytes for the
// caller (per-face gradients of the indexed path); returns the start of the extra region or null
static char* carve_backward_scratch(RasterArgs& A, void* workspace, size_t workspace_bytes, size_t extra) {
  if (!workspace || workspace_bytes < backward_scratch_bytes(A.N, A.F, A.S, A.K) + extra) {
    fail("raster_backward: scratch too small");
    return nullptr;
  }
  char* p = reinterpret_cast<char*>(workspace);
  A.big_list = reinterpret_cast<int*>(p);  p += align256((size_t)(kBigListCap + 1) * sizeof(int));
  A.pack_area = nullptr;
  if (pack_room(A.F, A.S, A.K)) {
    A.pack_area = reinterpret_cast<float4*>(p);
    p += align256((size_t)A.N * 2 * (size_t)A.S * A.S * sizeof(float4));
  }
  return p;
}

// forward launch: dispatch order first (if this launch uses it), then one CTA per (frame, tile)
static int launch_forward(const RasterArgs& A, cudaStream_t st) {
  const int S = A.S, F = A.F, N = A.N;
  const int tiles = (int)num_tiles(S);
  if (A.tile_order) {
    raster_order_kernel<<<1, 1024, 0, st>>>(A.tile_count, A.tile_order, tiles * N);
    if (check_launch("raster_order_kernel")) return 1;
  }
  int chunk = ((F + kFwdThreads - 1) / kFwdThreads) * kFwdThreads;   // a multiple of 32 faces per warp
  const bool bulk = LASR_FWD_BULK != 0 && (A.K > 1 || is_fast(A));
  // with the staging ring the list is capped so that four CTAs (three with several colour sets) still fit an SM
  const int max_chunk = bulk ? (A.K == 1 ? 7936 : 9984) : kMaxScanChunk;
  if (chunk > max_chunk) chunk = (max_chunk / kFwdThreads) * kFwdThreads;
  const size_t smem = (size_t)chunk * sizeof(int) + (bulk ? (size_t)kStageBufs * stage_bytes(A.K) : 0);
  if (bulk) {   // more than 48 KB of dynamic shared memory is an opt-in per kernel AND per device
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && !done[dev]) {
      if (cudaFuncSetAttribute(raster_forward_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(raster_forward_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(raster_forward_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
        return fail("raster_forward: cannot opt in to the staging ring's shared memory");
      done[dev] = true;
    }
  }
  const unsigned grid = (unsigned)((long long)tiles * N);
  if (A.K == 3) raster_forward_kernel<true, 3><<<grid, kFwdThreads, smem, st>>>(A, chunk);
  else if (A.K == 2) raster_forward_kernel<true, 2><<<grid, kFwdThreads, smem, st>>>(A, chunk);
  else if (is_fast(A)) raster_forward_kernel<true, 1><<<grid, kFwdThreads, smem, st>>>(A, chunk);
  else raster_forward_kernel<false, 1><<<grid, kFwdThreads, smem, st>>>(A, chunk);
  return check_launch("raster_forward_kernel");
}

// compact mode (records_valid == 2): the single-set op of the reference signature and the fused region's three packed sets
static int launch_backward_compact(const RasterArgs& A, unsigned blocks, bool pack_first, cudaStream_t st) {
  if (A.K == 1 && pack_first) {
    RasterArgs P = A;
    P.packed = A.pack_area;
    raster_backward_kernel<true, false, 1, 0, true, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(P);
    if (check_launch("raster_backward_kernel<packed, compact>")) return 1;
    raster_backward_kernel<true, true, 1, 0, true, true><<<592, 256, 0, st>>>(P);
    return check_launch("raster_backward_kernel<packed, big, compact>");
  }
  if (A.K == 1 && !A.packed) {
    if (is_fast(A)) {
      raster_backward_kernel<true, false, 1, 0, false, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
      if (check_launch("raster_backward_kernel<compact>")) return 1;
      raster_backward_kernel<true, true, 1, 0, false, true><<<592, 256, 0, st>>>(A);
    } else {
      raster_backward_kernel<false, false, 1, 0, false, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
      if (check_launch("raster_backward_kernel<compact>")) return 1;
      raster_backward_kernel<false, true, 1, 0, false, true><<<592, 256, 0, st>>>(A);
    }
    return check_launch("raster_backward_kernel<big, compact>");
  }
  if (A.K == 3 && A.packed && A.first_live_set == 1) {
    raster_backward_kernel<true, false, 3, 1, true, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
    if (check_launch("raster_backward_kernel<packed, compact>")) return 1;
    raster_backward_kernel<true, true, 3, 1, true, true><<<592, 256, 0, st>>>(A);
    return check_launch("raster_backward_kernel<packed, big, compact>");
  }
  return fail("raster_backward: compact mode serves one colour set, or three packed sets with the first detached");
}

// main pass (one warp per face, huge faces deferred) + CTA-per-face pass over the deferred list
static int launch_backward(const RasterArgs& A, long long nf, cudaStream_t st) {
  if (cudaMemsetAsync(A.big_list, 0, sizeof(int), st) != cudaSuccess) return fail("raster_backward: memset failed");
  const int faces_per_block = (LASR_BWD_THREADS / 32) * bwd_faces_per_warp(A.K);
  const unsigned blocks = (unsigned)((nf + faces_per_block - 1) / faces_per_block);
  const int kf = A.first_live_set;
  static const int pack_forced = [] { const char* e = getenv("LASR_B200_BWD_PACK"); return e ? atoi(e) : -1; }();
  if (!A.packed && A.pack_area && A.K == 1 && is_fast(A) &&
      (pack_forced == 1 || est_windows_per_pixel(A.F, A.S, sqrtf(A.dist_eps * A.sigma)) >= 16.f)) {   // pack per pixel first
    const long long SS = (long long)A.S * A.S;
    const int vec = SS % 4 == 0 ? 4 : 1;
    const unsigned gx = (unsigned)((SS / vec + 255) / 256);
    if (vec == 4) raster_pack_kernel<4><<<dim3(gx, A.N), 256, 0, st>>>(A.soft_colors, A.grad_soft_colors, A.aggrs_info, A.pack_area, SS);
    else raster_pack_kernel<1><<<dim3(gx, A.N), 256, 0, st>>>(A.soft_colors, A.grad_soft_colors, A.aggrs_info, A.pack_area, SS);
    if (check_launch("raster_pack_kernel")) return 1;
    if (A.records_valid == 2) return launch_backward_compact(A, blocks, true, st);
    RasterArgs P = A;
    P.packed = A.pack_area;
    raster_backward_kernel<true, false, 1, 0, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(P);
    if (check_launch("raster_backward_kernel<packed>")) return 1;
    raster_backward_kernel<true, true, 1, 0, true><<<592, 256, 0, st>>>(P);
    return check_launch("raster_backward_kernel<packed, big>");
  }
  if (A.records_valid == 2) return launch_backward_compact(A, blocks, false, st);
  if (A.packed) {   // the fused region's packed per-pixel gradients (lasr_b200_render_backward_packed: K = 3, set 0 detached)
    raster_backward_kernel<true, false, 3, 1, true><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
    if (check_launch("raster_backward_kernel<packed>")) return 1;
    raster_backward_kernel<true, true, 3, 1, true><<<592, 256, 0, st>>>(A);
    return check_launch("raster_backward_kernel<packed, big>");
  }
  if (A.K == 3 && kf == 1) raster_backward_kernel<true, false, 3, 1><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  else if (A.K == 3) raster_backward_kernel<true, false, 3, 0><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  else if (A.K == 2 && kf == 1) raster_backward_kernel<true, false, 2, 1><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  else if (A.K == 2) raster_backward_kernel<true, false, 2, 0><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  else if (is_fast(A)) raster_backward_kernel<true, false, 1, 0><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  else raster_backward_kernel<false, false, 1, 0><<<blocks, LASR_BWD_THREADS, 0, st>>>(A);
  if (check_launch("raster_backward_kernel")) return 1;
  const unsigned big_blocks = 592;   // 4 per SM; each loops over the (usually empty) deferred list
  if (A.K == 3 && kf == 1) raster_backward_kernel<true, true, 3, 1><<<big_blocks, 256, 0, st>>>(A);
  else if (A.K == 3) raster_backward_kernel<true, true, 3, 0><<<big_blocks, 256, 0, st>>>(A);
  else if (A.K == 2 && kf == 1) raster_backward_kernel<true, true, 2, 1><<<big_blocks, 256, 0, st>>>(A);
  else if (A.K == 2) raster_backward_kernel<true, true, 2, 0><<<big_blocks, 256, 0, st>>>(A);
  else if (is_fast(A)) raster_backward_kernel<true, true, 1, 0><<<big_blocks, 256, 0, st>>>(A);
  else raster_backward_kernel<false, true, 1, 0><<<big_blocks, 256, 0, st>>>(A);
  return check_launch("raster_backward_kernel<big>");
}

static bool is_fast(const RasterArgs& A) {
  // LASR's training configuration (nnutils/mesh_net.py:136-145): euclidean / softmax / prod /
  // vertex colours / fill_back
  return A.dist == 2 && A.rgb == 1 && A.alpha == 2 && A.sample == 1 && A.double_side != 0;
}

// ---------------------------------------------------------------------------------------------
// indexed mode: per-face gradients -> per-vertex gradients through the vertex->corner adjacency
// (CSR, ascending face order => fixed summation order, no atomics)
// ---------------------------------------------------------------------------------------------
struct SetGrads { float* p[kMaxSets]; int shift[kMaxSets]; };   // per colour set: [N,V,3] output (null: not wanted), frame shift

__global__ void __launch_bounds__(256) vertex_gather_kernel(const float* __restrict__ grad_faces,
                                                            const float* __restrict__ grad_textures,
                                                            const int* __restrict__ csr_offsets,
                                                            const int* __restrict__ csr_entries, float* grad_verts,
                                                            SetGrads out, int N, int V, int F, int K,
                                                            int first_live) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * V) return;
  const int n = (int)(i / V), v = (int)(i % V);
  const float* gf = grad_faces + (long long)n * F * 9;
  const float* gt = grad_textures + (long long)n * F * 9 * K;   // per face: K sets x 3 corners x rgb
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  float c[3 * kMaxSets];
#pragma unroll
  for (int q = 0; q < 3 * kMaxSets; ++q) c[q] = 0.f;
  for (int e = csr_offsets[v]; e < csr_offsets[v + 1]; ++e) {
    const int ent = csr_entries[e];     // entry = face*3 + corner
    const int k = ent * 3;              // offset of that corner's 3 position gradients
    a0 += gf[k]; a1 += gf[k + 1]; a2 += gf[k + 2];
    const float* gc = gt + (long long)(ent / 3) * 9 * K + (ent % 3) * 3;
#pragma unroll
    for (int s = 0; s < kMaxSets; ++s)
      if (s >= first_live && s < K) { c[3 * s] += gc[9 * s]; c[3 * s + 1] += gc[9 * s + 1]; c[3 * s + 2] += gc[9 * s + 2]; }
  }
  grad_verts[i * 3] = a0; grad_verts[i * 3 + 1] = a1; grad_verts[i * 3 + 2] = a2;
#pragma unroll
  for (int s = 0; s < kMaxSets; ++s)
    if (s < K && out.p[s]) {            // the gradient lands in the frame the colours were read from
      int m = n + out.shift[s];
      if (m >= N) m -= N;
      float* o = out.p[s] + ((long long)m * V + v) * 3;
      o[0] = c[3 * s]; o[1] = c[3 * s + 1]; o[2] = c[3 * s + 2];
    }
}

}  // namespace lasr

using namespace lasr;

extern "C" size_t lasr_b200_raster_workspace_bytes(int N, int F, int S) { return raster_ws_bytes(N, F, S, 1); }

static int raster_forward_impl(const float* faces, const float* textures, float* aggrs_info, float* soft_colors,
                               void* workspace, size_t workspace_bytes, int N, int F, int S, int T, float near_, float far_,
                               const float* near_far_dev, float eps, float sigma_val, int func_id_dist, float dist_eps,
                               float gamma_val, int func_id_rgb, int func_id_alpha, int texture_sample_type,
                               int double_side, const float* background_host, void* stream, void* saved = nullptr,
                               size_t saved_bytes = 0, bool compact = false) {
  RasterArgs A{};
  if (fill_args(A, N, F, S, T, near_, far_, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                func_id_rgb, func_id_alpha, texture_sample_type, double_side, workspace, workspace_bytes))
    return 1;
  if (compact && attach_saved(A, saved, saved_bytes)) return 1;
  if (!faces || !textures || !aggrs_info || !soft_colors) return fail("raster_forward: null tensor pointer");
  A.faces = faces; A.textures = textures; A.aggrs_info = aggrs_info; A.soft_colors = soft_colors;
  if (background_host) {   // no pre-fill pass over the image and no read of it in the forward
    A.bg_in_args = 1;
    for (int k = 0; k < 3; ++k) A.bg[k] = background_host[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long nf = (long long)N * F;
  const long long nthreads = nf > S ? nf : S;
  if (cudaMemsetAsync(A.frame_window, 0x7f, (size_t)N * 4 * sizeof(int), st) != cudaSuccess ||
      (A.tile_count && cudaMemsetAsync(A.tile_count, 0, (size_t)N * num_tiles(S) * sizeof(int), st) != cudaSuccess))
    return fail("render_forward: memset failed");
  raster_setup_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(A, 1);
  if (check_launch("raster_setup_kernel")) return 1;
  return launch_forward(A, st);
}

extern "C" int lasr_b200_raster_forward(const float* faces, const float* textures, float* aggrs_info,
                                        float* soft_colors, void* workspace, size_t workspace_bytes, int N, int F,
                                        int S, int T, float near_, float far_, const float* near_far_dev, float eps,
                                        float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                        int func_id_rgb, int func_id_alpha, int texture_sample_type, int double_side,
                                        void* stream) {
  return raster_forward_impl(faces, textures, aggrs_info, soft_colors, workspace, workspace_bytes, N, F, S, T, near_, far_,
                             near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                             texture_sample_type, double_side, nullptr, stream);
}

extern "C" int lasr_b200_raster_forward_bg(const float* faces, const float* textures, float* aggrs_info,
                                           float* soft_colors, void* workspace, size_t workspace_bytes, int N, int F,
                                           int S, int T, float near_, float far_, const float* near_far_dev, float eps,
                                           float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                           int func_id_rgb, int func_id_alpha, int texture_sample_type, int double_side,
                                           const float* background_host, void* stream) {
  if (!background_host) return fail("raster_forward_bg: null background");
  return raster_forward_impl(faces, textures, aggrs_info, soft_colors, workspace, workspace_bytes, N, F, S, T, near_, far_,
                             near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                             texture_sample_type, double_side, background_host, stream);
}

static int raster_backward_impl(const float* faces, const float* textures, const float* soft_colors,
                                const float* aggrs_info, float* grad_faces, float* grad_textures,
                                const float* grad_soft_colors, void* workspace, size_t workspace_bytes,
                                int N, int F, int S, int T, float near_, float far_,
                                const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                                float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                int texture_sample_type, int double_side, int setup_done, void* stream,
                                void* saved = nullptr, size_t saved_bytes = 0, bool compact = false) {
  RasterArgs A{};
  if (fill_args(A, N, F, S, T, near_, far_, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                func_id_rgb, func_id_alpha, texture_sample_type, double_side, workspace, workspace_bytes, 1, !compact))
    return 1;
  A.records_valid = setup_done;
  if (compact) {      // the forward's compact records + coordinate table; the workspace is this call's own scratch
    if (attach_saved(A, saved, saved_bytes) || !carve_backward_scratch(A, workspace, workspace_bytes, 0)) return 1;
    A.records_valid = 2;
    setup_done = 1;
  }
  if (!faces || !textures || !aggrs_info || !soft_colors || !grad_faces || !grad_textures || !grad_soft_colors)
    return fail("raster_backward: null tensor pointer");
  A.faces = faces; A.textures = textures;
  A.aggrs_info = const_cast<float*>(aggrs_info);
  A.soft_colors = const_cast<float*>(soft_colors);
  A.grad_soft_colors = grad_soft_colors; A.grad_faces = grad_faces; A.grad_textures = grad_textures;
  cudaStream_t st = (cudaStream_t)stream;
  const long long nf = (long long)N * F;
  if (texture_sample_type == 0) {  // 'surface' texel gradients are scattered with atomics
    cudaError_t e = cudaMemsetAsync(grad_textures, 0, (size_t)nf * T * 3 * sizeof(float), st);
    if (e != cudaSuccess) return fail(std::string("raster_backward memset: ") + cudaGetErrorString(e));
  }
  if (!setup_done) {  // only the coordinate table is needed from the setup pass
    RasterArgs B = A;
    B.N = 1; B.F = 1;  // one dummy face: fills coords[0..S)
    raster_setup_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(B, 0);
    if (check_launch("raster_setup_kernel")) return 1;
  }
  return launch_backward(A, nf, st);
}

extern "C" int lasr_b200_raster_backward(const float* faces, const float* textures, const float* soft_colors,
                                         const float* aggrs_info, float* grad_faces, float* grad_textures,
                                         const float* grad_soft_colors, void* workspace, size_t workspace_bytes,
                                         int N, int F, int S, int T, float near_, float far_,
                                         const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                                         float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                         int texture_sample_type, int double_side, int setup_done, void* stream) {
  return raster_backward_impl(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_soft_colors,
                              workspace, workspace_bytes, N, F, S, T, near_, far_, near_far_dev, eps, sigma_val,
                              func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, texture_sample_type,
                              double_side, setup_done, stream);
}

extern "C" size_t lasr_b200_raster_saved_bytes(int N, int F, int S) { return raster_saved_bytes(N, F, S); }
extern "C" size_t lasr_b200_raster_backward_scratch_bytes(int N, int F, int S) { return backward_scratch_bytes(N, F, S, 1); }

extern "C" int lasr_b200_raster_forward_compact(const float* faces, const float* textures, float* aggrs_info,
                                                float* soft_colors, void* saved, size_t saved_bytes, void* workspace,
                                                size_t workspace_bytes, int N, int F, int S, int T, float near_,
                                                float far_, const float* near_far_dev, float eps, float sigma_val,
                                                int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                                int func_id_alpha, int texture_sample_type, int double_side,
                                                const float* background_host, void* stream) {
  return raster_forward_impl(faces, textures, aggrs_info, soft_colors, workspace, workspace_bytes, N, F, S, T, near_, far_,
                             near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                             texture_sample_type, double_side, background_host, stream, saved, saved_bytes, true);
}

extern "C" int lasr_b200_raster_backward_compact(const float* faces, const float* textures, const float* soft_colors,
                                                 const float* aggrs_info, float* grad_faces, float* grad_textures,
                                                 const float* grad_soft_colors, const void* saved, size_t saved_bytes,
                                                 void* workspace, size_t workspace_bytes, int N, int F, int S, int T,
                                                 float near_, float far_, const float* near_far_dev, float eps,
                                                 float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                                 int func_id_rgb, int func_id_alpha, int texture_sample_type,
                                                 int double_side, void* stream) {
  return raster_backward_impl(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_soft_colors,
                              workspace, workspace_bytes, N, F, S, T, near_, far_, near_far_dev, eps, sigma_val,
                              func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, texture_sample_type,
                              double_side, 1, stream, const_cast<void*>(saved), saved_bytes, true);
}

// ---------------------------------------------------------------------------------------------
// indexed ("fused") entry points: vertices + shared topology + vertex colours in, image out;
// gradients come back per vertex.  No [N,F,9] gathered tensors in HBM on either pass.
// ---------------------------------------------------------------------------------------------
extern "C" size_t lasr_b200_render_workspace_bytes(int N, int F, int S, int num_sets) {
  if (N <= 0 || F <= 0 || S <= 0 || num_sets < 1 || num_sets > kMaxSets) return 0;
  return raster_ws_bytes(N, F, S, num_sets) + align256((size_t)N * F * 9 * sizeof(float)) +
         align256((size_t)N * F * 9 * num_sets * sizeof(float));
}

// colour sets of an indexed render: one stacked tensor [K][N,V,3], or K separate [N,V,3] tensors each read with a frame
// shift (frame n takes frame (n + shift) mod N: LASR pairs frame i with i +- bs, mesh_net.py:318-335)
struct ColourSets { const float* p[kMaxSets]; int shift[kMaxSets]; };

static ColourSets stacked_sets(const float* vcolors, int N, int V, int K) {
  ColourSets c{};
  for (int s = 0; s < K && s < kMaxSets; ++s) c.p[s] = vcolors ? vcolors + (size_t)s * N * V * 3 : nullptr;
  return c;
}

static int fill_indexed(RasterArgs& A, const float* verts, const int* tri, const ColourSets& sets, int V) {
  if (!verts || !tri || V <= 0) return fail("render: null vertex / topology pointer");
  if (A.sample != 1) return fail("render: the indexed path takes 'vertex' textures");
  for (int s = 0; s < A.K; ++s) {
    if (!sets.p[s]) return fail("render: null colour pointer");
    if (sets.shift[s] < 0 || sets.shift[s] >= A.N) return fail("render: a colour set's frame shift must lie in [0, N)");
    A.set_ptr[s] = sets.p[s]; A.set_shift[s] = sets.shift[s];
  }
  A.verts = verts; A.tri = tri; A.V = V;
  return 0;
}

static int render_forward_impl(const float* verts, const int32_t* tri, const ColourSets& sets,
                               float* aggrs_info, float* soft_colors, void* workspace,
                               size_t workspace_bytes, int N, int V, int F, int S, float near_, float far_,
                               const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                               float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                               int double_side, int num_sets, const float* background, void* stream,
                               void* saved = nullptr, size_t saved_bytes = 0, bool compact = false) {
  RasterArgs A{};
  if (num_sets < 1 || num_sets > kMaxSets) return fail("render: num_sets must be 1, 2 or 3");
  if (workspace_bytes < (compact ? raster_ws_bytes(N, F, S, num_sets) : lasr_b200_render_workspace_bytes(N, F, S, num_sets)))
    return fail("render: workspace too small");   // (compact: no room for the backward's per-face gradients needed)
  if (fill_args(A, N, F, S, 3, near_, far_, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                func_id_rgb, func_id_alpha, 1, double_side, workspace, workspace_bytes, num_sets))
    return 1;
  if (fill_indexed(A, verts, tri, sets, V)) return 1;
  if (compact && attach_saved(A, saved, saved_bytes)) return 1;
  if (!aggrs_info || !soft_colors) return fail("render_forward: null tensor pointer");
  A.aggrs_info = aggrs_info; A.soft_colors = soft_colors;
  if (background) {   // host array of 3*num_sets floats: no pre-fill pass over the image, no read of it here
    A.bg_in_args = 1;
    for (int k = 0; k < 3 * num_sets; ++k) A.bg[k] = background[k];
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long nf = (long long)N * F;
  const long long nthreads = nf > S ? nf : S;
  if (cudaMemsetAsync(A.frame_window, 0x7f, (size_t)N * 4 * sizeof(int), st) != cudaSuccess ||
      (A.tile_count && cudaMemsetAsync(A.tile_count, 0, (size_t)N * num_tiles(S) * sizeof(int), st) != cudaSuccess))
    return fail("render_forward: memset failed");
  raster_setup_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(A, 1);
  if (check_launch("raster_setup_kernel")) return 1;
  return launch_forward(A, st);
}

extern "C" int lasr_b200_render_forward(const float* verts, const int32_t* tri, const float* vcolors,
                                        float* aggrs_info, float* soft_colors, void* workspace,
                                        size_t workspace_bytes, int N, int V, int F, int S, float near_, float far_,
                                        const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                                        float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                        int double_side, int num_sets, const float* background, void* stream) {
  if (!vcolors || num_sets < 1 || num_sets > kMaxSets || N <= 0 || V <= 0)
    return fail("render_forward: null colour pointer or bad num_sets");
  return render_forward_impl(verts, tri, stacked_sets(vcolors, N, V, num_sets), aggrs_info, soft_colors, workspace,
                             workspace_bytes, N, V, F, S, near_, far_, near_far_dev, eps, sigma_val, func_id_dist,
                             dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side, num_sets, background, stream);
}

static ColourSets separate_sets(const float* const* set_ptrs, const int* set_shifts, int K) {
  ColourSets c{};
  for (int s = 0; s < K && s < kMaxSets; ++s) { c.p[s] = set_ptrs ? set_ptrs[s] : nullptr; c.shift[s] = set_shifts ? set_shifts[s] : 0; }
  return c;
}

extern "C" int lasr_b200_render_forward_sets(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                             const int* set_shifts, float* aggrs_info, float* soft_colors,
                                             void* workspace, size_t workspace_bytes, int N, int V, int F, int S,
                                             float near_, float far_, const float* near_far_dev, float eps,
                                             float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                             int func_id_rgb, int func_id_alpha, int double_side, int num_sets,
                                             const float* background, void* stream) {
  if (!set_ptrs || num_sets < 1 || num_sets > kMaxSets) return fail("render_forward_sets: null set array or bad num_sets");
  return render_forward_impl(verts, tri, separate_sets(set_ptrs, set_shifts, num_sets), aggrs_info, soft_colors,
                             workspace, workspace_bytes, N, V, F, S, near_, far_, near_far_dev, eps, sigma_val,
                             func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side, num_sets,
                             background, stream);
}

static SetGrads stacked_grads(float* grad_vcolors, int N, int V, int K) {
  SetGrads g{};
  for (int s = 0; s < K && s < kMaxSets; ++s) g.p[s] = grad_vcolors ? grad_vcolors + (size_t)s * N * V * 3 : nullptr;
  return g;
}

static int render_backward_impl(const float* verts, const int32_t* tri, const ColourSets& sets,
                                const float* soft_colors, const float* aggrs_info, const float* grad_soft_colors,
                                const float* packed_grad, const int32_t* csr_offsets, const int32_t* csr_entries,
                                float* grad_verts, const SetGrads& grad_sets, void* workspace, size_t workspace_bytes, int N,
                                int V, int F, int S, float near_, float far_, const float* near_far_dev, float eps,
                                float sigma_val, int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                int func_id_alpha, int double_side, int num_sets, int detached_sets, int setup_done,
                                void* stream, void* saved = nullptr, size_t saved_bytes = 0, bool compact = false) {
  RasterArgs A{};
  if (num_sets < 1 || num_sets > kMaxSets) return fail("render: num_sets must be 1, 2 or 3");
  if (detached_sets < 0 || detached_sets > 1 || detached_sets >= num_sets)
    return fail("render_backward: detached_sets must be 0, or 1 with num_sets >= 2");
  if (!compact && workspace_bytes < lasr_b200_render_workspace_bytes(N, F, S, num_sets)) return fail("render: workspace too small");
  if (fill_args(A, N, F, S, 3, near_, far_, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                func_id_rgb, func_id_alpha, 1, double_side, workspace, workspace_bytes, num_sets, !compact))
    return 1;
  A.records_valid = setup_done;
  const size_t grad_bytes = align256((size_t)N * F * 9 * sizeof(float)) + align256((size_t)N * F * 9 * num_sets * sizeof(float));
  char* scratch = nullptr;
  if (compact) {      // the forward's compact records + coordinate table; the workspace is this call's own scratch
    if (attach_saved(A, saved, saved_bytes)) return 1;
    scratch = carve_backward_scratch(A, workspace, workspace_bytes, grad_bytes);
    if (!scratch) return 1;
    A.records_valid = 2;
    setup_done = 1;
  } else {
    scratch = reinterpret_cast<char*>(workspace) + raster_ws_bytes(N, F, S, num_sets);
  }
  A.first_live_set = detached_sets;
  if (fill_indexed(A, verts, tri, sets, V)) return 1;
  if (!csr_offsets || !csr_entries || !grad_verts) return fail("render_backward: null tensor pointer");
  for (int s = detached_sets; s < num_sets; ++s)
    if (!grad_sets.p[s]) return fail("render_backward: null colour-gradient pointer of a live set");
  if (packed_grad) {
    if (!(num_sets == 3 && detached_sets == 1)) return fail("render_backward_packed: needs three colour sets with the first detached");
    A.packed = reinterpret_cast<const float4*>(packed_grad);
  } else if (!soft_colors || !aggrs_info || !grad_soft_colors) {
    return fail("render_backward: null tensor pointer");
  }
  A.aggrs_info = const_cast<float*>(aggrs_info);
  A.soft_colors = const_cast<float*>(soft_colors);
  A.grad_soft_colors = grad_soft_colors;
  const long long nf = (long long)N * F;
  A.grad_faces = reinterpret_cast<float*>(scratch);
  A.grad_textures = reinterpret_cast<float*>(scratch + align256((size_t)nf * 9 * sizeof(float)));
  cudaStream_t st = (cudaStream_t)stream;
  if (!setup_done) {
    RasterArgs B = A;
    B.N = 1; B.F = 1;
    raster_setup_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(B, 0);
    if (check_launch("raster_setup_kernel")) return 1;
  }
  if (launch_backward(A, nf, st)) return 1;
  const long long nv = (long long)N * V;
  vertex_gather_kernel<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(A.grad_faces, A.grad_textures, csr_offsets,
                                                                      csr_entries, grad_verts, grad_sets, N, V, F,
                                                                      num_sets, detached_sets);
  return check_launch("vertex_gather_kernel");
}


extern "C" int lasr_b200_render_backward(const float* verts, const int32_t* tri, const float* vcolors,
                                         const float* soft_colors, const float* aggrs_info,
                                         const float* grad_soft_colors, const int32_t* csr_offsets,
                                         const int32_t* csr_entries, float* grad_verts, float* grad_vcolors,
                                         void* workspace, size_t workspace_bytes, int N, int V, int F, int S,
                                         float near_, float far_, const float* near_far_dev, float eps,
                                         float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                         int func_id_rgb, int func_id_alpha, int double_side, int num_sets,
                                         int detached_sets, int setup_done, void* stream) {
  if (!vcolors || !grad_vcolors || num_sets < 1 || num_sets > kMaxSets || N <= 0 || V <= 0)
    return fail("render_backward: null colour pointer or bad num_sets");
  return render_backward_impl(verts, tri, stacked_sets(vcolors, N, V, num_sets), soft_colors, aggrs_info, grad_soft_colors,
                              nullptr, csr_offsets, csr_entries, grad_verts, stacked_grads(grad_vcolors, N, V, num_sets),
                              workspace, workspace_bytes, N, V, F, S, near_, far_,
                              near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha,
                              double_side, num_sets, detached_sets, setup_done, stream);
}

extern "C" int lasr_b200_render_backward_packed(const float* verts, const int32_t* tri, const float* vcolors,
                                                const float* packed_grad, const int32_t* csr_offsets,
                                                const int32_t* csr_entries, float* grad_verts, float* grad_vcolors,
                                                void* workspace, size_t workspace_bytes, int N, int V, int F, int S,
                                                float near_, float far_, const float* near_far_dev, float eps,
                                                float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                                int func_id_rgb, int func_id_alpha, int double_side, int num_sets,
                                                int detached_sets, int setup_done, void* stream) {
  if (!packed_grad) return fail("render_backward_packed: null packed gradient");
  if (!vcolors || !grad_vcolors || num_sets < 1 || num_sets > kMaxSets || N <= 0 || V <= 0)
    return fail("render_backward_packed: null colour pointer or bad num_sets");
  return render_backward_impl(verts, tri, stacked_sets(vcolors, N, V, num_sets), nullptr, nullptr, nullptr, packed_grad,
                              csr_offsets, csr_entries, grad_verts, stacked_grads(grad_vcolors, N, V, num_sets),
                              workspace, workspace_bytes, N, V, F, S, near_, far_, near_far_dev,
                              eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side,
                              num_sets, detached_sets, setup_done, stream);
}

extern "C" int lasr_b200_render_backward_sets_packed(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                                     const int* set_shifts, const float* packed_grad,
                                                     const int32_t* csr_offsets, const int32_t* csr_entries,
                                                     float* grad_verts, float* const* grad_set_ptrs, void* workspace,
                                                     size_t workspace_bytes, int N, int V, int F, int S, float near_,
                                                     float far_, const float* near_far_dev, float eps, float sigma_val,
                                                     int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                                     int func_id_alpha, int double_side, int num_sets, int detached_sets,
                                                     int setup_done, void* stream) {
  if (!packed_grad) return fail("render_backward_sets_packed: null packed gradient");
  if (!set_ptrs || !grad_set_ptrs || num_sets < 1 || num_sets > kMaxSets)
    return fail("render_backward_sets_packed: null set array or bad num_sets");
  const ColourSets sets = separate_sets(set_ptrs, set_shifts, num_sets);
  SetGrads g{};
  for (int s = 0; s < num_sets; ++s) { g.p[s] = grad_set_ptrs[s]; g.shift[s] = sets.shift[s]; }
  return render_backward_impl(verts, tri, sets, nullptr, nullptr, nullptr, packed_grad, csr_offsets, csr_entries,
                              grad_verts, g, workspace, workspace_bytes, N, V, F, S, near_, far_, near_far_dev, eps,
                              sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side,
                              num_sets, detached_sets, setup_done, stream);
}

extern "C" size_t lasr_b200_render_backward_scratch_bytes(int N, int F, int S, int num_sets) {
  if (N <= 0 || F <= 0 || S <= 0 || num_sets < 1 || num_sets > kMaxSets) return 0;
  return backward_scratch_bytes(N, F, S, num_sets) + align256((size_t)N * F * 9 * sizeof(float)) +
         align256((size_t)N * F * 9 * num_sets * sizeof(float));
}

extern "C" int lasr_b200_render_forward_sets_compact(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                                     const int* set_shifts, float* aggrs_info, float* soft_colors,
                                                     void* saved, size_t saved_bytes, void* workspace,
                                                     size_t workspace_bytes, int N, int V, int F, int S, float near_,
                                                     float far_, const float* near_far_dev, float eps, float sigma_val,
                                                     int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                                     int func_id_alpha, int double_side, int num_sets,
                                                     const float* background, void* stream) {
  if (!set_ptrs || num_sets < 1 || num_sets > kMaxSets) return fail("render_forward_sets_compact: null set array or bad num_sets");
  return render_forward_impl(verts, tri, separate_sets(set_ptrs, set_shifts, num_sets), aggrs_info, soft_colors,
                             workspace, workspace_bytes, N, V, F, S, near_, far_, near_far_dev, eps, sigma_val,
                             func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side, num_sets,
                             background, stream, saved, saved_bytes, true);
}

extern "C" int lasr_b200_render_backward_sets_packed_compact(const float* verts, const int32_t* tri,
                                                             const float* const* set_ptrs, const int* set_shifts,
                                                             const float* packed_grad, const int32_t* csr_offsets,
                                                             const int32_t* csr_entries, float* grad_verts,
                                                             float* const* grad_set_ptrs, const void* saved,
                                                             size_t saved_bytes, void* workspace, size_t workspace_bytes,
                                                             int N, int V, int F, int S, float near_, float far_,
                                                             const float* near_far_dev, float eps, float sigma_val,
                                                             int func_id_dist, float dist_eps, float gamma_val,
                                                             int func_id_rgb, int func_id_alpha, int double_side,
                                                             int num_sets, int detached_sets, void* stream) {
  if (!packed_grad) return fail("render_backward_sets_packed_compact: null packed gradient");
  if (!set_ptrs || !grad_set_ptrs || num_sets < 1 || num_sets > kMaxSets)
    return fail("render_backward_sets_packed_compact: null set array or bad num_sets");
  const ColourSets sets = separate_sets(set_ptrs, set_shifts, num_sets);
  SetGrads g{};
  for (int s = 0; s < num_sets; ++s) { g.p[s] = grad_set_ptrs[s]; g.shift[s] = sets.shift[s]; }
  return render_backward_impl(verts, tri, sets, nullptr, nullptr, nullptr, packed_grad, csr_offsets, csr_entries,
                              grad_verts, g, workspace, workspace_bytes, N, V, F, S, near_, far_, near_far_dev, eps,
                              sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side,
                              num_sets, detached_sets, 1, stream, const_cast<void*>(saved), saved_bytes, true);
}
