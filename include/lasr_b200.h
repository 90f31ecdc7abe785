/*
 * lasr_b200.h -- C ABI of the B200-native LASR render hot path (liblasr_b200.so).
 *
 * Drop-in boundary for the reference's pybind module `soft_renderer.cuda.soft_rasterize`
 * (third_party/softras/soft_renderer/cuda/soft_rasterize_cuda.cpp:59-76, 94-114, 135-138)
 * and for the dense geometry/loss glue either side of it (nnutils/geom_utils.py:27-71,
 * nnutils/mesh_net.py:264-267, 304-311, 372-447).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 *     `_host`; all tensors are contiguous, float32 / int32, laid out exactly as the reference's
 *     tensors are (shapes in the comments);
 *   - like the reference, the CALLER owns every buffer (soft_rasterize.py:47-62, 88-100); nothing
 *     is allocated inside; scratch comes from `workspace` (size from *_workspace_bytes);
 *   - `stream` is a cudaStream_t passed as void* (the reference launches on the legacy default
 *     stream, .cu:702,717,781; we take the caller's current stream explicitly);
 *   - every entry point returns 0 on success, non-zero on error (and records a message for
 *     lasr_b200_last_error()); CUDA launch errors are reported, not printf'ed and swallowed as in
 *     the reference (.cu:710-712, 741-743, 808-810);
 *   - enum encodings are the reference's (functional/soft_rasterize.py:22-25):
 *       func_id_dist  {hard:0, barycentric:1, euclidean:2}
 *       func_id_rgb   {hard:0, softmax:1}
 *       func_id_alpha {hard:0, sum:1, prod:2}
 *       texture_sample_type {surface:0, vertex:1}
 *   - `dist_eps` is the ALREADY TRANSFORMED value log(1/dist_eps - 1) (soft_rasterize.py:35), as
 *     in the reference's pybind call.
 */
#ifndef LASR_B200_H_
#define LASR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Library identification / diagnostics. */
int lasr_b200_abi_version(void);
const char* lasr_b200_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
uint64_t lasr_b200_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Soft rasterizer.
 * Replaces forward_soft_rasterize / backward_soft_rasterize
 * (soft_rasterize_cuda.cpp:59-91, 94-132; kernels soft_rasterize_cuda_kernel.cu:245-668).
 * ------------------------------------------------------------------------------------------- */

/* Scratch needed by one forward or backward call.  Replaces the reference's `faces_info`
 * [N,F,27] buffer (soft_rasterize.py:47): per face a packed pixel window (8 B), the dilated border
 * box (16 B) and a 176-byte record of derived constants, plus an S-entry pixel-coordinate table.
 * The workspace pointer must be 256-byte aligned (cudaMalloc and framework allocators are): the records are moved
 * by 16-byte-granular bulk copies. */
size_t lasr_b200_raster_workspace_bytes(int batch_size, int num_faces, int image_size);

/* forward_soft_rasterize(faces, textures, faces_info, aggrs_info, soft_colors, image_size, near,
 *                        far, eps, sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb,
 *                        func_id_alpha, texture_sample_type, double_side)
 *   faces        [N,F,9]    in   (x0 y0 z0 x1 y1 z1 x2 y2 z2 per face)
 *   textures     [N,F,T,3]  in   (T = texture_size: R*R texels for 'surface', 3 for 'vertex')
 *   aggrs_info   [N,2,S,S]  out  (softmax: sum,max | hard: depth_min, face_index)
 *   soft_colors  [N,4,S,S]  in/out: RGB planes pre-filled with the background colour by the
 *                                caller (read at .cu:361,364), overwritten with RGBA
 *   near_far_dev optional device pointer to {near, far}; when non-NULL it overrides the two
 *                scalars so LASR's 0-dim CUDA tensors (mesh_net.py:304-311) need no host sync. */
int lasr_b200_raster_forward(const float* faces, const float* textures, float* aggrs_info,
                             float* soft_colors, void* workspace, size_t workspace_bytes,
                             int batch_size, int num_faces, int image_size, int texture_size,
                             float near_, float far_, const float* near_far_dev, float eps,
                             float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                             int func_id_rgb, int func_id_alpha, int texture_sample_type,
                             int double_side, void* stream);

/* The same call with the background colour as an argument: `background_host` is a HOST array of 3 floats and
 * `soft_colors` is output only -- no pre-fill pass over the image by the caller (the three fills of
 * soft_rasterize.py:50-53: 12 S^2 bytes per frame written and read back) and no read of it here.  When the
 * hard z-buffer ('hard' rgb) leaves a pixel uncovered it writes the background itself. */
int lasr_b200_raster_forward_bg(const float* faces, const float* textures, float* aggrs_info,
                                float* soft_colors, void* workspace, size_t workspace_bytes,
                                int batch_size, int num_faces, int image_size, int texture_size,
                                float near_, float far_, const float* near_far_dev, float eps,
                                float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                int func_id_rgb, int func_id_alpha, int texture_sample_type,
                                int double_side, const float* background_host, void* stream);

/* backward_soft_rasterize(faces, textures, soft_colors, faces_info, aggrs_info, grad_faces,
 *                         grad_textures, grad_soft_colors, <same scalars>)
 *   grad_faces    [N,F,9]    out (overwritten; the reference accumulates into a zeroed buffer)
 *   grad_textures [N,F,T,3]  out (overwritten)
 * Deterministic for 'vertex' textures (LASR's mode): every face's gradient is a private, fixed-order sum (no
 * float atomics, unlike .cu:606,622,657-666).  'surface' texel gradients are still scattered with atomicAdd
 * (several faces' pixels never share a texel, but several pixels of one face do). */
int lasr_b200_raster_backward(const float* faces, const float* textures, const float* soft_colors,
                              const float* aggrs_info, float* grad_faces, float* grad_textures,
                              const float* grad_soft_colors, void* workspace,
                              size_t workspace_bytes, int batch_size, int num_faces,
                              int image_size, int texture_size, float near_, float far_,
                              const float* near_far_dev, float eps, float sigma_val,
                              int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                              int func_id_alpha, int texture_sample_type, int double_side,
                              int setup_done, void* stream);
/* setup_done = 1: `workspace` is the one the matching lasr_b200_raster_forward call filled for the SAME
 * inputs (the reference likewise hands its forward's `faces_info` to the backward,
 * soft_rasterize.py:64,71); the per-face records are then read back instead of recomputed. */

/* ---------------------------------------------------------------------------------------------
 * Indexed ("fused") rasterizer: vertices + shared topology + vertex colours in, image out, gradients
 * back per vertex.  Replaces the gather -> rasterize -> scatter chain of SoftRenderer.render_mesh for
 * 'vertex' textures (mesh.py:105-109,128-134 `face_vertices` / `face_textures` = vertices[faces],
 * functional/face_vertices.py:4-22, and their autograd scatter) so that the [N,F,9] gathered
 * tensors never exist in HBM.
 *   verts   [N,V,3]  clip-space vertices (what SoftRasterizer receives as mesh.vertices)
 *   tri     [F,3]    int32 topology shared by the N frames
 *   vcolors [K,N,V,3] vertex colours: num_sets = K (1, 2 or 3) colour sets interpolated in ONE geometry
 *                    pass (K > 1 needs LASR's configuration: euclidean / softmax / prod / fill_back).
 *                    LASR renders the same posed mesh three times per frame with different vertex
 *                    attributes -- forward flow, backward flow, texture (nnutils/mesh_net.py:318-363);
 *                    with K sets those share the per-pixel distance / depth-softmax work.
 *   soft_colors / grad_soft_colors [N,3K+1,S,S]: planes 3s..3s+2 = set s, plane 3K = alpha (K = 1 is
 *                    the reference's RGBA layout); soft_colors arrives pre-filled with the backgrounds
 *   csr_offsets [V+1], csr_entries [3F] (entry = face*3 + corner, grouped by vertex, ascending):
 *                    vertex -> corner adjacency used to sum per-face gradients in a fixed order
 *   background      HOST array of 3K floats (rgb per set), or NULL when soft_colors arrives pre-filled
 *                    with the background the way the reference's op expects it (soft_rasterize.py:50-53)
 *   grad_verts [N,V,3], grad_vcolors [K,N,V,3] out (overwritten)
 *   detached_sets (backward): 0, or 1 with K >= 2 -- colour set 0 is rendered for reading only, as if
 *                    its planes were `.detach()`ed: planes 0-2 of grad_soft_colors are not read and
 *                    grad_vcolors[0] is zero (LASR re-projects the frame's own positions only under
 *                    .detach(), nnutils/mesh_net.py:102-103)
 * Other arguments as lasr_b200_raster_forward/backward (texture_sample_type is 'vertex').
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_render_workspace_bytes(int batch_size, int num_faces, int image_size, int num_sets);

int lasr_b200_render_forward(const float* verts, const int32_t* tri, const float* vcolors,
                             float* aggrs_info, float* soft_colors, void* workspace,
                             size_t workspace_bytes, int batch_size, int num_vertices, int num_faces,
                             int image_size, float near_, float far_, const float* near_far_dev,
                             float eps, float sigma_val, int func_id_dist, float dist_eps,
                             float gamma_val, int func_id_rgb, int func_id_alpha, int double_side,
                             int num_sets, const float* background, void* stream);

int lasr_b200_render_backward(const float* verts, const int32_t* tri, const float* vcolors,
                              const float* soft_colors, const float* aggrs_info,
                              const float* grad_soft_colors, const int32_t* csr_offsets,
                              const int32_t* csr_entries, float* grad_verts, float* grad_vcolors,
                              void* workspace, size_t workspace_bytes, int batch_size,
                              int num_vertices, int num_faces, int image_size, float near_,
                              float far_, const float* near_far_dev, float eps, float sigma_val,
                              int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                              int func_id_alpha, int double_side, int num_sets, int detached_sets,
                              int setup_done, void* stream);

/* lasr_b200_render_backward fed by lasr_b200_head_backward_packed: `packed_grad` replaces soft_colors, aggrs_info and
 * grad_soft_colors (num_sets must be 3 and detached_sets 1).  Same outputs. */
int lasr_b200_render_backward_packed(const float* verts, const int32_t* tri, const float* vcolors,
                                     const float* packed_grad, const int32_t* csr_offsets,
                                     const int32_t* csr_entries, float* grad_verts, float* grad_vcolors,
                                     void* workspace, size_t workspace_bytes, int batch_size,
                                     int num_vertices, int num_faces, int image_size, float near_,
                                     float far_, const float* near_far_dev, float eps, float sigma_val,
                                     int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                     int func_id_alpha, int double_side, int num_sets, int detached_sets,
                                     int setup_done, void* stream);

/* The colour sets as SEPARATE tensors with a frame shift each, instead of one stacked [K][N,V,3] tensor:
 *   set_ptrs  HOST array of num_sets device pointers, each [N,V,3]
 *   set_shifts HOST array of num_sets ints in [0, N) (NULL = all 0): frame n renders frame (n + shift) mod N of the set
 * LASR's flow render colours every frame with its own camera-space vertices AND with its pair's -- frame i and frame
 * i +- bs of the SAME tensor (nnutils/mesh_net.py:84-87, 318-335).  Stacked, that is a roll and a concatenation per step
 * (and their adjoints); here both sets point at the one tensor, the second with shift bs.  The backward writes set s's
 * gradient [N,V,3] through grad_set_ptrs[s] (HOST array; NULL entries allowed for detached sets) in the frame order of
 * the set's own tensor, i.e. already shifted back. */
int lasr_b200_render_forward_sets(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                  const int* set_shifts, float* aggrs_info, float* soft_colors, void* workspace,
                                  size_t workspace_bytes, int batch_size, int num_vertices, int num_faces,
                                  int image_size, float near_, float far_, const float* near_far_dev, float eps,
                                  float sigma_val, int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                  int func_id_alpha, int double_side, int num_sets, const float* background,
                                  void* stream);
int lasr_b200_render_backward_sets_packed(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                          const int* set_shifts, const float* packed_grad,
                                          const int32_t* csr_offsets, const int32_t* csr_entries, float* grad_verts,
                                          float* const* grad_set_ptrs, void* workspace, size_t workspace_bytes,
                                          int batch_size, int num_vertices, int num_faces, int image_size, float near_,
                                          float far_, const float* near_far_dev, float eps, float sigma_val,
                                          int func_id_dist, float dist_eps, float gamma_val, int func_id_rgb,
                                          int func_id_alpha, int double_side, int num_sets, int detached_sets,
                                          int setup_done, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Compact mode: 80 bytes per face kept from forward to backward instead of the workspace.
 * The workspace of the calls above holds, per face, the full record the forward's inner loop reads (176 - 240 bytes),
 * the dilated box and the pixel window: 200 - 264 bytes per face that a caller who reuses it in the backward
 * (`setup_done = 1`) keeps alive in between -- the reference keeps 108 (`faces_info`, soft_rasterize_cuda.cpp:59-76).
 * The *_compact calls write, next to it, a SAVED buffer (lasr_b200_raster_saved_bytes: 80 bytes per face + the S-entry
 * coordinate table; 256-byte aligned) holding only what costs a division to rebuild (inverse matrix, 1/z, 1/edge^2, the
 * edge-prediction threshold, flags); the workspace is then scratch that may be freed when the forward call returns, and
 * the backward takes the saved buffer plus a small scratch of its own (lasr_b200_raster_backward_scratch_bytes /
 * lasr_b200_render_backward_scratch_bytes), reads the vertices and colours from the op's inputs and re-derives the edge
 * rows and the box with the forward's own expressions: same gradients bit for bit.
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_raster_saved_bytes(int batch_size, int num_faces, int image_size);
size_t lasr_b200_raster_backward_scratch_bytes(int batch_size, int num_faces, int image_size);
size_t lasr_b200_render_backward_scratch_bytes(int batch_size, int num_faces, int image_size, int num_sets);

int lasr_b200_raster_forward_compact(const float* faces, const float* textures, float* aggrs_info,
                                     float* soft_colors, void* saved, size_t saved_bytes, void* workspace,
                                     size_t workspace_bytes, int batch_size, int num_faces, int image_size,
                                     int texture_size, float near_, float far_, const float* near_far_dev, float eps,
                                     float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                     int func_id_rgb, int func_id_alpha, int texture_sample_type, int double_side,
                                     const float* background_host, void* stream);
int lasr_b200_raster_backward_compact(const float* faces, const float* textures, const float* soft_colors,
                                      const float* aggrs_info, float* grad_faces, float* grad_textures,
                                      const float* grad_soft_colors, const void* saved, size_t saved_bytes,
                                      void* workspace, size_t workspace_bytes, int batch_size, int num_faces,
                                      int image_size, int texture_size, float near_, float far_,
                                      const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                                      float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                      int texture_sample_type, int double_side, void* stream);
int lasr_b200_render_forward_sets_compact(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                          const int* set_shifts, float* aggrs_info, float* soft_colors, void* saved,
                                          size_t saved_bytes, void* workspace, size_t workspace_bytes, int batch_size,
                                          int num_vertices, int num_faces, int image_size, float near_, float far_,
                                          const float* near_far_dev, float eps, float sigma_val, int func_id_dist,
                                          float dist_eps, float gamma_val, int func_id_rgb, int func_id_alpha,
                                          int double_side, int num_sets, const float* background, void* stream);
int lasr_b200_render_backward_sets_packed_compact(const float* verts, const int32_t* tri, const float* const* set_ptrs,
                                                  const int* set_shifts, const float* packed_grad,
                                                  const int32_t* csr_offsets, const int32_t* csr_entries,
                                                  float* grad_verts, float* const* grad_set_ptrs, const void* saved,
                                                  size_t saved_bytes, void* workspace, size_t workspace_bytes,
                                                  int batch_size, int num_vertices, int num_faces, int image_size,
                                                  float near_, float far_, const float* near_far_dev, float eps,
                                                  float sigma_val, int func_id_dist, float dist_eps, float gamma_val,
                                                  int func_id_rgb, int func_id_alpha, int double_side, int num_sets,
                                                  int detached_sets, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused linear-blend skinning + body transform + pinhole projection + batch depth window
 * (SURVEY 8a rows G1, G2, G3).  Replaces, for CUDA tensors,
 *   obj_to_cam   nnutils/geom_utils.py:45-71     cam = (sum_k w_k (v R_k + T_k)) R_0 + T_0
 *   pinhole_cam  nnutils/geom_utils.py:27-34     x,y <- pp + xy * f / z
 *   near/far     nnutils/mesh_net.py:304-311     near = dmin-(dmax-dmin)/2, far = dmax+(dmax-dmin)/2
 *
 *   verts [Nv,V,3]  Nv = N, or 1 with verts_shared=1 (one mesh for every frame)
 *   Rmat  [N*B,3,3], Tmat [N*B,3]      bone 0 of each frame is the body-to-camera transform
 *   skin  [Ns,B-1,V] Ns = N, or 1 with skin_shared=1; NULL when B == 1
 *   pp [N,2], fl [N] per-frame principal point / focal length; both NULL = no projection
 *   cam [N,V,3] out (skinned only when tocam == 0);  px [N,V,4] out = (x, y, z, 1)
 *   near_far_dev [2] out, optional: depth window of the whole batch, left on the device
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_lbs_workspace_bytes(int num_frames, int num_vertices, int num_bones);

int lasr_b200_lbs_project_forward(const float* verts, const float* Rmat, const float* Tmat,
                                  const float* skin, const float* pp, const float* fl, float* cam,
                                  float* px, float* near_far_dev, void* workspace,
                                  size_t workspace_bytes, int num_frames, int num_vertices,
                                  int num_bones, int verts_shared, int skin_shared, int tocam,
                                  void* stream);

/* Gradients of the above.  grad_cam [N,V,3] and/or grad_px [N,V,4] in (either may be NULL);
 * grad_verts [N,V,3] and grad_skin [N,B-1,V] are written PER FRAME (callers sum over frames for a
 * shared mesh / shared weights); grad_R [N*B,3,3], grad_T [N*B,3], grad_pp [N,2], grad_fl [N].
 * Deterministic (fixed-order block reductions, no float atomics). */
int lasr_b200_lbs_project_backward(const float* verts, const float* Rmat, const float* Tmat,
                                   const float* skin, const float* pp, const float* fl,
                                   const float* grad_cam, const float* grad_px, float* grad_verts,
                                   float* grad_R, float* grad_T, float* grad_skin, float* grad_pp,
                                   float* grad_fl, void* workspace, size_t workspace_bytes,
                                   int num_frames, int num_vertices, int num_bones,
                                   int verts_shared, int skin_shared, int tocam, void* stream);

/* The general pose entry points: lasr_b200_lbs_project_* plus two optional pieces of LASR's per-step glue done inside
 * the kernels.
 * (1) rest_ts != NULL: the part transforms are COMPOSED ABOUT THEIR REST POSITIONS (nnutils/mesh_net.py:278-281:
 *     Tmat[:,1:] = -Rmat[:,1:] rest + Tmat[:,1:] + rest; Rmat[:,1:] transposed), which the reference does with ~8 small
 *     PyTorch kernels forward and ~20 backward per step:
 *       rest_ts   [n_hypo*(B-1), 3]  joint centres, shared by the frames (frame n uses hypothesis n % n_hypo)
 *       Rmat/Tmat the RAW network outputs (mesh_net.py:259-260), bone 0 the body transform, used as given
 *       grad_R / grad_T are gradients w.r.t. those raw inputs; grad_rest [N, B-1, 3] is per frame (the caller sums the
 *       frames of a hypothesis).
 * (2) view != NULL: a third output [N,V,3], the projected vertex as LASR hands it to the rasterizer -- + eye offset,
 *     y flip, then SoftRenderer's look_at(eye) with the eye at (0, 0, eye_z) and orthographic scale 1
 *     (mesh_net.py:81-83,354-357; transform.py:41-48): (x, -y, (z + eye_z) - eye_z); grad_view is its upstream gradient. */
int lasr_b200_lbs_pose_forward(const float* verts, const float* Rmat, const float* Tmat,
                               const float* rest_ts, const float* skin, const float* pp, const float* fl,
                               float* cam, float* px, float* view, float eye_z, float* near_far_dev,
                               void* workspace, size_t workspace_bytes, int N, int V, int B, int n_hypo,
                               int verts_shared, int skin_shared, int tocam, void* stream);

int lasr_b200_lbs_pose_backward(const float* verts, const float* Rmat, const float* Tmat,
                                const float* rest_ts, const float* skin, const float* pp, const float* fl,
                                const float* grad_cam, const float* grad_px, const float* grad_view,
                                float* grad_verts, float* grad_R, float* grad_T, float* grad_rest,
                                float* grad_skin, float* grad_pp, float* grad_fl, void* workspace,
                                size_t workspace_bytes, int N, int V, int B, int n_hypo, int verts_shared,
                                int skin_shared, int tocam, void* stream);

/* Sums over frames of the per-frame gradients of parameters shared by every frame (lasr_b200_lbs_*_backward writes
 * grad_verts [N,V,3], grad_skin [N,B-1,V] and grad_rest [N,B-1,3] per frame), one launch for the three:
 *   sum_verts [V,3], sum_skin [B-1,V], sum_rest [n_hypo,B-1,3] (frames laid out [.., n_hypo]: frame n belongs to
 *   hypothesis n % n_hypo).  Any input may be NULL together with its output.  Frames are added in order: deterministic. */
int lasr_b200_lbs_sum_frames(const float* grad_verts, const float* grad_skin, const float* grad_rest,
                             float* sum_verts, float* sum_skin, float* sum_rest, int num_frames, int num_vertices,
                             int num_bones, int n_hypo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Image losses: one fused masked-reduction pass.  Replaces the Python double
 * loops of nnutils/mesh_net.py:372-447 (silhouette :374-390, flow :393-416, texture L1 :420-441;
 * the 0.005*LPIPS term is outside this path).
 *   mask_pred [N,H,S,S]; flow_rd [N,H,S,S,2] + bgmask [N,H,S,S] uint8 + flow_obs [N,3,S,S] (NULL = no
 *   flow term); tex [N,H,3,S,S] + fg [N,H,S,S] + imgs [N,3,S,S] (NULL = no texture term);
 *   masks, occ [N,S,S].
 *   sums [N,H,7] out, per (frame, hypothesis):
 *     0: sum_v (mask_pred-mask)^2   1: sum_k s|flow_rd-flow_obs|   2: sum_k s   3: sum_k 1
 *     4: sum_v mean_c|obs-tex*fg|   5: sum_v mean_c|white-tex|     6: sum_v 1
 *   with v = occ!=0, k = !bgmask & v & mask>0, s = sigmoid(-occ).
 * The handful of scalar operations from sums to losses live in lasr_b200/losses.py.
 * backward: grad_sums [N,H,7] -> per-pixel gradients (any output may be NULL).  Deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_losses_workspace_bytes(int num_frames, int num_hypotheses, int image_size);

int lasr_b200_losses_forward(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                             const float* tex, const float* fg, const float* imgs, const float* masks,
                             const float* occ, const float* flow_obs, float* sums, void* workspace,
                             size_t workspace_bytes, int num_frames, int num_hypotheses, int image_size,
                             void* stream);

/* The same plus the scalar step of lasr_b200_losses_finalize in the launch that sums the block partials: sums [N,H,7],
 * sub [3][N*H] and out4 [4] out (both given or both NULL): out4 = the three means of lasr_b200_losses_finalize (same
 * bits as the two calls) and their sum (mask + flow) + texture, the objective LASR forms of them. */
int lasr_b200_losses_forward_total(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                                   const float* tex, const float* fg, const float* imgs, const float* masks,
                                   const float* occ, const float* flow_obs, float* sums, float* sub, float* out4,
                                   float l1tex_wt, void* workspace, size_t workspace_bytes, int num_frames,
                                   int num_hypotheses, int image_size, void* stream);

int lasr_b200_losses_backward(const float* mask_pred, const float* flow_rd, const uint8_t* bgmask,
                              const float* tex, const float* fg, const float* imgs, const float* masks,
                              const float* occ, const float* flow_obs, const float* grad_sums,
                              float* grad_mask_pred, float* grad_flow_rd, float* grad_tex, float* grad_fg,
                              int num_frames, int num_hypotheses, int image_size, void* stream);

/* sums [N,H,7] -> sub [3][N*H] (per-pair mask / flow / texture losses) and out3 [3] (their means), the
 * scalar arithmetic of mesh_net.py:388-390, 404-416, 440-445 in one tiny kernel; and its adjoint:
 * grad_out3 [3] (+ optional grad_sub [3][N*H]) -> grad_sums [N,H,7] for lasr_b200_losses_backward. */
int lasr_b200_losses_finalize(const float* sums, float* sub, float* out3, int num_frames, int num_hypotheses,
                              float l1tex_wt, void* stream);
int lasr_b200_losses_coefficients(const float* sums, const float* grad_out3, const float* grad_sub,
                                  float* grad_sums, int num_frames, int num_hypotheses, float l1tex_wt,
                                  void* stream);
/* ... for the four-entry output of the *_total calls: grad_out4 [4], the last entry being the gradient of the sum. */
int lasr_b200_losses_coefficients_total(const float* sums, const float* grad_out4, const float* grad_sub,
                                        float* grad_sums, int num_frames, int num_hypotheses, float l1tex_wt,
                                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Flow epilogue: the post-processing of render_flow_soft_2,
 * nnutils/mesh_net.py:88-104, in one pass.  px [2n,4,S,S] is the rasterizer's output for the stacked
 * batch (first n: frame-0 positions, last n: frame-1 positions); pp0/pp1 [n,2], f0/f1 [n].
 * flow [n,S,S,2] and bgmask [n,S,S] (uint8) out.  backward: grad_flow -> grad_px [2n,4,S,S] (only the
 * frame-1 half carries gradient, p0 is detached and background pixels are cut, :102-103), grad_pp1,
 * grad_f1 (pp0 / f0 get none).
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_flow_epilogue_workspace_bytes(int num_frames, int image_size);
int lasr_b200_flow_epilogue_forward(const float* px, const float* pp0, const float* pp1, const float* f0,
                                    const float* f1, float* flow, uint8_t* bgmask, int num_frames,
                                    int image_size, void* stream);
int lasr_b200_flow_epilogue_backward(const float* px, const float* f1, const float* grad_flow, float* grad_px,
                                     float* grad_pp1, float* grad_f1, void* workspace, size_t workspace_bytes,
                                     int num_frames, int image_size, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused image head of the three-set render (SURVEY 8a rows R1 epilogue + L1-L3 in one pass).
 * `image` [N*H,10,S,S] is the output of lasr_b200_render_forward with num_sets = 3 and colour sets
 * (own camera-space positions, the paired frame's positions, texture): planes 0-2 / 3-5 / 6-8 and the
 * silhouette in plane 9.  Per frame it re-projects both position sets with the frame's own and its
 * pair's intrinsics (render_flow_soft_2's post-processing, nnutils/mesh_net.py:88-104; the forward AND
 * the backward flow, :318-335, in one call because frame i of the first half pairs with i+bs and vice
 * versa) and accumulates the seven masked sums of lasr_b200_losses_forward (mesh_net.py:372-447).
 *   pp_own, pp_pair [N*H,2]; f_own, f_pair [N*H]; imgs, masks, occ, flow_obs as lasr_b200_losses_forward
 *   sums [N,H,7] out; flow_rd [N*H,S,S,2] and bgmask [N*H,S,S] (uint8) optional outputs (both or none)
 * backward: grad_sums [N,H,7] (from lasr_b200_losses_coefficients) -> grad_image [N*H,10,S,S] (all
 * planes written), grad_pp_pair [N*H,2], grad_f_pair [N*H].  Deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_head_workspace_bytes(int N, int H, int image_size);

int lasr_b200_head_forward(const float* image, const float* pp_own, const float* pp_pair,
                           const float* f_own, const float* f_pair, const float* imgs,
                           const float* masks, const float* occ, const float* flow_obs, float* sums,
                           float* flow_rd, uint8_t* bgmask, void* workspace, size_t workspace_bytes,
                           int N, int H, int image_size, void* stream);

/* head_forward + lasr_b200_losses_finalize, the scalar step riding the partial-sum launch (sub [3][N*H]; out4 [4] as in
 * lasr_b200_losses_forward_total).  pair_shift in [0, N*H): frame ij takes entry (ij + pair_shift) mod N*H of pp_pair /
 * f_pair -- LASR pairs frame i with frame i +- bs, so the caller passes the OWN arrays again with pair_shift = bs*H
 * instead of rolled copies (0: pp_pair / f_pair are indexed like pp_own / f_own). */
int lasr_b200_head_forward_total(const float* image, const float* pp_own, const float* pp_pair,
                                 const float* f_own, const float* f_pair, const float* imgs,
                                 const float* masks, const float* occ, const float* flow_obs, float* sums,
                                 float* sub, float* out4, float l1tex_wt, float* flow_rd, uint8_t* bgmask,
                                 void* workspace, size_t workspace_bytes, int N, int H, int image_size,
                                 int pair_shift, void* stream);

int lasr_b200_head_backward(const float* image, const float* pp_own, const float* pp_pair,
                            const float* f_own, const float* f_pair, const float* imgs,
                            const float* masks, const float* occ, const float* flow_obs,
                            const float* grad_sums, float* grad_image, float* grad_pp_pair,
                            float* grad_f_pair, void* workspace, size_t workspace_bytes, int N, int H,
                            int image_size, void* stream);

/* The same backward with the image gradient in PACKED per-pixel form for lasr_b200_render_backward_packed (K = 3 colour
 * sets, set 0 detached -- LASR's hot path): packed_grad [N*H][3][S*S] float4 =
 *   { g_alpha * (1 - alpha), sum_k g_k * image_k (six live colour planes), 1 / softmax sum, softmax max },
 *   { g3, g4, g5, g6 }, { g7, g8, 0, 0 }
 * aggrs_info [N*H,2,S,S] is the rasterizer forward's.  The ten planar gradient planes are never written, and a
 * (pixel, face) pair of the rasterizer's backward reads three coalesced 128-bit words instead of sixteen scalars. */
int lasr_b200_head_backward_packed(const float* image, const float* pp_own, const float* pp_pair,
                                   const float* f_own, const float* f_pair, const float* imgs,
                                   const float* masks, const float* occ, const float* flow_obs,
                                   const float* grad_sums, const float* aggrs_info, float* packed_grad,
                                   float* grad_pp_pair, float* grad_f_pair, void* workspace,
                                   size_t workspace_bytes, int N, int H, int image_size, void* stream);
/* ... with the pair_shift of lasr_b200_head_forward_total: entry (ij + pair_shift) mod N*H of grad_pp_pair / grad_f_pair
 * receives frame ij's gradient, i.e. the gradients come out in the order of the arrays that were read. */
int lasr_b200_head_backward_packed_shift(const float* image, const float* pp_own, const float* pp_pair,
                                         const float* f_own, const float* f_pair, const float* imgs,
                                         const float* masks, const float* occ, const float* flow_obs,
                                         const float* grad_sums, const float* aggrs_info, float* packed_grad,
                                         float* grad_pp_pair, float* grad_f_pair, void* workspace,
                                         size_t workspace_bytes, int N, int H, int image_size, int pair_shift,
                                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Skinning weights, nnutils/mesh_net.py:264-267:
 *   skin[h,b,v] = softmax_b( -10 * sum_d exp(log_ctl[h,b,d]) * ((ctl_ts[h,b] - verts[h,v]) . R(ctl_rs[h,b]))_d^2 )
 * with R = kornia's quaternion_to_rotation_matrix (XYZW, normalised first).
 *   ctl_ts, log_ctl [H*Bp,3]; ctl_rs [H*Bp,4]; verts [H,V,3] (the caller's detached mean shape)
 *   skin [H,Bp,V] out -- the reference's [H,B-1,V,1] without the trailing 1; Bp = B-1 part bones, <= 128
 * backward: grad_skin [H,Bp,V] -> grad_ctl_ts, grad_log_ctl [H*Bp,3], grad_ctl_rs [H*Bp,4] (overwritten).
 * Deterministic (fixed-order partial sums).
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_skin_workspace_bytes(int n_hypo, int n_part_bones, int num_vertices);

int lasr_b200_skin_weights_forward(const float* ctl_ts, const float* ctl_rs, const float* log_ctl,
                                   const float* verts, float* skin, int n_hypo, int n_part_bones,
                                   int num_vertices, void* stream);

int lasr_b200_skin_weights_backward(const float* ctl_ts, const float* ctl_rs, const float* log_ctl,
                                    const float* verts, const float* skin, const float* grad_skin,
                                    float* grad_ctl_ts, float* grad_ctl_rs, float* grad_log_ctl,
                                    void* workspace, size_t workspace_bytes, int n_hypo,
                                    int n_part_bones, int num_vertices, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mesh regularisers on the sparse topology.  The reference evaluates them through
 * dense V x V matrices; these take a CSR vertex adjacency / an edge list built once per topology.
 *   nbr_offsets [V+1], nbr [E2]: unique neighbours of every vertex, ascending (E2 = directed edges =
 *                    non-zeros of the reference's adjacency matrix)
 *   LaplacianLoss (third_party/ext_nnutils/loss_utils.py:34-65):  lap = L x with L = I - D^-1 A (rows of
 *                    isolated vertices zero), loss[n] = sum lap^2.   x, lap, grad_x [N,V,3]; loss [N]
 *   ARAPLoss (nnutils/loss_utils.py:29-64):  loss[n] = mean over directed edges (a,b) of
 *                    | |x_b-x_a|^2 - |dx_b-dx_a|^2 |
 *   FlattenLoss (third_party/ext_nnutils/loss_utils.py:67-152):  edges [E,4] = (v0, v1, v2, v3): an edge
 *                    and the opposite vertices of its two faces; loss[n] = sum_e (cos(dihedral)+1)^2;
 *                    inc_offsets [V+1], inc [4E] (entry = edge*4 + slot, grouped by vertex, ascending):
 *                    vertex -> edge-corner incidence used to sum the per-edge gradients in a fixed order
 * Every gradient is formed in gather form: no atomics, deterministic.
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_meshreg_workspace_bytes(int N, int num_vertices, int num_edges);

int lasr_b200_laplacian_forward(const float* x, const int32_t* nbr_offsets, const int32_t* nbr, float* lap,
                                float* loss, void* workspace, size_t workspace_bytes, int N,
                                int num_vertices, void* stream);
int lasr_b200_laplacian_backward(const float* lap, const float* grad_loss, const int32_t* nbr_offsets,
                                 const int32_t* nbr, float* grad_x, int N, int num_vertices, void* stream);

int lasr_b200_arap_forward(const float* dx, const float* x, const int32_t* nbr_offsets, const int32_t* nbr,
                           float* loss, void* workspace, size_t workspace_bytes, int N, int num_vertices,
                           int num_directed_edges, void* stream);
int lasr_b200_arap_backward(const float* dx, const float* x, const float* grad_loss,
                            const int32_t* nbr_offsets, const int32_t* nbr, float* grad_dx, float* grad_x,
                            int N, int num_vertices, int num_directed_edges, void* stream);

int lasr_b200_flatten_forward(const float* x, const int32_t* edges, float* loss, void* workspace,
                              size_t workspace_bytes, int N, int num_vertices, int num_edges, float eps,
                              void* stream);
int lasr_b200_flatten_backward(const float* x, const int32_t* edges, const float* grad_loss,
                               const int32_t* inc_offsets, const int32_t* inc, float* grad_x,
                               void* workspace, size_t workspace_bytes, int N, int num_vertices,
                               int num_edges, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer3D nearest neighbours.
 * Replaces the pybind module `chamfer_3D` (third_party/chamfer3D/chamfer_cuda.cpp:9-32; kernels
 * chamfer3D.cu:12-196) behind chamfer_3DFunction (third_party/chamfer3D/dist_chamfer_3D.py:27-63);
 * LASR's call site: nnutils/mesh_net.py:474-478, scripts/eval_mesh.py:168.
 *   forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
 *     xyz1 [B,n,3], xyz2 [B,m,3] in;  dist1 [B,n] = min_k |xyz1_j - xyz2_k|^2 (SQUARED distance),
 *     idx1 [B,n] int32 = the lowest k attaining it;  dist2 [B,m], idx2 [B,m] the same with the roles swapped.
 *   backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
 *     gradxyz1 [B,n,3], gradxyz2 [B,m,3] are fully OVERWRITTEN (the reference accumulates into
 *     caller-zeroed buffers, dist_chamfer_3D.py:54-58; same values).  The matched-target terms are
 *     scattered with float atomics, as in the reference (chamfer3D.cu:169-174).
 * ------------------------------------------------------------------------------------------- */
int lasr_b200_chamfer_forward(const float* xyz1, const float* xyz2, float* dist1, float* dist2,
                              int32_t* idx1, int32_t* idx2, int batch_size, int n, int m, void* stream);
int lasr_b200_chamfer_backward(const float* xyz1, const float* xyz2, const float* grad_dist1,
                               const float* grad_dist2, const int32_t* idx1, const int32_t* idx2,
                               float* grad_xyz1, float* grad_xyz2, int batch_size, int n, int m,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * Texture atlas kernels of the SoftRas extension family.
 * Replace the pybind modules `soft_renderer.cuda.load_textures` (load_textures_cuda.cpp:19-36, kernel
 * load_textures_cuda_kernel.cu:8-66; caller functional/load_obj.py:100) and
 * `soft_renderer.cuda.create_texture_image` (create_texture_image_cuda.cpp:17-34, kernel
 * create_texture_image_cuda_kernel.cu:10-73; caller functional/save_obj.py:27).
 *   load_textures(image, faces, textures, is_update):
 *     image [H,W,3] (already flipped, 0..1), faces [F,3,2] texture coordinates, is_update [F] int32,
 *     textures [F,R*R,3] in/out: faces with is_update != 0 are resampled bilinearly, the rest kept.
 *   create_texture_image(vertices_all, textures, image, eps):
 *     faces [F,3,2] texture coordinates in ATLAS PIXELS, textures [F,R*R,3], image [H,W,3] out
 *     (cells of faces >= F are left untouched), tile_width = int(sqrt(F-1))+1, cell = W / tile_width.
 * ------------------------------------------------------------------------------------------- */
int lasr_b200_load_textures(const float* image, const float* faces, const int32_t* is_update,
                            float* textures, int num_faces, int texture_res, int image_height,
                            int image_width, void* stream);
int lasr_b200_create_texture_image(const float* faces, const float* textures, float* image,
                                   int num_faces, int texture_res_in, int image_height, int image_width,
                                   float eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dataloader -> device staging.
 * Replaces the per-item preprocessing of BaseDataset.__getitem__ (dataloader/vidbase.py:52-193) between file
 * decode and collate -- complement-colour background, mask-box crop (cv2.remap at integer offsets), cv2.resize
 * to img_size, flow re-normalisation between the two crops, focal scale and principal-point offsets -- for
 * `num_pairs` dataset items at once.  A pair is frame 0 and its partner frame n.
 *   in   frames_bgr [P,2,H,W,3] u8 (cv2.imread order), masks [P,2,H,W] u8, flows [P,2,H,W,3] f32 (readPFM:
 *        u, v, valid; slot 0 holds the flow 0->n, slot 1 the flow n->0), occs [P,2,H,W] f32
 *   out  img [P,2,3,S,S] RGB 0..1 (elem 'img' / 'imgn'), mask [P,2,S,S] 0/1 ('mask'), flow [P,2,3,S,S]
 *        ('flow' / 'flown'), occ [P,2,S,S] ('occ' / 'occn'), cam_focal [P,2] (cam[0] = 1/alp), pps [P,2,2],
 *        boxes [P,2,4] int32 = (center x, center y, half size, mask pixel count); a frame whose mask is empty
 *        has count 0 and its pair's outputs are left untouched (the reference raises there).
 * ------------------------------------------------------------------------------------------- */
size_t lasr_b200_stage_workspace_bytes(int num_pairs);
int lasr_b200_stage_pairs(const uint8_t* frames_bgr, const uint8_t* masks, const float* flows,
                          const float* occs, float* img, float* mask_out, float* flow_out,
                          float* occ_out, float* cam_focal, float* pps, int32_t* boxes,
                          void* workspace, size_t workspace_bytes, int num_pairs, int height,
                          int width, int image_size, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LASR_B200_H_ */
