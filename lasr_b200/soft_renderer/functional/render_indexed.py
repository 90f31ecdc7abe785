"""Indexed ("fused") rasterization: ``soft_rasterize(vertices[faces], textures[faces], ...)`` without
the two gathers and their autograd scatter -- the kernels fetch vertices / colours through the
topology themselves and hand back per-vertex gradients (lasr_b200_render_forward/backward).

Semantically identical to ``srf.soft_rasterize(srf.face_vertices(v, f), srf.face_vertices(c, f),
..., texture_type='vertex')`` (mesh.py:105-109,128-134 + functional/soft_rasterize.py) when every
frame shares one topology, which is how LASR calls it (faces repeated per frame,
nnutils/mesh_net.py:77,357)."""
from __future__ import annotations

import sys

import numpy as np
import torch
from torch.autograd import Function

from ... import _cabi
from .soft_rasterize import ALPHA_FUNCS, DIST_FUNCS, RGB_FUNCS, SoftRasterizeFunction, _near_far

import weakref

# Caches are keyed by the identity of the tensor OBJECT (id + a weak reference that must still resolve to the
# same object) and its in-place version counter -- not by data_ptr, which the allocator may hand to a
# different tensor of the same shape later.
_topology_cache = {}
_csr_cache = {}


def _cache_get(cache, t, extra=()):
    hit = cache.get(id(t))
    if hit is not None and hit[0]() is t and hit[1] == (t._version,) + tuple(extra):
        return hit[2]
    return None


def _cache_put(cache, t, value, extra=()):
    key = id(t)
    cache[key] = (weakref.ref(t, lambda _r, k=key, c=cache: c.pop(k, None)), (t._version,) + tuple(extra), value)
    return value


def shared_topology(faces):
    """[N,F,3] faces -> [F,3] int32 if all frames share one topology, else None.  The (one-off,
    synchronising) equality check is cached per tensor object."""
    if faces.dim() == 2:
        return faces if (faces.dtype == torch.int32 and faces.is_contiguous()) else faces.to(torch.int32).contiguous()
    hit = _cache_get(_topology_cache, faces)
    if hit is None:
        same = bool((faces == faces[:1]).all()) if faces.shape[0] > 1 else True
        hit = _cache_put(_topology_cache, faces, faces[0].to(torch.int32).contiguous() if same else False)
    return hit if hit is not False else None


def vertex_adjacency(tri, num_vertices):
    """vertex -> (face, corner) adjacency as CSR; entries of a vertex are in ascending face order so
    the gradient sum has a fixed order.  Cached per topology tensor object."""
    hit = _cache_get(_csr_cache, tri, (num_vertices,))
    if hit is None:
        flat = tri.reshape(-1).long()
        # validated once per topology (this path is cached): the kernels fetch vertices through these indices
        # unchecked, where the reference's gather device-asserts (functional/face_vertices.py:16-22)
        lo, hi = int(flat.min()), int(flat.max())
        if lo < 0 or hi >= num_vertices:
            raise IndexError(f"faces index vertices {lo}..{hi} but the mesh has {num_vertices}")
        order = torch.argsort(flat, stable=True)
        counts = torch.bincount(flat, minlength=num_vertices)
        offsets = torch.zeros(num_vertices + 1, dtype=torch.int32, device=tri.device)
        offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
        hit = _cache_put(_csr_cache, tri, (offsets.contiguous(), order.to(torch.int32).contiguous()), (num_vertices,))
    return hit


class PackedGradLink:
    """Side channel between the two ends of the fused region's backward (lasr_b200.losses.fused_head ->
    render_indexed): the head's backward writes the image gradient in the packed per-pixel form the rasterizer's backward
    reads (lasr_b200_head_backward_packed / lasr_b200_render_backward_packed) and hands autograd a zero-stride
    placeholder instead of ten planar gradient planes.  Valid only when the rendered image feeds the head and nothing
    else differentiable -- the rasterizer's backward checks that it received the placeholder and raises otherwise."""

    def __init__(self):
        self.aggrs = None        # the forward's [N,2,S,S] softmax (sum, max), set by render_indexed
        self.packed = None       # [N,3,S*S,4] set by the head's backward
        self.placeholder = None

    _zero = {}                   # (device, dtype) -> 0-dim zero, made once: a placeholder costs no fill kernel per step

    def make_placeholder(self, like):
        key = (like.device, like.dtype)
        z = PackedGradLink._zero.get(key)
        if z is None:
            z = torch.zeros((), dtype=like.dtype, device=like.device)
            if not (like.is_cuda and torch.cuda.is_current_stream_capturing()):
                PackedGradLink._zero[key] = z      # (a tensor born inside a graph capture belongs to that graph's pool)
        self.placeholder = z.expand(like.shape)
        return self.placeholder

    def is_placeholder(self, grad):
        p = self.placeholder
        return (p is not None and self.packed is not None and grad.data_ptr() == p.data_ptr()
                and all(st == 0 for st in grad.stride()))


class RenderIndexedFunction(Function):

    @staticmethod
    def forward(ctx, vertices, vcolors, tri, image_size, background_color, near, far, fill_back, eps, sigma_val,
                dist_func, dist_eps, gamma_val, aggr_func_rgb, aggr_func_alpha, detached_sets=0, grad_link=None):
        if not vertices.is_cuda:
            raise _cabi.CudaTensorRequired("Rasterize module supports only cuda Tensors")
        dev = vertices.device
        N, V = vertices.shape[:2]
        S = int(image_size)
        v = vertices.detach().float().contiguous()
        c = vcolors.detach().float().contiguous()
        ctx.scal = (eps, sigma_val, DIST_FUNCS[dist_func], float(np.log(1.0 / dist_eps - 1.0)), gamma_val,
                    RGB_FUNCS[aggr_func_rgb], ALPHA_FUNCS[aggr_func_alpha], bool(fill_back))
        ctx.near, ctx.far, ctx.nf = _near_far(near, far, dev)
        ctx.S = S
        K = 1 if c.dim() == 3 else int(c.shape[0])
        if detached_sets not in (0, 1) or detached_sets >= K:
            raise ValueError("detached_sets must be 0, or 1 with at least two colour sets")
        ctx.detached = int(detached_sets)
        bg = [float(x) for x in np.asarray(background_color, dtype=np.float64).reshape(-1)]
        if len(bg) != 3 * K:
            raise ValueError("background_color needs 3 values per colour set")
        aggrs = torch.empty(N, 2, S, S, dtype=torch.float32, device=dev)
        img = torch.empty(N, 3 * K + 1, S, S, dtype=torch.float32, device=dev)
        ws = _cabi.render_workspace(N, tri.shape[0], S, dev, K)
        # background colours travel as kernel arguments: no fill pass over the image (CUDA-graph safe)
        _cabi.render_forward(v, tri, c, aggrs, img, ws, S, ctx.near, ctx.far, ctx.nf, *ctx.scal, background=bg)
        # the adjacency is looked up here, where `tri` is the caller's (cached) tensor object; the backward
        # only sees re-wrapped saved tensors
        offsets, entries = vertex_adjacency(tri, V)
        # (the package re-exports the FUNCTION soft_rasterize under the module's name: fetch the module itself)
        # ("compact" exists for the reference-signature op and for render_sets; the stacked form keeps the workspace)
        ctx.keep_ws = sys.modules[SoftRasterizeFunction.__module__].keep_mode() != "none"
        if ctx.keep_ws:
            ctx.save_for_backward(v, c, tri, img, aggrs, ws, offsets, entries)
        else:
            ctx.save_for_backward(v, c, tri, img, aggrs, offsets, entries)
        ctx.shapes = (vertices.shape, vcolors.shape)
        ctx.grad_link = grad_link
        if grad_link is not None:
            if not (K == 3 and ctx.detached == 1):
                raise ValueError("grad_link needs three colour sets with the first detached (the fused region)")
            grad_link.aggrs = aggrs
        return img

    @staticmethod
    def backward(ctx, grad):
        if ctx.keep_ws:
            v, c, tri, img, aggrs, ws, offsets, entries = ctx.saved_tensors
        else:
            v, c, tri, img, aggrs, offsets, entries = ctx.saved_tensors
            ws = _cabi.render_workspace(v.shape[0], tri.shape[0], ctx.S, v.device, 1 if c.dim() == 3 else int(c.shape[0]))
        N, V = v.shape[:2]
        gv, gc = torch.empty_like(v), torch.empty_like(c)
        link = ctx.grad_link
        if link is not None and link.packed is not None:
            if not link.is_placeholder(grad):
                raise RuntimeError("render_indexed(grad_link=...): the rendered image received gradient from something "
                                   "other than the fused head; render without grad_link for that")
            packed, link.packed, link.placeholder = link.packed, None, None
            _cabi.render_backward_packed(v, tri, c, packed, offsets, entries, gv, gc, ws, ctx.S, ctx.near, ctx.far,
                                         ctx.nf, *ctx.scal, setup_done=ctx.keep_ws, detached_sets=ctx.detached)
        else:
            _cabi.render_backward(v, tri, c, img, aggrs, grad.float().contiguous(), offsets, entries, gv, gc, ws, ctx.S,
                                  ctx.near, ctx.far, ctx.nf, *ctx.scal, setup_done=ctx.keep_ws, detached_sets=ctx.detached)
        vshape, cshape = ctx.shapes
        return (gv.reshape(vshape), gc.reshape(cshape)) + (None,) * 15


class RenderSetsFunction(Function):
    """Three colour sets given as SEPARATE [N,V,3] tensors, each read with a frame shift (lasr_b200_render_forward_sets /
    lasr_b200_render_backward_sets_packed): the fused region's form of render_indexed -- set 0 (detached) and set 1 are
    the same camera-space vertices, set 1 shifted by half the batch, so no rolled copy and no stacked tensor exist and
    the colour gradients come back per tensor.  Differentiated through a PackedGradLink only."""

    @staticmethod
    def forward(ctx, vertices, set0, set1, set2, shifts, tri, image_size, background_color, near, far, fill_back, eps,
                sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb, aggr_func_alpha, grad_link=None):
        if not vertices.is_cuda:
            raise _cabi.CudaTensorRequired("Rasterize module supports only cuda Tensors")
        dev = vertices.device
        N, V = vertices.shape[:2]
        S = int(image_size)
        v = vertices.detach().float().contiguous()
        sets = [t.detach().float().reshape(N, V, 3).contiguous() for t in (set0, set1, set2)]
        ctx.shifts = tuple(int(x) for x in shifts)
        ctx.scal = (eps, sigma_val, DIST_FUNCS[dist_func], float(np.log(1.0 / dist_eps - 1.0)), gamma_val,
                    RGB_FUNCS[aggr_func_rgb], ALPHA_FUNCS[aggr_func_alpha], bool(fill_back))
        ctx.near, ctx.far, ctx.nf = _near_far(near, far, dev)
        ctx.S = S
        bg = [float(x) for x in np.asarray(background_color, dtype=np.float64).reshape(-1)]
        if len(bg) != 9:
            raise ValueError("background_color needs 3 values per colour set")
        aggrs = torch.empty(N, 2, S, S, dtype=torch.float32, device=dev)
        img = torch.empty(N, 10, S, S, dtype=torch.float32, device=dev)
        ws = _cabi.render_workspace(N, tri.shape[0], S, dev, 3)
        ctx.keep = sys.modules[SoftRasterizeFunction.__module__].keep_mode()
        saved = _cabi.saved_buffer(N, tri.shape[0], S, dev) if ctx.keep == "compact" else None
        _cabi.render_forward_sets(v, tri, sets, ctx.shifts, aggrs, img, ws, S, ctx.near, ctx.far, ctx.nf, *ctx.scal,
                                  background=bg, saved=saved)
        offsets, entries = vertex_adjacency(tri, V)
        kept = [ws] if ctx.keep == "full" else ([saved] if ctx.keep == "compact" else [])   # compact: 80 B per face
        ctx.save_for_backward(v, tri, offsets, entries, *sets, *kept)
        ctx.shapes = (vertices.shape, set1.shape, set2.shape)
        ctx.grad_link = grad_link
        if grad_link is not None:
            grad_link.aggrs = aggrs
        return img

    @staticmethod
    def backward(ctx, grad):
        v, tri, offsets, entries, s0, s1, s2 = ctx.saved_tensors[:7]
        N, V = v.shape[:2]
        saved = None
        if ctx.keep == "full":
            ws = ctx.saved_tensors[7]
        elif ctx.keep == "compact":
            saved = ctx.saved_tensors[7]
            ws = _cabi.render_backward_scratch(N, tri.shape[0], ctx.S, v.device, 3)
        else:
            ws = _cabi.render_workspace(N, tri.shape[0], ctx.S, v.device, 3)
        link = ctx.grad_link
        if link is None or link.packed is None:
            raise RuntimeError("render_sets: the backward takes the packed gradient of lasr_b200.losses.fused_head "
                               "(pass the same PackedGradLink to both)")
        if not link.is_placeholder(grad):
            raise RuntimeError("render_sets(grad_link=...): the rendered image received gradient from something other "
                               "than the fused head")
        packed, link.packed, link.placeholder = link.packed, None, None
        gv, g1, g2 = torch.empty_like(v), torch.empty_like(s1), torch.empty_like(s2)
        _cabi.render_backward_sets_packed(v, tri, (s0, s1, s2), ctx.shifts, packed, offsets, entries, gv, (None, g1, g2),
                                          ws, ctx.S, ctx.near, ctx.far, ctx.nf, *ctx.scal, setup_done=ctx.keep == "full",
                                          detached_sets=1, saved=saved)
        vshape, s1shape, s2shape = ctx.shapes
        return (gv.reshape(vshape), None, g1.reshape(s1shape), g2.reshape(s2shape)) + (None,) * 15


def render_sets(vertices, sets, shifts, tri, image_size=256, background_color=(0,) * 9, near=1, far=100, fill_back=True,
                eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
                aggr_func_alpha="prod", grad_link=None):
    """sets: three [N,V,3] tensors (the first rendered for reading only), shifts: three ints in [0, N) -- frame n of
    the image [N,10,S,S] shows frame (n + shift) mod N of each set.  Same image as render_indexed of the stacked,
    rolled sets with detached_sets=1."""
    return RenderSetsFunction.apply(vertices, sets[0], sets[1], sets[2], tuple(shifts), tri, image_size, background_color,
                                    near, far, fill_back, eps, sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb,
                                    aggr_func_alpha, grad_link)


def render_indexed(vertices, vcolors, tri, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                   fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                   aggr_func_rgb="softmax", aggr_func_alpha="prod", detached_sets=0, grad_link=None):
    """vcolors [N,V,3] -> image [N,4,S,S] (RGBA, the reference's layout).  vcolors [K,N,V,3] with
    K = 2 or 3 renders K colour sets over ONE geometry pass -> image [N,3K+1,S,S] (set s in planes
    3s..3s+2, alpha last); background_color then holds K rgb triples.  Per set the result equals a
    separate render of that set; the vertex gradient is the sum over the sets.  detached_sets=1 declares
    set 0 an output for reading only (as if its planes were ``.detach()``ed): the backward skips it.
    grad_link (a PackedGradLink shared with ``lasr_b200.losses.fused_head``): the image gradient travels in packed
    per-pixel form between the two backward kernels instead of as ten planar planes."""
    return RenderIndexedFunction.apply(vertices, vcolors, tri, image_size, background_color, near, far, fill_back,
                                       eps, sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb, aggr_func_alpha,
                                       detached_sets, grad_link)
