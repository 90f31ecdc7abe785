"""``SoftRasterizeFunction`` / ``soft_rasterize`` with the reference's signatures
(third_party/softras/soft_renderer/functional/soft_rasterize.py:9-119), backed by the sm_100a
kernels in lasr_b200/csrc through the C ABI (include/lasr_b200.h).

Differences a caller can observe, all deliberate:
  * ``near`` / ``far`` may be python numbers or 0-dim CUDA tensors; tensors are handed to the
    kernel as a device pointer, so LASR's per-step depth window (nnutils/mesh_net.py:304-311)
    costs no host synchronisation (the reference's pybind ``float`` coercion does);
  * the scratch tensor the reference calls ``faces_info`` [N,F,27] is replaced by a workspace of packed
    per-face windows, border boxes and records, which IS saved for backward (``setup_done=True``), exactly as
    the reference hands ``faces_info`` from its forward to its backward (soft_rasterize.py:64,71);
  * backward is deterministic (no float atomics).
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.autograd import Function

from ... import _cabi

# Keep the forward's per-face scratch (200 B per face: windows, boxes, records) alive for the backward, as the reference
# keeps its `faces_info` (108 B per face, soft_rasterize.py:64,71).  False: nothing but the op's own inputs and outputs is
# saved and the backward recomputes the records (`setup_done=0`): 0 B per face kept, a few per cent slower backward
# (profiles/r2/README.md).  "compact": 80 B per face kept -- the record's division results in a separate buffer, the
# workspace itself freed when the forward returns; the backward re-derives the rest from the op's inputs, same bits
# (lasr_b200_raster_forward_compact / _backward_compact; the colour-set route of the fused region likewise).
# Read at forward time; LASR_B200_KEEP_WORKSPACE=0 / =compact in the environment set the default.
_env_keep = os.environ.get("LASR_B200_KEEP_WORKSPACE", "1")
KEEP_WORKSPACE = "compact" if _env_keep == "compact" else _env_keep != "0"


def keep_mode():
    """-> "full" | "compact" | "none" from the module switch (read at forward time)."""
    k = KEEP_WORKSPACE
    if isinstance(k, str):
        if k != "compact":
            raise ValueError("KEEP_WORKSPACE is True, False or 'compact'")
        return "compact"
    return "full" if k else "none"

DIST_FUNCS = {"hard": 0, "barycentric": 1, "euclidean": 2}
RGB_FUNCS = {"hard": 0, "softmax": 1}
ALPHA_FUNCS = {"hard": 0, "sum": 1, "prod": 2}
SAMPLE_TYPES = {"surface": 0, "vertex": 1}


def _near_far(near, far, device):
    """-> (near_float, far_float, device_tensor_or_None)."""
    if torch.is_tensor(near) or torch.is_tensor(far):
        nf = torch.stack([torch.as_tensor(near, dtype=torch.float32, device=device).reshape(()),
                          torch.as_tensor(far, dtype=torch.float32, device=device).reshape(())]).detach()
        return 0.0, 0.0, nf.contiguous()
    return float(near), float(far), None


class SoftRasterizeFunction(Function):

    @staticmethod
    def forward(ctx, face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface"):
        # face_vertices [nb, nf, 3, 3] (or [nb, nf, 9]); textures [nb, nf, T, 3]
        if not face_vertices.is_cuda:
            raise _cabi.CudaTensorRequired("Rasterize module supports only cuda Tensors")
        ctx.image_size = int(image_size)
        ctx.eps = eps
        ctx.sigma_val = sigma_val
        ctx.gamma_val = gamma_val
        ctx.func_dist_type = DIST_FUNCS[dist_func]
        ctx.dist_eps = float(np.log(1.0 / dist_eps - 1.0))
        ctx.func_rgb_type = RGB_FUNCS[aggr_func_rgb]
        ctx.func_alpha_type = ALPHA_FUNCS[aggr_func_alpha]
        ctx.texture_type = SAMPLE_TYPES[texture_type]
        ctx.fill_back = bool(fill_back)

        device = face_vertices.device
        nb, nf = face_vertices.shape[:2]
        fv = face_vertices.detach().to(torch.float32).reshape(nb, nf, 9).contiguous()
        tex = textures.detach().to(torch.float32).reshape(nb, nf, -1, 3).contiguous()
        ctx.in_shapes = (face_vertices.shape, textures.shape)
        ctx.near, ctx.far, ctx.near_far_dev = _near_far(near, far, device)

        S = ctx.image_size
        aggrs_info = torch.empty(nb, 2, S, S, dtype=torch.float32, device=device)
        soft_colors = torch.empty(nb, 4, S, S, dtype=torch.float32, device=device)
        # the background colour travels as a kernel argument (lasr_b200_raster_forward_bg): the reference's three fills
        # of soft_colors (:50-53) and the kernel's read of them are 24 S^2 bytes of HBM traffic per frame for a constant
        workspace = _cabi.raster_workspace(nb, nf, S, device)
        ctx.keep = keep_mode()
        saved = _cabi.saved_buffer(nb, nf, S, device) if ctx.keep == "compact" else None
        _cabi.raster_forward(fv, tex, aggrs_info, soft_colors, workspace, S, ctx.near, ctx.far, ctx.near_far_dev,
                             eps, sigma_val, ctx.func_dist_type, ctx.dist_eps, gamma_val, ctx.func_rgb_type,
                             ctx.func_alpha_type, ctx.texture_type, ctx.fill_back,
                             background=[float(background_color[k]) for k in range(3)], saved=saved)
        if ctx.keep == "full":
            ctx.save_for_backward(fv, tex, soft_colors, aggrs_info, workspace)   # workspace plays the reference's faces_info
        elif ctx.keep == "compact":
            ctx.save_for_backward(fv, tex, soft_colors, aggrs_info, saved)       # 80 B per face; the workspace dies here
        else:
            ctx.save_for_backward(fv, tex, soft_colors, aggrs_info)
        return soft_colors

    @staticmethod
    def backward(ctx, grad_soft_colors):
        saved = None
        if ctx.keep == "full":
            fv, tex, soft_colors, aggrs_info, workspace = ctx.saved_tensors
        elif ctx.keep == "compact":
            fv, tex, soft_colors, aggrs_info, saved = ctx.saved_tensors
            workspace = _cabi.raster_backward_scratch(fv.shape[0], fv.shape[1], ctx.image_size, fv.device)
        else:
            fv, tex, soft_colors, aggrs_info = ctx.saved_tensors
            workspace = _cabi.raster_workspace(fv.shape[0], fv.shape[1], ctx.image_size, fv.device)
        nb, nf = fv.shape[:2]
        S = ctx.image_size
        grad_faces = torch.empty_like(fv)
        grad_textures = torch.empty_like(tex)
        g = grad_soft_colors.to(torch.float32).contiguous()
        _cabi.raster_backward(fv, tex, soft_colors, aggrs_info, grad_faces, grad_textures, g, workspace, S,
                              ctx.near, ctx.far, ctx.near_far_dev, ctx.eps, ctx.sigma_val, ctx.func_dist_type,
                              ctx.dist_eps, ctx.gamma_val, ctx.func_rgb_type, ctx.func_alpha_type, ctx.texture_type,
                              ctx.fill_back, setup_done=ctx.keep == "full", saved=saved)
        fshape, tshape = ctx.in_shapes
        return (grad_faces.reshape(fshape), grad_textures.reshape(tshape)) + (None,) * 13


def soft_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                   fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                   aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface"):
    if not face_vertices.is_cuda:
        raise _cabi.CudaTensorRequired("Rasterize module supports only cuda Tensors")
    return SoftRasterizeFunction.apply(face_vertices, textures, image_size, background_color, near, far, fill_back,
                                       eps, sigma_val, dist_func, dist_eps, gamma_val, aggr_func_rgb,
                                       aggr_func_alpha, texture_type)
