"""oracle/ref_gpu.py -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Driver for the UNMODIFIED reference SoftRas CUDA extension built into ``oracle/_ref`` from the
sources where they lie under /root/reference (``make -C oracle ref``; host-glue shim
``oracle/ref_shim/ref_compat.h`` is force-included, no reference file is copied or edited).

The reference's Python wrapper (functional/soft_rasterize.py) cannot travel to the GPU box, so the
buffer set-up it performs around the pybind calls is restated here (allocation, background
pre-fill, zero-fill, enum maps, ``dist_eps -> log(1/eps-1)``: soft_rasterize.py:22-53,88-100).
Used (a) to mint and check golden vectors and (b) as the ``--impl reference`` arm of bench.py.
"""
from __future__ import annotations

import importlib.util
import glob
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ext = None

DIST = {"hard": 0, "barycentric": 1, "euclidean": 2}
RGB = {"hard": 0, "softmax": 1}
ALPHA = {"hard": 0, "sum": 1, "prod": 2}
SAMPLE = {"surface": 0, "vertex": 1}


def available() -> bool:
    return bool(glob.glob(os.path.join(_HERE, "_ref", "soft_rasterize*.so")))


def ext():
    global _ext
    if _ext is None:
        cands = glob.glob(os.path.join(_HERE, "_ref", "soft_rasterize*.so"))
        if not cands:
            raise RuntimeError("oracle/_ref/soft_rasterize*.so missing: run `make -C oracle ref` where "
                               "/root/reference exists")
        spec = importlib.util.spec_from_file_location("soft_rasterize", cands[0])
        _ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ext)
    return _ext


class RefSoftRasterize(torch.autograd.Function):
    """The reference's SoftRasterizeFunction, calling the reference's own compiled kernels."""

    @staticmethod
    def forward(ctx, face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface"):
        ctx.args = (int(image_size), float(near), float(far), float(eps), float(sigma_val), DIST[dist_func],
                    float(np.log(1.0 / dist_eps - 1.0)), float(gamma_val), RGB[aggr_func_rgb],
                    ALPHA[aggr_func_alpha], SAMPLE[texture_type], bool(fill_back))
        fv = face_vertices.clone()
        tex = textures.clone()
        dev = fv.device
        nb, nf = fv.shape[:2]
        S = int(image_size)
        faces_info = torch.zeros(nb, nf, 27, dtype=torch.float32, device=dev)
        aggrs_info = torch.zeros(nb, 2, S, S, dtype=torch.float32, device=dev)
        soft_colors = torch.ones(nb, 4, S, S, dtype=torch.float32, device=dev)
        for k in range(3):
            soft_colors[:, k] *= background_color[k]
        faces_info, aggrs_info, soft_colors = ext().forward_soft_rasterize(fv, tex, faces_info, aggrs_info,
                                                                           soft_colors, *ctx.args)
        ctx.save_for_backward(fv, tex, soft_colors, faces_info, aggrs_info)
        ctx.aggrs_info = aggrs_info
        return soft_colors

    @staticmethod
    def backward(ctx, grad_soft_colors):
        fv, tex, soft_colors, faces_info, aggrs_info = ctx.saved_tensors
        grad_faces = torch.zeros_like(fv)
        grad_textures = torch.zeros_like(tex)
        g = grad_soft_colors.contiguous()
        grad_faces, grad_textures = ext().backward_soft_rasterize(fv, tex, soft_colors, faces_info, aggrs_info,
                                                                  grad_faces, grad_textures, g, *ctx.args)
        return (grad_faces, grad_textures) + (None,) * 13


def soft_rasterize(face_vertices, textures, **settings):
    order = ["image_size", "background_color", "near", "far", "fill_back", "eps", "sigma_val", "dist_func",
             "dist_eps", "gamma_val", "aggr_func_rgb", "aggr_func_alpha", "texture_type"]
    defaults = dict(image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True, eps=1e-3,
                    sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
                    aggr_func_alpha="prod", texture_type="surface")
    defaults.update(settings)
    return RefSoftRasterize.apply(face_vertices, textures, *[defaults[k] for k in order])
