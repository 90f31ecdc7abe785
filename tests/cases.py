"""Seeded input cases shared by the parity tests and by the golden-vector scripts.

Each case is a dict: face_vertices [N,F,3,3] f32, textures [N,F,T,3] f32, settings (the 13
non-tensor arguments of SoftRasterizeFunction.forward), grad_soft_colors [N,4,S,S] f32.
Inputs are synthetic: posed Spot / icosphere / bumpy-sphere meshes under
the render_syn.py camera sweep with LASR's training rasterizer settings, plus small cases that
exercise every dist / rgb / alpha / texture mode of the operator."""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lasr_b200 import synth  # noqa: E402


def _lasr_settings(S, near, far, sigma=1e-4, bg=(0, 0, 0)):
    return dict(image_size=S, background_color=list(bg), near=near, far=far, fill_back=True, eps=1e-3,
                sigma_val=sigma, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax",
                aggr_func_alpha="prod", texture_type="vertex")


def _mesh(name):
    if name == "spot":
        return synth.load_spot()
    if name.startswith("ico"):
        return synth.icosphere(int(name[3:]), 0.5)
    if name.startswith("bumpy"):
        return synth.bumpy_sphere(int(name[5:]), 0.5)
    raise KeyError(name)


def lasr_case(mesh="ico2", N=2, S=64, sigma=1e-4, bg=(0, 0, 0), seed=0, grad_seed=1):
    v, f = _mesh(mesh)
    px, cam, near, far = synth.render_inputs(v, f, N, seed)
    fv = synth.gather_faces(px, f)
    tex = synth.gather_faces(np.repeat(synth.vertex_colors(len(v), seed)[None], N, 0), f)
    g = (np.random.RandomState(grad_seed).randn(N, 4, S, S) * 1e-3).astype(np.float32)
    return dict(face_vertices=fv.astype(np.float32), textures=tex.astype(np.float32),
                settings=_lasr_settings(S, near, far, sigma, bg), grad_soft_colors=g)


def mode_case(dist_func, aggr_func_rgb, aggr_func_alpha, texture_type, fill_back=True, mesh="bumpy1", N=2, S=48,
              texture_res=2, sigma=1e-4, seed=0):
    """Small case exercising one combination of operator modes."""
    c = lasr_case(mesh, N, S, sigma, (0.2, 0.4, 0.6), seed)
    s = c["settings"]
    s.update(dist_func=dist_func, aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha,
             texture_type=texture_type, fill_back=fill_back)
    if dist_func == "barycentric":
        s["sigma_val"] = 1e-2      # barycentric distances are O(1), not O(pixel^2)
    if aggr_func_rgb == "hard":
        s["gamma_val"] = 1e-4
    if texture_type == "surface":
        F = c["face_vertices"].shape[1]
        rng = np.random.RandomState(seed + 7)
        c["textures"] = rng.rand(N, F, texture_res * texture_res, 3).astype(np.float32)
        # The reference's 'surface' texel formula (soft_rasterize_cuda_kernel.cu:182-188) indexes one
        # texel block too far when a clipped w[0] == 1; for the LAST face that read is out of bounds
        # (undefined).  Park the last face off screen so the fixture is well defined.
        c["face_vertices"][:, -1, :, 0] += 10.0
    return c


# name -> zero-argument constructor.  Small ones are stored as golden fixtures.
GOLDEN_CASES = {
    "lasr_ico2_s64": lambda: lasr_case("ico2", 2, 64),
    "lasr_ico2_s64_sig1e-5_white": lambda: lasr_case("ico2", 2, 64, sigma=1e-5, bg=(1, 1, 1)),
    "lasr_bumpy2_s96": lambda: lasr_case("bumpy2", 2, 96),
    "lasr_spot_s128": lambda: lasr_case("spot", 1, 128),
    "mode_hard_hard_hard_surface": lambda: mode_case("hard", "hard", "hard", "surface"),
    "mode_bary_softmax_sum_surface": lambda: mode_case("barycentric", "softmax", "sum", "surface"),
    "mode_euclid_hard_prod_vertex": lambda: mode_case("euclidean", "hard", "prod", "vertex"),
    "mode_euclid_softmax_prod_surface_r1": lambda: mode_case("euclidean", "softmax", "prod", "surface",
                                                           texture_res=1),
    "mode_euclid_softmax_prod_vertex_nofill": lambda: mode_case("euclidean", "softmax", "prod", "vertex",
                                                               fill_back=False),
    "mode_bary_hard_sum_vertex": lambda: mode_case("barycentric", "hard", "sum", "vertex"),
}

# Larger cases compared live on the GPU only (not stored).
LIVE_CASES = {
    "lasr_spot_s256_n2": lambda: lasr_case("spot", 2, 256),
    "lasr_spot_s256_n2_sig1e-5": lambda: lasr_case("spot", 2, 256, sigma=1e-5),
    "lasr_ico4_s256_n2": lambda: lasr_case("ico4", 2, 256),
    "lasr_ico5_s512_n1": lambda: lasr_case("ico5", 1, 512),          # BASELINE config D shape (20480 faces, 512^2)
    "lasr_ico3_s256_n2_white": lambda: lasr_case("ico3", 2, 256, bg=(1, 1, 1)),   # the shape LASR trains at
}
