"""ctypes binding of liblasr_b200.so -- the C ABI declared in include/lasr_b200.h.

This is the only place Python touches native code.  PyTorch supplies device memory
(``tensor.data_ptr()``) and the current CUDA stream; no torch type crosses the boundary.
The library must exist: there is no CPU or eager fallback, a missing library raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

from . import _build

_lib = None
_lib_lock = threading.Lock()      # lib() may be reached first from a DataLoader / stager thread

_c = ctypes
_RASTER_SCALARS = [_c.c_int, _c.c_int, _c.c_int, _c.c_int,      # N, F, S, T
                   _c.c_float, _c.c_float, _c.c_void_p,          # near, far, near_far_dev
                   _c.c_float, _c.c_float, _c.c_int, _c.c_float, _c.c_float,  # eps, sigma, dist, dist_eps, gamma
                   _c.c_int, _c.c_int, _c.c_int, _c.c_int]       # rgb, alpha, sample, double_side


EXPECTED_ABI = 8     # include/lasr_b200.h; bumped whenever an entry point's argument list changes


class NativeLibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        return _load_locked()


def _load_locked():
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if not os.path.exists(path):
            raise NativeLibraryMissing(
                f"{path} not found: build it with `python -m lasr_b200._build` "
                "(lasr_b200 has no CPU/eager fallback for its CUDA kernels)")
        L = ctypes.CDLL(path)
        L.lasr_b200_abi_version.restype = _c.c_int
        if L.lasr_b200_abi_version() != EXPECTED_ABI:
            # *.so files are build products (LASR_B200_LIB may point anywhere): a stale library with the same symbols
            # and an older argument list would be called with shifted arguments
            raise NativeLibraryMissing(
                f"{path} has ABI version {L.lasr_b200_abi_version()}, these bindings are written for {EXPECTED_ABI}: "
                "rebuild it with `python -m lasr_b200._build`")
        L.lasr_b200_last_error.restype = _c.c_char_p
        L.lasr_b200_launch_count.restype = _c.c_uint64
        L.lasr_b200_raster_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_raster_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_raster_forward.restype = _c.c_int
        L.lasr_b200_raster_forward.argtypes = [_c.c_void_p] * 5 + [_c.c_size_t] + _RASTER_SCALARS + [_c.c_void_p]
        L.lasr_b200_raster_forward_bg.restype = _c.c_int
        L.lasr_b200_raster_forward_bg.argtypes = ([_c.c_void_p] * 5 + [_c.c_size_t] + _RASTER_SCALARS
                                                  + [_c.POINTER(_c.c_float), _c.c_void_p])
        L.lasr_b200_raster_backward.restype = _c.c_int
        L.lasr_b200_raster_backward.argtypes = [_c.c_void_p] * 8 + [_c.c_size_t] + _RASTER_SCALARS + [_c.c_int, _c.c_void_p]
        for nm in ("lasr_b200_raster_saved_bytes", "lasr_b200_raster_backward_scratch_bytes"):
            getattr(L, nm).restype = _c.c_size_t
            getattr(L, nm).argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_render_backward_scratch_bytes.restype = _c.c_size_t
        L.lasr_b200_render_backward_scratch_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_raster_forward_compact.restype = _c.c_int
        L.lasr_b200_raster_forward_compact.argtypes = ([_c.c_void_p] * 5 + [_c.c_size_t, _c.c_void_p, _c.c_size_t]
                                                       + _RASTER_SCALARS + [_c.POINTER(_c.c_float), _c.c_void_p])
        L.lasr_b200_raster_backward_compact.restype = _c.c_int
        L.lasr_b200_raster_backward_compact.argtypes = ([_c.c_void_p] * 8 + [_c.c_size_t, _c.c_void_p, _c.c_size_t]
                                                        + _RASTER_SCALARS + [_c.c_void_p])
        L.lasr_b200_lbs_sum_frames.restype = _c.c_int
        L.lasr_b200_lbs_sum_frames.argtypes = [_c.c_void_p] * 6 + [_c.c_int] * 4 + [_c.c_void_p]
        L.lasr_b200_render_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_render_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int, _c.c_int]
        _render_scalars = [_c.c_int] * 4 + [_c.c_float, _c.c_float, _c.c_void_p, _c.c_float, _c.c_float, _c.c_int,
                                            _c.c_float, _c.c_float, _c.c_int, _c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_render_forward.restype = _c.c_int
        L.lasr_b200_render_forward.argtypes = ([_c.c_void_p] * 6 + [_c.c_size_t] + _render_scalars
                                               + [_c.POINTER(_c.c_float), _c.c_void_p])
        L.lasr_b200_render_backward.restype = _c.c_int
        L.lasr_b200_render_backward.argtypes = ([_c.c_void_p] * 11 + [_c.c_size_t] + _render_scalars
                                                + [_c.c_int, _c.c_int, _c.c_void_p])
        L.lasr_b200_render_backward_packed.restype = _c.c_int
        L.lasr_b200_render_backward_packed.argtypes = ([_c.c_void_p] * 9 + [_c.c_size_t] + _render_scalars
                                                       + [_c.c_int, _c.c_int, _c.c_void_p])
        L.lasr_b200_render_forward_sets.restype = _c.c_int
        L.lasr_b200_render_forward_sets.argtypes = ([_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int),
                                                     _c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_size_t] + _render_scalars
                                                    + [_c.POINTER(_c.c_float), _c.c_void_p])
        L.lasr_b200_render_backward_sets_packed.restype = _c.c_int
        L.lasr_b200_render_backward_sets_packed.argtypes = ([_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p),
                                                             _c.POINTER(_c.c_int), _c.c_void_p, _c.c_void_p, _c.c_void_p,
                                                             _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_void_p, _c.c_size_t]
                                                            + _render_scalars + [_c.c_int, _c.c_int, _c.c_void_p])
        L.lasr_b200_render_forward_sets_compact.restype = _c.c_int
        L.lasr_b200_render_forward_sets_compact.argtypes = (
            [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int), _c.c_void_p, _c.c_void_p, _c.c_void_p,
             _c.c_size_t, _c.c_void_p, _c.c_size_t] + _render_scalars + [_c.POINTER(_c.c_float), _c.c_void_p])
        L.lasr_b200_render_backward_sets_packed_compact.restype = _c.c_int
        L.lasr_b200_render_backward_sets_packed_compact.argtypes = (
            [_c.c_void_p, _c.c_void_p, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int), _c.c_void_p, _c.c_void_p, _c.c_void_p,
             _c.c_void_p, _c.POINTER(_c.c_void_p), _c.c_void_p, _c.c_size_t, _c.c_void_p, _c.c_size_t] + _render_scalars
            + [_c.c_int, _c.c_void_p])
        L.lasr_b200_meshreg_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_meshreg_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        vp, ci = _c.c_void_p, _c.c_int
        for name, args in (("laplacian_forward", [vp] * 6 + [_c.c_size_t, ci, ci, vp]),
                           ("laplacian_backward", [vp] * 5 + [ci, ci, vp]),
                           ("arap_forward", [vp] * 6 + [_c.c_size_t, ci, ci, ci, vp]),
                           ("arap_backward", [vp] * 7 + [ci, ci, ci, vp]),
                           ("flatten_forward", [vp] * 4 + [_c.c_size_t, ci, ci, ci, _c.c_float, vp]),
                           ("flatten_backward", [vp] * 7 + [_c.c_size_t, ci, ci, ci, _c.c_float, vp]),
                           ("chamfer_forward", [vp] * 6 + [ci, ci, ci, vp]),
                           ("chamfer_backward", [vp] * 8 + [ci, ci, ci, vp]),
                           ("load_textures", [vp] * 4 + [ci, ci, ci, ci, vp]),
                           ("create_texture_image", [vp] * 3 + [ci, ci, ci, ci, _c.c_float, vp]),
                           ("stage_pairs", [vp] * 12 + [_c.c_size_t, ci, ci, ci, ci, vp])):
            fn = getattr(L, "lasr_b200_" + name)
            fn.restype, fn.argtypes = ci, args
        L.lasr_b200_stage_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_stage_workspace_bytes.argtypes = [ci]
        L.lasr_b200_skin_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_skin_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_skin_weights_forward.restype = _c.c_int
        L.lasr_b200_skin_weights_forward.argtypes = [_c.c_void_p] * 5 + [_c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_skin_weights_backward.restype = _c.c_int
        L.lasr_b200_skin_weights_backward.argtypes = ([_c.c_void_p] * 10 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int,
                                                                            _c.c_void_p])
        L.lasr_b200_head_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_head_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_head_forward.restype = _c.c_int
        L.lasr_b200_head_forward.argtypes = [_c.c_void_p] * 13 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_head_forward_total.restype = _c.c_int
        L.lasr_b200_head_forward_total.argtypes = ([_c.c_void_p] * 12 + [_c.c_float] + [_c.c_void_p] * 3 +
                                                   [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p])
        L.lasr_b200_head_backward_packed_shift.restype = _c.c_int
        L.lasr_b200_head_backward_packed_shift.argtypes = ([_c.c_void_p] * 15 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int,
                                                                                  _c.c_int, _c.c_void_p])
        L.lasr_b200_head_backward.restype = _c.c_int
        L.lasr_b200_head_backward.argtypes = [_c.c_void_p] * 14 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_head_backward_packed.restype = _c.c_int
        L.lasr_b200_head_backward_packed.argtypes = [_c.c_void_p] * 15 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_losses_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_losses_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_losses_forward.restype = _c.c_int
        L.lasr_b200_losses_forward.argtypes = [_c.c_void_p] * 11 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_losses_forward_total.restype = _c.c_int
        L.lasr_b200_losses_forward_total.argtypes = ([_c.c_void_p] * 12 + [_c.c_float, _c.c_void_p, _c.c_size_t, _c.c_int,
                                                                            _c.c_int, _c.c_int, _c.c_void_p])
        L.lasr_b200_losses_backward.restype = _c.c_int
        L.lasr_b200_losses_backward.argtypes = [_c.c_void_p] * 14 + [_c.c_int, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_losses_finalize.restype = _c.c_int
        L.lasr_b200_losses_finalize.argtypes = [_c.c_void_p] * 3 + [_c.c_int, _c.c_int, _c.c_float, _c.c_void_p]
        L.lasr_b200_losses_coefficients.restype = _c.c_int
        L.lasr_b200_losses_coefficients.argtypes = [_c.c_void_p] * 4 + [_c.c_int, _c.c_int, _c.c_float, _c.c_void_p]
        L.lasr_b200_losses_coefficients_total.restype = _c.c_int
        L.lasr_b200_losses_coefficients_total.argtypes = [_c.c_void_p] * 4 + [_c.c_int, _c.c_int, _c.c_float, _c.c_void_p]
        L.lasr_b200_flow_epilogue_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_flow_epilogue_workspace_bytes.argtypes = [_c.c_int, _c.c_int]
        L.lasr_b200_flow_epilogue_forward.restype = _c.c_int
        L.lasr_b200_flow_epilogue_forward.argtypes = [_c.c_void_p] * 7 + [_c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_flow_epilogue_backward.restype = _c.c_int
        L.lasr_b200_flow_epilogue_backward.argtypes = [_c.c_void_p] * 7 + [_c.c_size_t, _c.c_int, _c.c_int, _c.c_void_p]
        L.lasr_b200_lbs_workspace_bytes.restype = _c.c_size_t
        L.lasr_b200_lbs_workspace_bytes.argtypes = [_c.c_int, _c.c_int, _c.c_int]
        L.lasr_b200_lbs_project_forward.restype = _c.c_int
        L.lasr_b200_lbs_project_forward.argtypes = [_c.c_void_p] * 10 + [_c.c_size_t] + [_c.c_int] * 6 + [_c.c_void_p]
        L.lasr_b200_lbs_pose_forward.restype = _c.c_int
        L.lasr_b200_lbs_pose_forward.argtypes = ([_c.c_void_p] * 10 + [_c.c_float, _c.c_void_p, _c.c_void_p, _c.c_size_t]
                                                 + [_c.c_int] * 7 + [_c.c_void_p])
        L.lasr_b200_lbs_pose_backward.restype = _c.c_int
        L.lasr_b200_lbs_pose_backward.argtypes = [_c.c_void_p] * 18 + [_c.c_size_t] + [_c.c_int] * 7 + [_c.c_void_p]
        L.lasr_b200_lbs_project_backward.restype = _c.c_int
        L.lasr_b200_lbs_project_backward.argtypes = [_c.c_void_p] * 15 + [_c.c_size_t] + [_c.c_int] * 6 + [_c.c_void_p]
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "lasr_b200_abi_version", "lasr_b200_last_error", "lasr_b200_launch_count",
    "lasr_b200_raster_workspace_bytes", "lasr_b200_raster_forward", "lasr_b200_raster_forward_bg",
    "lasr_b200_raster_backward",
    "lasr_b200_render_workspace_bytes", "lasr_b200_render_forward", "lasr_b200_render_backward",
    "lasr_b200_render_backward_packed", "lasr_b200_render_forward_sets", "lasr_b200_render_backward_sets_packed",
    "lasr_b200_lbs_sum_frames",
    "lasr_b200_raster_saved_bytes", "lasr_b200_raster_backward_scratch_bytes", "lasr_b200_render_backward_scratch_bytes",
    "lasr_b200_raster_forward_compact", "lasr_b200_raster_backward_compact", "lasr_b200_render_forward_sets_compact",
    "lasr_b200_render_backward_sets_packed_compact", "lasr_b200_head_backward_packed", "lasr_b200_head_backward_packed_shift",
    "lasr_b200_losses_workspace_bytes", "lasr_b200_losses_forward", "lasr_b200_losses_forward_total",
    "lasr_b200_losses_backward",
    "lasr_b200_losses_finalize", "lasr_b200_losses_coefficients", "lasr_b200_losses_coefficients_total",
    "lasr_b200_flow_epilogue_workspace_bytes", "lasr_b200_flow_epilogue_forward", "lasr_b200_flow_epilogue_backward",
    "lasr_b200_lbs_workspace_bytes", "lasr_b200_lbs_project_forward", "lasr_b200_lbs_project_backward",
    "lasr_b200_lbs_pose_forward", "lasr_b200_lbs_pose_backward",
    "lasr_b200_head_workspace_bytes", "lasr_b200_head_forward", "lasr_b200_head_forward_total", "lasr_b200_head_backward",
    "lasr_b200_skin_workspace_bytes", "lasr_b200_skin_weights_forward", "lasr_b200_skin_weights_backward",
    "lasr_b200_meshreg_workspace_bytes", "lasr_b200_laplacian_forward", "lasr_b200_laplacian_backward",
    "lasr_b200_arap_forward", "lasr_b200_arap_backward", "lasr_b200_flatten_forward", "lasr_b200_flatten_backward",
    "lasr_b200_chamfer_forward", "lasr_b200_chamfer_backward",
    "lasr_b200_load_textures", "lasr_b200_create_texture_image",
    "lasr_b200_stage_workspace_bytes", "lasr_b200_stage_pairs",
]


def launch_count() -> int:
    return int(lib().lasr_b200_launch_count())


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {lib().lasr_b200_last_error().decode()}")


def _require_cuda_f32(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise CudaTensorRequired(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


class CudaTensorRequired(TypeError, RuntimeError):
    """The reference means to raise TypeError for CPU tensors (functional/soft_rasterize.py:111-112)
    and in practice raises RuntimeError from CHECK_CUDA (soft_rasterize_cuda.cpp:54,78-82); this is
    both."""


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def raster_workspace(N: int, F: int, S: int, device) -> torch.Tensor:
    nbytes = int(lib().lasr_b200_raster_workspace_bytes(N, F, S))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def raster_forward(faces, textures, aggrs_info, soft_colors, workspace, image_size, near, far, near_far_dev, eps,
                   sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, texture_sample_type,
                   double_side, background=None, saved=None):
    """background=None: the reference's contract, soft_colors arrives pre-filled with the background colour
    (lasr_b200_raster_forward).  background=[r, g, b]: soft_colors is output only (lasr_b200_raster_forward_bg).
    saved (a `saved_buffer`, with background): compact mode -- the 80 bytes per face the backward needs go there and the
    workspace is scratch from the moment this call returns (lasr_b200_raster_forward_compact)."""
    for t, nm in ((faces, "faces"), (textures, "textures"), (aggrs_info, "aggrs_info"), (soft_colors, "soft_colors")):
        _require_cuda_f32(t, nm)
    N, F = faces.shape[:2]
    T = textures.shape[2]
    if saved is not None:
        if background is None:
            raise RuntimeError("raster_forward(saved=...): compact mode takes the background as an argument")
        bg = (_c.c_float * 3)(*[float(x) for x in background])
        with torch.cuda.device(faces.device):
            rc = lib().lasr_b200_raster_forward_compact(
                faces.data_ptr(), textures.data_ptr(), aggrs_info.data_ptr(), soft_colors.data_ptr(), saved.data_ptr(),
                saved.numel(), workspace.data_ptr(), workspace.numel(), N, F, int(image_size), T, float(near), float(far),
                near_far_dev.data_ptr() if near_far_dev is not None else None, float(eps), float(sigma_val),
                int(func_id_dist), float(dist_eps), float(gamma_val), int(func_id_rgb), int(func_id_alpha),
                int(texture_sample_type), int(bool(double_side)), bg, _stream_ptr())
        _check(rc, "lasr_b200_raster_forward_compact")
        return
    args = (faces.data_ptr(), textures.data_ptr(), aggrs_info.data_ptr(), soft_colors.data_ptr(),
            workspace.data_ptr(), workspace.numel(), N, F, int(image_size), T, float(near), float(far),
            near_far_dev.data_ptr() if near_far_dev is not None else None, float(eps), float(sigma_val),
            int(func_id_dist), float(dist_eps), float(gamma_val), int(func_id_rgb), int(func_id_alpha),
            int(texture_sample_type), int(bool(double_side)))
    with torch.cuda.device(faces.device):
        if background is None:
            rc = lib().lasr_b200_raster_forward(*args, _stream_ptr())
        else:
            bg = (_c.c_float * 3)(*[float(x) for x in background])
            rc = lib().lasr_b200_raster_forward_bg(*args, bg, _stream_ptr())
    _check(rc, "lasr_b200_raster_forward")


def raster_backward(faces, textures, soft_colors, aggrs_info, grad_faces, grad_textures, grad_soft_colors, workspace,
                    image_size, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                    func_id_rgb, func_id_alpha, texture_sample_type, double_side, setup_done=False, saved=None):
    """saved: the forward's `saved_buffer` (compact mode); `workspace` is then a `raster_backward_scratch`."""
    for t, nm in ((faces, "faces"), (textures, "textures"), (soft_colors, "soft_colors"), (aggrs_info, "aggrs_info"),
                  (grad_faces, "grad_faces"), (grad_textures, "grad_textures"),
                  (grad_soft_colors, "grad_soft_colors")):
        _require_cuda_f32(t, nm)
    N, F = faces.shape[:2]
    T = textures.shape[2]
    if saved is not None:
        with torch.cuda.device(faces.device):
            rc = lib().lasr_b200_raster_backward_compact(
                faces.data_ptr(), textures.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(),
                grad_faces.data_ptr(), grad_textures.data_ptr(), grad_soft_colors.data_ptr(), saved.data_ptr(),
                saved.numel(), workspace.data_ptr(), workspace.numel(), N, F, int(image_size), T, float(near), float(far),
                near_far_dev.data_ptr() if near_far_dev is not None else None, float(eps), float(sigma_val),
                int(func_id_dist), float(dist_eps), float(gamma_val), int(func_id_rgb), int(func_id_alpha),
                int(texture_sample_type), int(bool(double_side)), _stream_ptr())
        _check(rc, "lasr_b200_raster_backward_compact")
        return
    with torch.cuda.device(faces.device):
        rc = lib().lasr_b200_raster_backward(
            faces.data_ptr(), textures.data_ptr(), soft_colors.data_ptr(), aggrs_info.data_ptr(),
            grad_faces.data_ptr(), grad_textures.data_ptr(), grad_soft_colors.data_ptr(), workspace.data_ptr(),
            workspace.numel(), N, F, int(image_size), T, float(near), float(far),
            near_far_dev.data_ptr() if near_far_dev is not None else None, float(eps), float(sigma_val),
            int(func_id_dist), float(dist_eps), float(gamma_val), int(func_id_rgb), int(func_id_alpha),
            int(texture_sample_type), int(bool(double_side)), int(bool(setup_done)), _stream_ptr())
    _check(rc, "lasr_b200_raster_backward")


def _p(t):
    return t.data_ptr() if t is not None else None


def saved_buffer(N: int, F: int, S: int, device) -> torch.Tensor:
    """What compact mode keeps from forward to backward: 80 bytes per face + the coordinate table."""
    return torch.empty(int(lib().lasr_b200_raster_saved_bytes(N, F, int(S))), dtype=torch.uint8, device=device)


def raster_backward_scratch(N: int, F: int, S: int, device) -> torch.Tensor:
    return torch.empty(int(lib().lasr_b200_raster_backward_scratch_bytes(N, F, int(S))), dtype=torch.uint8, device=device)


def render_backward_scratch(N: int, F: int, S: int, device, num_sets: int) -> torch.Tensor:
    return torch.empty(int(lib().lasr_b200_render_backward_scratch_bytes(N, F, int(S), int(num_sets))), dtype=torch.uint8,
                       device=device)


def lbs_workspace(N: int, V: int, B: int, device) -> torch.Tensor:
    return torch.empty(int(lib().lasr_b200_lbs_workspace_bytes(N, V, B)), dtype=torch.uint8, device=device)


def lbs_project_forward(verts, Rmat, Tmat, skin, pp, fl, cam, px, near_far, workspace, N, V, B, verts_shared,
                        skin_shared, tocam):
    for t, nm in ((verts, "verts"), (Rmat, "Rmat"), (Tmat, "Tmat"), (cam, "cam")):
        _require_cuda_f32(t, nm)
    for t, nm in ((skin, "skin"), (pp, "pp"), (fl, "fl"), (px, "px"), (near_far, "near_far")):
        if t is not None:
            _require_cuda_f32(t, nm)
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_lbs_project_forward(_p(verts), _p(Rmat), _p(Tmat), _p(skin), _p(pp), _p(fl), _p(cam),
                                                 _p(px), _p(near_far), _p(workspace), workspace.numel(), N, V, B,
                                                 int(verts_shared), int(skin_shared), int(tocam), _stream_ptr())
    _check(rc, "lasr_b200_lbs_project_forward")


def lbs_project_backward(verts, Rmat, Tmat, skin, pp, fl, grad_cam, grad_px, grad_verts, grad_R, grad_T, grad_skin,
                         grad_pp, grad_fl, workspace, N, V, B, verts_shared, skin_shared, tocam):
    for t in (verts, Rmat, Tmat, grad_verts, grad_R, grad_T):
        _require_cuda_f32(t, "lbs tensor")
    for t in (skin, pp, fl, grad_cam, grad_px, grad_skin, grad_pp, grad_fl):
        if t is not None:
            _require_cuda_f32(t, "lbs tensor")
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_lbs_project_backward(_p(verts), _p(Rmat), _p(Tmat), _p(skin), _p(pp), _p(fl),
                                                  _p(grad_cam), _p(grad_px), _p(grad_verts), _p(grad_R), _p(grad_T),
                                                  _p(grad_skin), _p(grad_pp), _p(grad_fl), _p(workspace),
                                                  workspace.numel(), N, V, B, int(verts_shared), int(skin_shared),
                                                  int(tocam), _stream_ptr())
    _check(rc, "lasr_b200_lbs_project_backward")


def lbs_pose_forward(verts, Rmat, Tmat, rest_ts, skin, pp, fl, cam, px, view, eye_z, near_far, workspace, N, V, B, n_hypo,
                     verts_shared, skin_shared, tocam):
    for t, nm in ((verts, "verts"), (Rmat, "Rmat"), (Tmat, "Tmat"), (cam, "cam")):
        _require_cuda_f32(t, nm)
    for t, nm in ((rest_ts, "rest_ts"), (skin, "skin"), (pp, "pp"), (fl, "fl"), (px, "px"), (view, "view"),
                  (near_far, "near_far")):
        if t is not None:
            _require_cuda_f32(t, nm)
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_lbs_pose_forward(_p(verts), _p(Rmat), _p(Tmat), _p(rest_ts), _p(skin), _p(pp), _p(fl),
                                              _p(cam), _p(px), _p(view), float(eye_z), _p(near_far), _p(workspace),
                                              workspace.numel(), N, V, B, int(n_hypo), int(verts_shared),
                                              int(skin_shared), int(tocam), _stream_ptr())
    _check(rc, "lasr_b200_lbs_pose_forward")


def lbs_pose_backward(verts, Rmat, Tmat, rest_ts, skin, pp, fl, grad_cam, grad_px, grad_view, grad_verts, grad_R, grad_T,
                      grad_rest, grad_skin, grad_pp, grad_fl, workspace, N, V, B, n_hypo, verts_shared, skin_shared, tocam):
    for t in (verts, Rmat, Tmat, grad_verts, grad_R, grad_T):
        _require_cuda_f32(t, "lbs tensor")
    for t in (rest_ts, skin, pp, fl, grad_cam, grad_px, grad_view, grad_rest, grad_skin, grad_pp, grad_fl):
        if t is not None:
            _require_cuda_f32(t, "lbs tensor")
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_lbs_pose_backward(_p(verts), _p(Rmat), _p(Tmat), _p(rest_ts), _p(skin), _p(pp), _p(fl),
                                               _p(grad_cam), _p(grad_px), _p(grad_view), _p(grad_verts), _p(grad_R),
                                               _p(grad_T), _p(grad_rest), _p(grad_skin), _p(grad_pp), _p(grad_fl),
                                               _p(workspace), workspace.numel(), N, V, B, int(n_hypo),
                                               int(verts_shared), int(skin_shared), int(tocam), _stream_ptr())
    _check(rc, "lasr_b200_lbs_pose_backward")


def lbs_sum_frames(grad_verts, grad_skin, grad_rest, N, V, B, n_hypo):
    """Per-frame gradients of shared parameters summed over frames in one launch: -> (sum_verts [V,3] | None,
    sum_skin [B-1,V] | None, sum_rest [n_hypo,B-1,3] | None), one per non-None input."""
    ref = next(t for t in (grad_verts, grad_skin, grad_rest) if t is not None)
    for t in (grad_verts, grad_skin, grad_rest):
        if t is not None:
            _require_cuda_f32(t, "lbs gradient")
    mk = lambda t, *sh: None if t is None else torch.empty(*sh, dtype=torch.float32, device=ref.device)
    sv, ss, sr = mk(grad_verts, V, 3), mk(grad_skin, max(B - 1, 0), V), mk(grad_rest, int(n_hypo), max(B - 1, 0), 3)
    with torch.cuda.device(ref.device):
        rc = lib().lasr_b200_lbs_sum_frames(_p(grad_verts), _p(grad_skin), _p(grad_rest), _p(sv), _p(ss), _p(sr), N, V, B,
                                            int(n_hypo), _stream_ptr())
    _check(rc, "lasr_b200_lbs_sum_frames")
    return sv, ss, sr


def render_workspace(N: int, F: int, S: int, device, num_sets: int = 1) -> torch.Tensor:
    nbytes = int(lib().lasr_b200_render_workspace_bytes(N, F, S, num_sets))
    if nbytes == 0:
        raise RuntimeError("render: bad sizes (num_sets must be 1, 2 or 3)")
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                        func_id_rgb, func_id_alpha, double_side, num_sets):
    return (N, V, F, int(S), float(near), float(far), _p(near_far_dev), float(eps), float(sigma_val),
            int(func_id_dist), float(dist_eps), float(gamma_val), int(func_id_rgb), int(func_id_alpha),
            int(bool(double_side)), int(num_sets))


def _num_sets(verts, vcolors):
    """vcolors [N,V,3] -> 1; [K,N,V,3] -> K."""
    if vcolors.dim() == verts.dim():
        return 1
    if vcolors.dim() != verts.dim() + 1 or tuple(vcolors.shape[1:]) != tuple(verts.shape):
        raise RuntimeError("vcolors must be [N,V,3] or [K,N,V,3]")
    return int(vcolors.shape[0])


def render_forward(verts, tri, vcolors, aggrs_info, soft_colors, workspace, S, near, far, near_far_dev, eps,
                   sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side,
                   background=None):
    """background: 3K floats (then soft_colors need not be pre-filled) or None (pre-filled image)."""
    for t, nm in ((verts, "verts"), (vcolors, "vcolors"), (aggrs_info, "aggrs_info"), (soft_colors, "soft_colors")):
        _require_cuda_f32(t, nm)
    if tri.dtype != torch.int32 or not tri.is_cuda or not tri.is_contiguous():
        raise RuntimeError("tri must be a contiguous CUDA int32 tensor [F,3]")
    N, V = verts.shape[:2]
    F = tri.shape[0]
    K = _num_sets(verts, vcolors)
    if tuple(soft_colors.shape[:2]) != (N, 3 * K + 1):
        raise RuntimeError("soft_colors must be [N,3K+1,S,S]")
    bg_arr = None
    if background is not None:
        if len(background) != 3 * K:
            raise RuntimeError("background needs 3 floats per colour set")
        bg_arr = (_c.c_float * (3 * K))(*[float(x) for x in background])
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_render_forward(
            _p(verts), _p(tri), _p(vcolors), _p(aggrs_info), _p(soft_colors), _p(workspace), workspace.numel(),
            *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                 gamma_val, func_id_rgb, func_id_alpha, double_side, K), bg_arr, _stream_ptr())
    _check(rc, "lasr_b200_render_forward")


def render_backward(verts, tri, vcolors, soft_colors, aggrs_info, grad_soft_colors, csr_offsets, csr_entries,
                    grad_verts, grad_vcolors, workspace, S, near, far, near_far_dev, eps, sigma_val, func_id_dist,
                    dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side, setup_done=False, detached_sets=0):
    for t in (verts, vcolors, soft_colors, aggrs_info, grad_soft_colors, grad_verts, grad_vcolors):
        _require_cuda_f32(t, "render tensor")
    N, V = verts.shape[:2]
    F = tri.shape[0]
    K = _num_sets(verts, vcolors)
    if grad_vcolors.shape != vcolors.shape or grad_soft_colors.shape != soft_colors.shape:
        raise RuntimeError("render_backward: gradient shapes must match vcolors / soft_colors")
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_render_backward(
            _p(verts), _p(tri), _p(vcolors), _p(soft_colors), _p(aggrs_info), _p(grad_soft_colors), _p(csr_offsets),
            _p(csr_entries), _p(grad_verts), _p(grad_vcolors), _p(workspace), workspace.numel(),
            *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                 gamma_val, func_id_rgb, func_id_alpha, double_side, K), int(detached_sets),
            int(bool(setup_done)), _stream_ptr())
    _check(rc, "lasr_b200_render_backward")


def _set_arrays(sets, shifts, N, V, what):
    K = len(sets)
    if K < 1 or K > 3 or len(shifts) != K:
        raise RuntimeError(f"{what}: one to three colour sets, one frame shift each")
    for t, sh in zip(sets, shifts):
        if t is not None:
            _require_cuda_f32(t, "colour set")
            if tuple(t.shape) != (N, V, 3):
                raise RuntimeError(f"{what}: every colour set is [N,V,3]")
        if not 0 <= int(sh) < N:
            raise RuntimeError(f"{what}: frame shifts lie in [0, N)")
    return (_c.c_void_p * K)(*[_p(t) for t in sets]), (_c.c_int * K)(*[int(x) for x in shifts])


def render_forward_sets(verts, tri, sets, shifts, aggrs_info, soft_colors, workspace, S, near, far, near_far_dev, eps,
                        sigma_val, func_id_dist, dist_eps, gamma_val, func_id_rgb, func_id_alpha, double_side,
                        background, saved=None):
    """lasr_b200_render_forward_sets: `sets` K tensors [N,V,3], frame n of the image reads frame (n + shifts[s]) mod N.
    saved (a `saved_buffer`): compact mode, the workspace is scratch once the call returns."""
    for t, nm in ((verts, "verts"), (aggrs_info, "aggrs_info"), (soft_colors, "soft_colors")):
        _require_cuda_f32(t, nm)
    if tri.dtype != torch.int32 or not tri.is_cuda or not tri.is_contiguous():
        raise RuntimeError("tri must be a contiguous CUDA int32 tensor [F,3]")
    N, V = verts.shape[:2]
    F, K = tri.shape[0], len(sets)
    if any(t is None for t in sets):
        raise RuntimeError("render_forward_sets: null colour set")
    ptrs, sh = _set_arrays(sets, shifts, N, V, "render_forward_sets")
    if tuple(soft_colors.shape[:2]) != (N, 3 * K + 1) or len(background) != 3 * K:
        raise RuntimeError("soft_colors must be [N,3K+1,S,S] and background 3 floats per colour set")
    bg_arr = (_c.c_float * (3 * K))(*[float(x) for x in background])
    with torch.cuda.device(verts.device):
        if saved is not None:
            rc = lib().lasr_b200_render_forward_sets_compact(
                _p(verts), _p(tri), ptrs, sh, _p(aggrs_info), _p(soft_colors), _p(saved), saved.numel(), _p(workspace),
                workspace.numel(),
                *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                     gamma_val, func_id_rgb, func_id_alpha, double_side, K), bg_arr, _stream_ptr())
        else:
            rc = lib().lasr_b200_render_forward_sets(
                _p(verts), _p(tri), ptrs, sh, _p(aggrs_info), _p(soft_colors), _p(workspace), workspace.numel(),
                *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                     gamma_val, func_id_rgb, func_id_alpha, double_side, K), bg_arr, _stream_ptr())
    _check(rc, "lasr_b200_render_forward_sets")


def render_backward_sets_packed(verts, tri, sets, shifts, packed_grad, csr_offsets, csr_entries, grad_verts, grad_sets,
                                workspace, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                                func_id_rgb, func_id_alpha, double_side, setup_done=False, detached_sets=1, saved=None):
    """grad_sets: K tensors [N,V,3] or None (detached sets), written in each set's own frame order.
    saved: the forward's `saved_buffer` (compact mode); `workspace` is then a `render_backward_scratch`."""
    for t in (verts, packed_grad, grad_verts):
        _require_cuda_f32(t, "render tensor")
    N, V = verts.shape[:2]
    F, K = tri.shape[0], len(sets)
    ptrs, sh = _set_arrays(sets, shifts, N, V, "render_backward_sets_packed")
    gptrs, _ = _set_arrays(grad_sets, shifts, N, V, "render_backward_sets_packed")
    if packed_grad.numel() != N * 3 * int(S) * int(S) * 4:
        raise RuntimeError("render_backward_sets_packed: packed_grad is [N,3,S*S,4]")
    if saved is not None:
        with torch.cuda.device(verts.device):
            rc = lib().lasr_b200_render_backward_sets_packed_compact(
                _p(verts), _p(tri), ptrs, sh, _p(packed_grad), _p(csr_offsets), _p(csr_entries), _p(grad_verts), gptrs,
                _p(saved), saved.numel(), _p(workspace), workspace.numel(),
                *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                     gamma_val, func_id_rgb, func_id_alpha, double_side, K), int(detached_sets),
                _stream_ptr())
        _check(rc, "lasr_b200_render_backward_sets_packed_compact")
        return
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_render_backward_sets_packed(
            _p(verts), _p(tri), ptrs, sh, _p(packed_grad), _p(csr_offsets), _p(csr_entries), _p(grad_verts), gptrs,
            _p(workspace), workspace.numel(),
            *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                 gamma_val, func_id_rgb, func_id_alpha, double_side, K), int(detached_sets),
            int(bool(setup_done)), _stream_ptr())
    _check(rc, "lasr_b200_render_backward_sets_packed")


def render_backward_packed(verts, tri, vcolors, packed_grad, csr_offsets, csr_entries, grad_verts, grad_vcolors,
                           workspace, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps, gamma_val,
                           func_id_rgb, func_id_alpha, double_side, setup_done=False, detached_sets=1):
    for t in (verts, vcolors, packed_grad, grad_verts, grad_vcolors):
        _require_cuda_f32(t, "render tensor")
    N, V = verts.shape[:2]
    F = tri.shape[0]
    K = _num_sets(verts, vcolors)
    if grad_vcolors.shape != vcolors.shape or packed_grad.numel() != N * 3 * int(S) * int(S) * 4:
        raise RuntimeError("render_backward_packed: gradient shapes must match vcolors / [N,3,S*S,4]")
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_render_backward_packed(
            _p(verts), _p(tri), _p(vcolors), _p(packed_grad), _p(csr_offsets), _p(csr_entries), _p(grad_verts),
            _p(grad_vcolors), _p(workspace), workspace.numel(),
            *_render_scalar_args(N, V, F, S, near, far, near_far_dev, eps, sigma_val, func_id_dist, dist_eps,
                                 gamma_val, func_id_rgb, func_id_alpha, double_side, K), int(detached_sets),
            int(bool(setup_done)), _stream_ptr())
    _check(rc, "lasr_b200_render_backward_packed")


def losses_forward(mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, sums, N, H, S, sub=None, out3=None,
                   l1tex_wt=0.0):
    """sums [N,H,7]; with `sub` [3,N,H] and `out3` -- FOUR entries: the three means and their sum -- also the per-pair
    losses, their means and the objective, in the same launches."""
    if out3 is not None and out3.numel() != 4:
        raise RuntimeError("losses_forward: out3 takes four entries (mask, flow, texture means and their sum)")
    ws = torch.empty(int(lib().lasr_b200_losses_workspace_bytes(N, H, S)), dtype=torch.uint8, device=mask_pred.device)
    with torch.cuda.device(mask_pred.device):
        rc = lib().lasr_b200_losses_forward_total(_p(mask_pred), _p(flow_rd), _p(bgmask), _p(tex), _p(fg), _p(imgs),
                                                  _p(masks), _p(occ), _p(flow_obs), _p(sums), _p(sub), _p(out3),
                                                  float(l1tex_wt), _p(ws), ws.numel(), N, H, S, _stream_ptr())
    _check(rc, "lasr_b200_losses_forward_total")


def losses_backward(mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, grad_sums, g_mask_pred,
                    g_flow_rd, g_tex, g_fg, N, H, S):
    with torch.cuda.device(mask_pred.device):
        rc = lib().lasr_b200_losses_backward(_p(mask_pred), _p(flow_rd), _p(bgmask), _p(tex), _p(fg), _p(imgs),
                                             _p(masks), _p(occ), _p(flow_obs), _p(grad_sums), _p(g_mask_pred),
                                             _p(g_flow_rd), _p(g_tex), _p(g_fg), N, H, S, _stream_ptr())
    _check(rc, "lasr_b200_losses_backward")


def losses_finalize(sums, sub, out3, N, H, l1tex_wt):
    with torch.cuda.device(sums.device):
        rc = lib().lasr_b200_losses_finalize(_p(sums), _p(sub), _p(out3), N, H, float(l1tex_wt), _stream_ptr())
    _check(rc, "lasr_b200_losses_finalize")


def losses_coefficients(sums, g3, gsub, gsums, N, H, l1tex_wt):
    """g3 [3] -> gsums; g3 with FOUR entries (the output of the *_total forward calls: three means and their sum) takes
    the gradient of the sum from the last one."""
    if g3.numel() not in (3, 4):
        raise RuntimeError("losses_coefficients: the upstream gradient has 3 entries (or 4 with the sum)")
    fn = lib().lasr_b200_losses_coefficients_total if g3.numel() == 4 else lib().lasr_b200_losses_coefficients
    with torch.cuda.device(sums.device):
        rc = fn(_p(sums), _p(g3), _p(gsub), _p(gsums), N, H, float(l1tex_wt), _stream_ptr())
    _check(rc, "lasr_b200_losses_coefficients")


def meshreg_workspace(N, V, E, device):
    return torch.empty(int(lib().lasr_b200_meshreg_workspace_bytes(N, V, E)), dtype=torch.uint8, device=device)


def native_call(name, *args):
    return meshreg_call(name, *args)


def meshreg_call(name, *args):
    """lasr_b200_<name>(*args, stream): tensors become device pointers, numbers pass through."""
    dev = next(a.device for a in args if torch.is_tensor(a))
    conv = []
    for a in args:
        if torch.is_tensor(a):
            if not a.is_cuda or not a.is_contiguous():
                raise CudaTensorRequired(f"lasr_b200_{name}: tensors must be contiguous CUDA tensors")
            if a.dtype not in (torch.float32, torch.int32, torch.uint8):
                raise RuntimeError(f"lasr_b200_{name}: unsupported dtype {a.dtype}")
            conv.append(_p(a))
        else:
            conv.append(a)
    with torch.cuda.device(dev):
        rc = getattr(lib(), "lasr_b200_" + name)(*conv, _stream_ptr())
    _check(rc, "lasr_b200_" + name)


def skin_weights_forward(ctl_ts, ctl_rs, log_ctl, verts, skin, H, Bp, V):
    for t in (ctl_ts, ctl_rs, log_ctl, verts, skin):
        _require_cuda_f32(t, "skin tensor")
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_skin_weights_forward(_p(ctl_ts), _p(ctl_rs), _p(log_ctl), _p(verts), _p(skin), H, Bp, V,
                                                  _stream_ptr())
    _check(rc, "lasr_b200_skin_weights_forward")


def skin_weights_backward(ctl_ts, ctl_rs, log_ctl, verts, skin, grad_skin, g_ts, g_rs, g_log, H, Bp, V):
    for t in (ctl_ts, ctl_rs, log_ctl, verts, skin, grad_skin, g_ts, g_rs, g_log):
        _require_cuda_f32(t, "skin tensor")
    ws = torch.empty(int(lib().lasr_b200_skin_workspace_bytes(H, Bp, V)), dtype=torch.uint8, device=verts.device)
    with torch.cuda.device(verts.device):
        rc = lib().lasr_b200_skin_weights_backward(_p(ctl_ts), _p(ctl_rs), _p(log_ctl), _p(verts), _p(skin),
                                                   _p(grad_skin), _p(g_ts), _p(g_rs), _p(g_log), _p(ws), ws.numel(),
                                                   H, Bp, V, _stream_ptr())
    _check(rc, "lasr_b200_skin_weights_backward")


def head_forward(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, sums, flow_rd, bgmask, N, H, S,
                 sub=None, out3=None, l1tex_wt=0.0, pair_shift=0):
    for t in (image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, sums):
        _require_cuda_f32(t, "head tensor")
    if tuple(image.shape) != (N * H, 10, S, S):
        raise RuntimeError("head: image must be [N*H,10,S,S] (three colour sets + silhouette)")
    if out3 is not None and out3.numel() != 4:
        raise RuntimeError("head_forward: out3 takes four entries (mask, flow, texture means and their sum)")
    ws = torch.empty(int(lib().lasr_b200_head_workspace_bytes(N, H, S)), dtype=torch.uint8, device=image.device)
    with torch.cuda.device(image.device):
        rc = lib().lasr_b200_head_forward_total(_p(image), _p(pp_own), _p(pp_pair), _p(f_own), _p(f_pair), _p(imgs),
                                                _p(masks), _p(occ), _p(flow_obs), _p(sums), _p(sub), _p(out3),
                                                float(l1tex_wt), _p(flow_rd), _p(bgmask), _p(ws), ws.numel(), N, H, S,
                                                int(pair_shift), _stream_ptr())
    _check(rc, "lasr_b200_head_forward_total")


def head_backward(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, grad_sums, grad_image,
                  grad_pp_pair, grad_f_pair, N, H, S):
    for t in (image, grad_sums, grad_image, grad_pp_pair, grad_f_pair):
        _require_cuda_f32(t, "head tensor")
    ws = torch.empty(int(lib().lasr_b200_head_workspace_bytes(N, H, S)), dtype=torch.uint8, device=image.device)
    with torch.cuda.device(image.device):
        rc = lib().lasr_b200_head_backward(_p(image), _p(pp_own), _p(pp_pair), _p(f_own), _p(f_pair), _p(imgs),
                                           _p(masks), _p(occ), _p(flow_obs), _p(grad_sums), _p(grad_image),
                                           _p(grad_pp_pair), _p(grad_f_pair), _p(ws), ws.numel(), N, H, S,
                                           _stream_ptr())
    _check(rc, "lasr_b200_head_backward")


def head_backward_packed(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, grad_sums, aggrs_info,
                         packed_grad, grad_pp_pair, grad_f_pair, N, H, S, pair_shift=0):
    for t in (image, grad_sums, aggrs_info, packed_grad, grad_pp_pair, grad_f_pair):
        _require_cuda_f32(t, "head tensor")
    ws = torch.empty(int(lib().lasr_b200_head_workspace_bytes(N, H, S)), dtype=torch.uint8, device=image.device)
    with torch.cuda.device(image.device):
        rc = lib().lasr_b200_head_backward_packed_shift(
            _p(image), _p(pp_own), _p(pp_pair), _p(f_own), _p(f_pair), _p(imgs), _p(masks), _p(occ), _p(flow_obs),
            _p(grad_sums), _p(aggrs_info), _p(packed_grad), _p(grad_pp_pair), _p(grad_f_pair), _p(ws), ws.numel(), N, H, S,
            int(pair_shift), _stream_ptr())
    _check(rc, "lasr_b200_head_backward_packed_shift")


def flow_epilogue_forward(px, pp0, pp1, f0, f1, flow, bgmask, n, S):
    with torch.cuda.device(px.device):
        rc = lib().lasr_b200_flow_epilogue_forward(_p(px), _p(pp0), _p(pp1), _p(f0), _p(f1), _p(flow), _p(bgmask), n, S,
                                                   _stream_ptr())
    _check(rc, "lasr_b200_flow_epilogue_forward")


def flow_epilogue_backward(px, f1, grad_flow, grad_px, grad_pp1, grad_f1, n, S):
    ws = torch.empty(int(lib().lasr_b200_flow_epilogue_workspace_bytes(n, S)), dtype=torch.uint8, device=px.device)
    with torch.cuda.device(px.device):
        rc = lib().lasr_b200_flow_epilogue_backward(_p(px), _p(f1), _p(grad_flow), _p(grad_px), _p(grad_pp1), _p(grad_f1),
                                                    _p(ws), ws.numel(), n, S, _stream_ptr())
    _check(rc, "lasr_b200_flow_epilogue_backward")
