"""Geometry glue of the LASR render hot path with the reference's signatures
(nnutils/geom_utils.py:27-95, nnutils/mesh_net.py:75-104, 264-267, 304-311).

``obj_to_cam`` (linear-blend skinning), ``pinhole_cam`` and the batch depth window are served by
hand-written sm_100a kernels (lasr_b200/csrc/lbs.cu, through the C ABI) whenever the tensors live
on a CUDA device; there is no silent fallback on a GPU box -- a missing library raises.  CPU
tensors take the plain PyTorch expressions below, which is what the reference itself executes for
CPU tensors (these functions are device-agnostic PyTorch there) and what the host-side tests pin
against golden vectors minted from the reference's own Python.
"""
from __future__ import annotations

import numpy as np
import torch

from . import soft_renderer as sr


# ----------------------------------------------------------------------------------------------
# S0: skinning weights (nnutils/mesh_net.py:264-267)
# ----------------------------------------------------------------------------------------------
def quaternion_to_rotation_matrix(quaternion: torch.Tensor) -> torch.Tensor:
    """kornia==0.5.3 ``quaternion_to_rotation_matrix`` (XYZW order; identity is [0,0,0,1],
    nnutils/mesh_net.py:118,120).  kornia is not vendored by the reference: restated from its
    published algorithm (normalise, then the unit-quaternion matrix).  Parity unpinned."""
    q = torch.nn.functional.normalize(quaternion, p=2, dim=-1, eps=1e-12)
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = torch.ones_like(x)
    m = torch.stack([one - (tyy + tzz), txy - twz, txz + twy,
                     txy + twz, one - (txx + tzz), tyz - twx,
                     txz - twy, tyz + twx, one - (txx + tyy)], dim=-1)
    return m.view(quaternion.shape[:-1] + (3, 3))


def skin_weights(ctl_ts, ctl_rs, log_ctl, mean_v, n_hypo):
    """Gaussian-mixture skinning weights, softmax over bones (nnutils/mesh_net.py:264-267).

    ctl_ts [H*(B-1),3], ctl_rs [H*(B-1),4] (XYZW), log_ctl [H*(B-1),3], mean_v [H,V,3] (detached by
    the caller, as the reference does) -> skin [H,B-1,V,1]."""
    H = n_hypo
    if _use_native(mean_v):
        from . import native_geom
        return native_geom.SkinWeightsFunction.apply(ctl_ts, ctl_rs, log_ctl, mean_v, H)
    dis = ctl_ts.view(H, -1, 1, 3) - mean_v.view(H, 1, -1, 3)
    dis = dis.matmul(quaternion_to_rotation_matrix(ctl_rs).view(H, -1, 3, 3))
    dis = log_ctl.exp().view(H, -1, 1, 3) * dis.pow(2)
    return (-10 * dis.sum(3)).softmax(1)[:, :, :, None]


# ----------------------------------------------------------------------------------------------
# G1 / G2: LBS and camera projection (nnutils/geom_utils.py:27-71)
# ----------------------------------------------------------------------------------------------
def _use_native(*tensors):
    return any(torch.is_tensor(t) and t.is_cuda for t in tensors)


def pinhole_cam(verts, pp, fl):
    """x,y <- pp + xy * f / z; remaining channels kept (geom_utils.py:27-34).  Out of place (the
    reference assigns into slices of a clone; values are identical)."""
    if _use_native(verts):
        from . import native_geom
        return native_geom.pinhole_cam(verts, pp, fl)
    n_hypo = verts.shape[0] // pp.shape[0]
    pp = pp[:, None].repeat(1, n_hypo, 1).view(-1, 2)
    fl = fl[:, None].reshape(-1, 1)
    z = verts[:, :, 2]
    x = pp[:, 0:1] + verts[:, :, 0] * fl / z
    y = pp[:, 1:2] + verts[:, :, 1] * fl / z
    return torch.cat([x[..., None], y[..., None], verts[:, :, 2:]], dim=-1)


def orthographic_cam(verts, pp, fl):
    """geom_utils.py:36-43."""
    n_hypo = verts.shape[0] // pp.shape[0]
    pp = pp[:, None].repeat(1, n_hypo, 1).view(-1, 2)
    fl = fl[:, None].reshape(-1, 1)
    x = pp[:, 0:1] + verts[:, :, 0] * fl
    y = pp[:, 1:2] + verts[:, :, 1] * fl
    return torch.cat([x[..., None], y[..., None], verts[:, :, 2:]], dim=-1)


def obj_to_cam(verts, Rmat, Tmat, nmesh, n_hypo, skin, tocam=True):
    """Canonical object coordinates -> camera coordinates with linear-blend skinning
    (geom_utils.py:45-71), row-vector convention:
        v' = (sum_k w_k (v R_k + T_k)) R_0 + T_0
    verts [N,V,3]; Rmat [N*nmesh,3,3]; Tmat [N*nmesh,1,3]; skin [N,nmesh-1,V,1] (or None)."""
    verts = verts.reshape(-1, verts.shape[1], 3)
    if _use_native(verts):
        from . import native_geom
        return native_geom.obj_to_cam(verts, Rmat, Tmat, nmesh, skin, tocam)
    Rmat = Rmat.reshape(-1, 3, 3)
    Tmat = Tmat.reshape(-1, 1, 3)
    if nmesh > 1:
        N = verts.shape[0]
        R = Rmat.view(N, nmesh, 3, 3)
        T = Tmat.view(N, nmesh, 1, 3)
        parts = verts[:, None].matmul(R[:, 1:]) + T[:, 1:]       # [N, K, V, 3]
        vs = (parts * skin).sum(1)
    else:
        vs = verts
    if tocam:
        vs = vs.matmul(Rmat[::nmesh]) + Tmat[::nmesh]
    return vs


def depth_window(verts_z):
    """near/far of the batch (nnutils/mesh_net.py:304-311): near = dmin - (dmax-dmin)/2,
    far = dmax + (dmax-dmin)/2.  Returned as 0-dim tensors on the input's device; the rasterizer
    takes them as a device pointer, so no host synchronisation happens."""
    dmax = verts_z.max()
    dmin = verts_z.min()
    half = (dmax - dmin) / 2
    return dmin - half, dmax + half


# ----------------------------------------------------------------------------------------------
# R1: flow rendering (nnutils/mesh_net.py:75-104, nnutils/geom_utils.py:73-95)
# ----------------------------------------------------------------------------------------------
def _eye_offset(renderer_soft, device):
    eye = renderer_soft.transform.transformer._eye
    if torch.is_tensor(eye):
        return eye.to(device=device, dtype=torch.float32).reshape(1, 1, 3)
    from .soft_renderer.functional.geometry import _const      # uploaded once per (value, device): graph-capturable
    return _const(eye, device)[None, None]


def _reproject(pos_px, pp, fl):
    z = pos_px[..., 2]
    x = pp[:, 0:1, None] + pos_px[..., 0] * fl[:, :1, None] / z
    y = pp[:, 1:2, None] + pos_px[..., 1] * fl[:, :1, None] / z
    return torch.stack([x, y], dim=-1)


class _FlowEpilogue(torch.autograd.Function):
    """mesh_net.py:88-104 on CUDA tensors in one kernel (lasr_b200_flow_epilogue_forward/backward)."""

    @staticmethod
    def forward(ctx, px, pp0, pp1, f0, f1):
        from . import _cabi
        n2, _, S, _ = px.shape
        n = n2 // 2
        pxc = px.detach().float().contiguous()
        c = lambda t, sh: t.detach().float().reshape(sh).contiguous()
        pp0c, pp1c, f0c, f1c = c(pp0, (n, 2)), c(pp1, (n, 2)), c(f0, (n,)), c(f1, (n,))
        flow = torch.empty(n, S, S, 2, dtype=torch.float32, device=px.device)
        bg = torch.empty(n, S, S, dtype=torch.uint8, device=px.device)
        _cabi.flow_epilogue_forward(pxc, pp0c, pp1c, f0c, f1c, flow, bg, n, S)
        ctx.save_for_backward(pxc, f1c)
        ctx.shapes = (pp1.shape, f1.shape)
        bgb = bg.bool()
        ctx.mark_non_differentiable(bgb)
        return flow, bgb

    @staticmethod
    def backward(ctx, gflow, _gbg):
        from . import _cabi
        pxc, f1c = ctx.saved_tensors
        n2, _, S, _ = pxc.shape
        n = n2 // 2
        gpx = torch.empty_like(pxc)
        gpp1 = torch.empty(n, 2, dtype=torch.float32, device=pxc.device)
        gf1 = torch.empty(n, dtype=torch.float32, device=pxc.device)
        _cabi.flow_epilogue_backward(pxc, f1c, gflow.float().contiguous(), gpx, gpp1, gf1, n, S)
        return gpx, None, gpp1.reshape(ctx.shapes[0]), None, gf1.reshape(ctx.shapes[1])


def render_flow_soft_2(renderer_soft, verts, faces, verts_pos0, verts_pos1, pp0, pp1, proj_cam0, proj_cam1,
                       shared_faces=None):
    """Render the 3-D positions of two frames as vertex 'colours' through the soft rasterizer and
    re-project them with each frame's intrinsics to get optical flow (mesh_net.py:75-104).
    Returns (flow_fw [n,S,S,2], bgmask [n,S,S] bool, fgmask [n,S,S])."""
    n_hypo = verts.shape[0] // faces.shape[0]
    if shared_faces is None and faces.is_cuda:
        # the one-topology check is cached per tensor OBJECT: ask it of the caller's tensor, not of the fresh
        # repeat / cat copies built below (those would re-run the synchronising comparison at every call)
        shared_faces = sr.functional.shared_topology(faces)
    faces = faces[:, None].repeat(1, n_hypo, 1, 1).view(-1, faces.shape[1], 3)
    offset = _eye_offset(renderer_soft, verts.device)
    pre = verts[:, :, :3] + offset
    pre = torch.cat([pre[:, :, :1], -pre[:, :, 1:2], pre[:, :, 2:]], dim=-1)
    nb = verts.shape[0]
    px = renderer_soft.render_mesh(sr.Mesh(torch.cat([pre, pre], 0), torch.cat([faces, faces], 0),
                                           textures=torch.cat([verts_pos0[:, :, :3], verts_pos1[:, :, :3]], 0),
                                           texture_type="vertex", shared_faces=shared_faces))
    fgmask = px[:nb, -1]
    if px.is_cuda:
        flow_fw, bgmask = _FlowEpilogue.apply(px, pp0, pp1, proj_cam0[:, :1], proj_cam1[:, :1])
        return flow_fw, bgmask, fgmask
    px = px[:, :3]
    pos0 = px[:nb].permute(0, 2, 3, 1)
    pos1 = px[nb:].permute(0, 2, 3, 1)
    bgmask = (pos0[..., 2] < 1e-9) | (pos1[..., 2] < 1e-9)
    ten = torch.full_like(pos0, 10.0)
    pos0 = torch.where(bgmask[..., None], ten, pos0)
    pos1 = torch.where(bgmask[..., None], ten, pos1)
    p0 = _reproject(pos0, pp0, proj_cam0)
    p1 = _reproject(pos1, pp1, proj_cam1)
    flow_fw = p1 - p0.detach()
    flow_fw = torch.where(bgmask[..., None], flow_fw.detach(), flow_fw)
    return flow_fw, bgmask, fgmask


def render_flow_soft_3(renderer_soft, verts, verts_target, faces):
    """geom_utils.py:73-95 -- flow from rendered target positions against the pixel grid."""
    offset = _eye_offset(renderer_soft, verts.device)
    pre = verts[:, :, :3] + offset
    pre = torch.cat([pre[:, :, :1], -pre[:, :, 1:2], pre[:, :, 2:]], dim=-1)
    px = renderer_soft.render_mesh(sr.Mesh(pre, faces, textures=verts_target[:, :, :3],
                                           texture_type="vertex")).clone()
    px = px.permute(0, 2, 3, 1)
    bgmask = px[..., 2] < 1e-9
    px = torch.where(bgmask[..., None], torch.full_like(px, 10.0), px)
    # the reference's fgmask is a view of the render it then overwrites with 10 on the background (all four channels,
    # geom_utils.py:82-86), so its foreground mask reads 10 there; kept
    fgmask = px[..., 3]
    H, W = bgmask.shape[1], bgmask.shape[2]
    gx = torch.arange(W, device=verts.device, dtype=torch.float32) * 2 / (W - 1) - 1
    gy = torch.arange(H, device=verts.device, dtype=torch.float32) * 2 / (H - 1) - 1
    grid = torch.stack(torch.meshgrid(gx, gy, indexing="xy"), dim=-1)[None]
    flow_fw = px[..., :2] - grid
    flow_fw = torch.where(bgmask[..., None], flow_fw.detach(), flow_fw)
    return flow_fw, bgmask, fgmask


def label_colormap():
    """The reference ships a 150-row segmentation colour table here (geom_utils.py:97-254); it only
    colours the bone visualisation and is outside the hot path."""
    raise NotImplementedError("label_colormap is visualisation-only and outside the render hot path")
