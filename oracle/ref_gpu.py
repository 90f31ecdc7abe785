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


# ---- the reference's render_mesh glue, restated (it cannot travel to the GPU box) -------------
def _vertex_normals(vertices, faces):
    """functional/vertex_normals.py:4-37."""
    import torch.nn.functional as F
    bs, nv = vertices.shape[:2]
    normals = torch.zeros(bs * nv, 3, device=vertices.device)
    faces = faces + (torch.arange(bs, dtype=torch.int32, device=vertices.device) * nv)[:, None, None]
    vf = vertices.reshape(bs * nv, 3)[faces.long()].view(-1, 3, 3)
    faces = faces.view(-1, 3)
    normals.index_add_(0, faces[:, 1].long(), torch.cross(vf[:, 2] - vf[:, 1], vf[:, 0] - vf[:, 1], dim=1))
    normals.index_add_(0, faces[:, 2].long(), torch.cross(vf[:, 0] - vf[:, 2], vf[:, 1] - vf[:, 2], dim=1))
    normals.index_add_(0, faces[:, 0].long(), torch.cross(vf[:, 1] - vf[:, 0], vf[:, 2] - vf[:, 0], dim=1))
    return F.normalize(normals, eps=1e-6, dim=1).reshape(bs, nv, 3)


def _face_gather(x, faces):
    """functional/face_vertices.py:4-22."""
    bs, nv = x.shape[:2]
    faces = faces + (torch.arange(bs, dtype=torch.int32, device=x.device) * nv)[:, None, None]
    return x.reshape(bs * nv, 3)[faces.long()]


def render_mesh_lasr(vertices, faces, textures, eye, settings):
    """SoftRenderer.render_mesh with LASR's renderer settings (nnutils/mesh_net.py:132-149):
    'vertex' lighting with ambient 1 / directional 0 (lighting.py:50-67 -- the vertex normals ARE
    computed and multiplied by 0, as the reference does), look_at + orthogonal(1)
    (transform.py:41-48), two gathers (mesh.py:105-109,128-134), then the reference kernels."""
    import torch.nn.functional as F
    light = torch.zeros_like(vertices)
    light = light + 1.0 * torch.ones(1, 1, 3, device=vertices.device)
    normals = _vertex_normals(vertices, faces)
    direction = torch.tensor([0.0, 1.0, 0.0], device=vertices.device)[None]
    cosine = F.relu(torch.sum(normals * direction, dim=2))
    light = light + 0.0 * (torch.ones(1, 1, 3, device=vertices.device) * cosine[:, :, None])
    textures = textures * light
    # look_at(eye) with at = 0, up = +y (look_at.py:44-60)
    bs = vertices.shape[0]
    e = eye[None].repeat(bs, 1)
    z_axis = F.normalize(-e, eps=1e-5)
    up = torch.tensor([0.0, 1.0, 0.0], device=vertices.device)[None].repeat(bs, 1)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.cat((x_axis[:, None], y_axis[:, None], z_axis[:, None]), dim=1)
    v = torch.matmul(vertices - e[:, None], r.transpose(1, 2))
    v = torch.stack((v[:, :, 0] * 1.0, v[:, :, 1] * 1.0, v[:, :, 2]), dim=2)  # orthogonal(scale=1)
    return soft_rasterize(_face_gather(v, faces), _face_gather(textures, faces), **settings)


# ---- the whole hot-path region of LASR.forward, reference style -------------------------------
def _ref_obj_to_cam(verts, Rmat, Tmat, nmesh, skin):
    """nnutils/geom_utils.py:45-71 as the reference executes it: one matmul per bone, cat, weighted sum."""
    verts, Rmat, Tmat = verts.clone(), Rmat.clone(), Tmat.clone()
    bodyR, bodyT = Rmat[::nmesh].clone(), Tmat[::nmesh].clone()
    if nmesh > 1:
        parts = []
        for k in range(nmesh - 1):
            parts.append((verts.matmul(Rmat[k + 1::nmesh].clone()) + Tmat[k + 1::nmesh].clone())[:, None])
        vs = (torch.cat(parts, 1) * skin).sum(1)
    else:
        vs = verts
    return vs.clone().matmul(bodyR) + bodyT


def _ref_pinhole(verts, pp, fl):
    """nnutils/geom_utils.py:27-34 (clones + in-place slice assigns)."""
    n_hypo = verts.shape[0] // pp.shape[0]
    pp = pp.clone()[:, None].repeat(1, n_hypo, 1).view(-1, 2)
    fl = fl.clone()[:, None].view(-1, 1)
    verts = verts.clone()
    verts[:, :, 1] = pp[:, 1:2] + verts[:, :, 1].clone() * fl / verts[:, :, 2].clone()
    verts[:, :, 0] = pp[:, 0:1] + verts[:, :, 0].clone() * fl / verts[:, :, 2].clone()
    return verts


def _ref_render_flow(verts, faces, pos0, pos1, pp0, pp1, f0, f1, eye, settings):
    """nnutils/mesh_net.py:75-104 with the reference kernels."""
    n_hypo = verts.shape[0] // faces.shape[0]
    faces = faces[:, None].repeat(1, n_hypo, 1, 1).view(-1, faces.shape[1], 3)
    pos0, pos1 = pos0.clone(), pos1.clone()
    pre = verts[:, :, :3] + eye[None, None]
    pre[:, :, 1] = -1 * pre[:, :, 1]
    nb = verts.shape[0]
    px = render_mesh_lasr(torch.cat([pre, pre], 0), torch.cat([faces, faces], 0),
                          torch.cat([pos0[:, :, :3], pos1[:, :, :3]], 0), eye, settings).clone()
    fg = px[:nb, -1]
    px = px[:, :3]
    p0 = px[:nb].permute(0, 2, 3, 1)
    p1 = px[nb:].permute(0, 2, 3, 1)
    bg = (p0[:, :, :, 2] < 1e-9) | (p1[:, :, :, 2] < 1e-9)
    p0[bg] = 10
    p1[bg] = 10
    p0[:, :, :, 1] = pp0[:, 1:2, None] + p0[:, :, :, 1].clone() * f0[:, :1, None] / p0[:, :, :, 2].clone()
    p0[:, :, :, 0] = pp0[:, 0:1, None] + p0[:, :, :, 0].clone() * f0[:, :1, None] / p0[:, :, :, 2].clone()
    p1[:, :, :, 1] = pp1[:, 1:2, None] + p1[:, :, :, 1].clone() * f1[:, :1, None] / p1[:, :, :, 2].clone()
    p1[:, :, :, 0] = pp1[:, 0:1, None] + p1[:, :, :, 0].clone() * f1[:, :1, None] / p1[:, :, :, 2].clone()
    flow = (p1 - p0.detach())[:, :, :, :2]
    flow[bg] = flow[bg].detach()
    return flow, bg, fg


def lasr_region_reference(pred_v, tex, faces, Rmat, Tmat, ppoint, scale, batch, n_hypo, n_bones, S, skin, eye,
                          sigval=1e-4, l1tex_wt=1.0):
    """Restatement of nnutils/mesh_net.py:298-447 (LBS -> projection -> near/far with the host sync ->
    flow fw / bw renders -> texture render -> mask / flow / texture-L1 losses in Python double loops)
    on the reference's own kernels.  Skin weights and the rest-pose bone transforms are taken as inputs.
    Bench / comparison infrastructure only."""
    H, B = n_hypo, n_bones
    n2 = ppoint.shape[0]
    bs = n2 // 2
    vfl = _ref_obj_to_cam(pred_v, Rmat, Tmat[:, None, :], B, skin)
    vfl = torch.cat([vfl, torch.ones_like(vfl[:, :, 0:1])], dim=-1)
    pos0 = vfl.view(n2, H, -1, 4)[:bs].clone().view(bs * H, -1, 4)
    pos1 = vfl.view(n2, H, -1, 4)[bs:].clone().view(bs * H, -1, 4)
    vpx = _ref_pinhole(vfl, ppoint, scale)
    dmax, dmin = vpx[:, :, -2].max(), vpx[:, :, -2].min()
    near, far = float(dmin - (dmax - dmin) / 2), float(dmax + (dmax - dmin) / 2)   # pybind `float` coercion = sync
    st = dict(image_size=S, background_color=[0, 0, 0], near=near, far=far, fill_back=True, eps=1e-3, sigma_val=sigval,
              dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax", aggr_func_alpha="prod",
              texture_type="vertex")
    v4 = vpx.view(n2, H, -1, 4)
    rp = lambda t, lo, hi: t[:, None][lo:hi].repeat(1, H, 1).view(-1, 2)
    sc = scale.view(n2, H, 1)
    ffw, bgf, _ = _ref_render_flow(v4[:bs].reshape(bs * H, -1, 4), faces[:bs], pos0, pos1, rp(ppoint, 0, bs),
                                   rp(ppoint, bs, n2), sc[:bs].reshape(-1, 1), sc[bs:].reshape(-1, 1), eye, st)
    fbw, bgb, _ = _ref_render_flow(v4[bs:].reshape(bs * H, -1, 4), faces[bs:], pos1, pos0, rp(ppoint, bs, n2),
                                   rp(ppoint, 0, bs), sc[bs:].reshape(-1, 1), sc[:bs].reshape(-1, 1), eye, st)
    bgmask = torch.cat([bgf, bgb], 0)
    flow_rd = torch.cat([ffw, fbw], 0)
    pre = vpx[:, :, :3] + eye[None, None]
    pre[:, :, 1] = -1 * pre[:, :, 1]
    rend = render_mesh_lasr(pre, faces[:, None].repeat(1, H, 1, 1).view(-1, faces.shape[1], 3), tex, eye,
                            dict(st, background_color=[1, 1, 1])).clone()
    mask_pred_all, rgb = rend[:, -1], rend[:, :3]
    masks, occ, imgs, flow = batch["masks"], batch["occ"], batch["imgs"], batch["flow"]
    img_obs = imgs * (masks > 0).float()[:, None]
    img_rnd = rgb * mask_pred_all[:, None]
    img_white = 1 - (masks > 0).float()[:, None] + img_obs
    # 1) mask loss (:374-390)
    mask_pred = mask_pred_all.view(n2, -1, S, S)
    msub = (mask_pred - masks[:, None]).pow(2)
    tmp = torch.zeros(n2, H, device=pred_v.device)
    for i in range(n2):
        for j in range(H):
            tmp[i, j] = msub[i, j][occ[i] != 0].mean()
    mask_loss = (0.5 * tmp).mean()
    # 2) flow loss (:393-416)
    frd = flow_rd.view(n2, -1, S, S, 2)
    mask = (~bgmask).view(n2, -1, S, S) & ((occ != 0)[:, None] & (masks > 0)[:, None]).repeat(1, H, 1, 1)
    fmap = torch.norm((frd - flow[:, None, :2].permute(0, 1, 3, 4, 2)), 2, -1)
    w = (-occ).sigmoid()[:, None].repeat(1, H, 1, 1)
    for i in range(n2):
        w[i] = w[i] / w[i][mask[i]].mean()
    fmap = fmap * w
    tmp = torch.zeros(n2, H, device=pred_v.device)
    for i in range(n2):
        for j in range(H):
            tmp[i, j] = fmap[i, j][mask[i, j]].mean()
            if mask[i, j].sum() == 0:
                tmp[i, j] = 0
    flow_loss = (0.5 * tmp).mean()
    # 3) texture L1 (:420-441, LPIPS term excluded)
    tmp = torch.zeros(n2, H, device=pred_v.device)
    for i in range(n2):
        for j in range(H):
            tmp[i, j] += (img_obs[i] - img_rnd.view(n2, -1, 3, S, S)[i, j]).abs().mean(0)[occ[i] != 0].mean()
            tmp[i, j] += (img_white[i] - rgb.view(n2, -1, 3, S, S)[i, j]).abs().mean(0)[occ[i] != 0].mean()
            tmp[i, j] *= 2 * l1tex_wt
    tex_loss = (0.25 * tmp).mean()
    return mask_loss + flow_loss + tex_loss, mask_loss, flow_loss, tex_loss


# ---- the part of the region upstream of obj_to_cam, reference style (mesh_net.py:259-289) ------------------
def _ref_quat_to_mat(q):
    """kornia 0.5.3 quaternion_to_rotation_matrix, XYZW (not vendored by the reference; restated)."""
    q = torch.nn.functional.normalize(q, p=2, dim=-1, eps=1e-12)
    x, y, z, w = q.unbind(-1)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    one = torch.ones_like(x)
    m = torch.stack([one - (ty * y + tz * z), ty * x - tz * w, tz * x + ty * w,
                     ty * x + tz * w, one - (tx * x + tz * z), tz * y - tx * w,
                     tz * x - ty * w, tz * y + tx * w, one - (tx * x + ty * y)], dim=-1)
    return m.view(q.shape[:-1] + (3, 3))


def ref_skin_and_bones(pred_v, Rmat, Tmat, bones, n2, n_hypo, n_bones):
    """nnutils/mesh_net.py:261-289 as the reference executes it: the softmax skinning weights in ~6 PyTorch kernels,
    repeated per frame, and the part transforms composed about their rest positions."""
    H, B = n_hypo, n_bones
    dis = bones["ctl_ts"].view(H, -1, 1, 3) - pred_v.view(n2, H, -1, 3)[0].detach().view(H, 1, -1, 3)
    dis = dis.matmul(_ref_quat_to_mat(bones["ctl_rs"]).view(H, -1, 3, 3))
    dis = bones["log_ctl"].exp().view(H, -1, 1, 3) * dis.pow(2)
    skin = (-10 * dis.sum(3)).softmax(1)[:, :, :, None].repeat(n2, 1, 1, 1)
    rest = bones["rest_ts"][:, None, :, None].repeat(n2, 1, 1, 1).view(-1, B - 1, 3, 1)
    R4, T4 = Rmat.view(-1, B, 3, 3), Tmat.view(-1, B, 3, 1)
    Tm = torch.cat([T4[:, :1], -R4[:, 1:].matmul(rest) + T4[:, 1:] + rest], 1).view(-1, 3)
    Rm = torch.cat([R4[:, :1], R4[:, 1:].permute(0, 1, 3, 2)], 1).view(-1, 3, 3)
    return skin, Rm, Tm


LASR_EYE = (0.0, 0.0, -2.7320508075688776)     # SoftRenderer's default look_at eye: -(1 / tan(30 deg) + 1) (renderer.py:60-63)
