"""CUDA-tensor implementations of the geometry glue, backed by lasr_b200/csrc/lbs.cu through the C ABI.

``skin_project`` is the fused entry point the LASR step wants: one kernel skins the vertices with
B bones, applies the per-frame camera, projects, and reduces the batch depth window; its backward is
one kernel plus a tiny fixed-order reduction.  ``obj_to_cam`` / ``pinhole_cam`` are the reference's
two signatures mapped onto the same kernels."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _cabi


def _batch_shared(t, N):
    """A [N, ...] tensor that is an expand()/repeat-free view of one row is passed with stride 0."""
    return t.shape[0] == 1 and N > 1


class SkinProjectFunction(Function):
    """(verts, Rmat, Tmat, skin, pp, fl) -> (cam [N,V,3], px [N,V,4] or None, near_far [2] or None)."""

    @staticmethod
    def forward(ctx, verts, Rmat, Tmat, skin, pp, fl, nmesh, tocam, want_depth_window, rest_ts=None, n_hypo=1,
                view_eye_z=None):
        N = Rmat.shape[0] // nmesh
        V = verts.shape[1]
        dev = verts.device
        verts_c = verts.detach().float().contiguous()
        R = Rmat.detach().float().reshape(N * nmesh, 3, 3).contiguous()
        T = Tmat.detach().float().reshape(N * nmesh, 3).contiguous()
        sk = None
        if nmesh > 1:
            sk = skin.detach().float().reshape(skin.shape[0], nmesh - 1, V).contiguous()
        project = pp is not None
        ppc = pp.detach().float().reshape(N, 2).contiguous() if project else None
        flc = fl.detach().float().reshape(N).contiguous() if project else None
        cam = torch.empty(N, V, 3, dtype=torch.float32, device=dev)
        px = torch.empty(N, V, 4, dtype=torch.float32, device=dev) if project else None
        nf = torch.empty(2, dtype=torch.float32, device=dev) if want_depth_window else None
        ws = _cabi.lbs_workspace(N, V, nmesh, dev)
        vs, ss = _batch_shared(verts_c, N), (sk is not None and _batch_shared(sk, N))
        rest = view = None
        if rest_ts is not None or view_eye_z is not None:
            # raw part transforms composed about their rest positions inside the kernel, and / or the render-space vertices
            # (x, -y, (z + eye_z) - eye_z) as a third output
            if rest_ts is not None:
                rest = rest_ts.detach().float().reshape(int(n_hypo) * (nmesh - 1), 3).contiguous()
            if view_eye_z is not None:
                view = torch.empty(N, V, 3, dtype=torch.float32, device=dev)
            _cabi.lbs_pose_forward(verts_c, R, T, rest, sk, ppc, flc, cam, px, view, view_eye_z or 0.0, nf, ws, N, V, nmesh,
                                   n_hypo, vs, ss, tocam)
        else:
            _cabi.lbs_project_forward(verts_c, R, T, sk, ppc, flc, cam, px, nf, ws, N, V, nmesh, vs, ss, tocam)
        ctx.save_for_backward(verts_c, R, T, sk, ppc, flc, rest)
        ctx.has_view = view is not None
        ctx.meta = (N, V, nmesh, vs, ss, bool(tocam), verts.shape, Rmat.shape, Tmat.shape,
                    None if skin is None else skin.shape, None if pp is None else pp.shape,
                    None if fl is None else fl.shape, int(n_hypo), None if rest_ts is None else rest_ts.shape)
        outs = [cam, px if project else cam.new_empty(0), nf if want_depth_window else cam.new_empty(0),
                view if view is not None else cam.new_empty(0)]
        ctx.mark_non_differentiable(outs[2])
        ctx.set_materialize_grads(False)         # an output nobody differentiates (px on the fused route) costs no zero-fill
        if not project:
            ctx.mark_non_differentiable(outs[1])
        if view is None:
            ctx.mark_non_differentiable(outs[3])
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_cam, g_px, _g_nf, g_view):
        verts, R, T, sk, pp, fl, rest = ctx.saved_tensors
        N, V, B, vs, ss, tocam, vshape, rshape, tshape, sshape, ppshape, flshape, n_hypo, restshape = ctx.meta
        dev = verts.device
        project = pp is not None
        g_cam = g_cam.float().contiguous() if g_cam is not None else None
        g_px = g_px.float().contiguous() if (project and g_px is not None and g_px.numel()) else None
        g_view = g_view.float().contiguous() if (ctx.has_view and g_view is not None and g_view.numel()) else None
        if g_cam is None and g_px is None and g_view is None:
            return (None,) * 12
        gv = torch.empty(N, V, 3, dtype=torch.float32, device=dev)
        gR = torch.empty(N * B, 3, 3, dtype=torch.float32, device=dev)
        gT = torch.empty(N * B, 3, dtype=torch.float32, device=dev)
        gs = torch.empty(N, B - 1, V, dtype=torch.float32, device=dev) if B > 1 else None
        gpp = torch.empty(N, 2, dtype=torch.float32, device=dev) if project else None
        gfl = torch.empty(N, dtype=torch.float32, device=dev) if project else None
        ws = _cabi.lbs_workspace(N, V, B, dev)
        g_rest = g_rest_f = None
        if rest is not None or g_view is not None:
            g_rest_f = torch.empty(N, B - 1, 3, dtype=torch.float32, device=dev) if rest is not None else None
            _cabi.lbs_pose_backward(verts, R, T, rest, sk, pp, fl, g_cam, g_px, g_view, gv, gR, gT, g_rest_f, gs, gpp, gfl,
                                    ws, N, V, B, n_hypo, vs, ss, tocam)
        else:
            _cabi.lbs_project_backward(verts, R, T, sk, pp, fl, g_cam, g_px, gv, gR, gT, gs, gpp, gfl, ws, N, V, B, vs,
                                       ss, tocam)
        # gradients of parameters every frame shares: summed over frames (and g_rest over the frames of a hypothesis) in
        # ONE launch -- they were three PyTorch reductions per step
        g_rest_f = g_rest_f if rest is not None else None
        want_v, want_s = bool(vs), bool(gs is not None and ss)
        if want_v or want_s or g_rest_f is not None:
            sv, ssum, srest = _cabi.lbs_sum_frames(gv if want_v else None, gs if want_s else None, g_rest_f, N, V, B, n_hypo)
            if want_v:
                gv = sv[None]
            if want_s:
                gs = ssum[None]
            if srest is not None:
                g_rest = srest.reshape(restshape)
        return (gv.reshape(vshape), gR.reshape(rshape), gT.reshape(tshape),
                gs.reshape(sshape) if gs is not None else None,
                gpp.reshape(ppshape) if project else None, gfl.reshape(flshape) if project else None,
                None, None, None, g_rest, None, None)


class SkinWeightsFunction(Function):
    """nnutils/mesh_net.py:264-267 on CUDA tensors (lasr_b200_skin_weights_forward/backward).
    (ctl_ts [H*Bp,3], ctl_rs [H*Bp,4], log_ctl [H*Bp,3], mean_v [H,V,3]) -> skin [H,Bp,V,1]; mean_v gets no
    gradient (the reference detaches it, :266)."""

    @staticmethod
    def forward(ctx, ctl_ts, ctl_rs, log_ctl, mean_v, H):
        c = lambda t, sh: t.detach().float().reshape(sh).contiguous()
        Bp = ctl_ts.numel() // (3 * H)
        V = mean_v.numel() // (3 * H)
        ts, rs, lg, mv = c(ctl_ts, (H * Bp, 3)), c(ctl_rs, (H * Bp, 4)), c(log_ctl, (H * Bp, 3)), c(mean_v, (H, V, 3))
        skin = torch.empty(H, Bp, V, dtype=torch.float32, device=mean_v.device)
        _cabi.skin_weights_forward(ts, rs, lg, mv, skin, H, Bp, V)
        ctx.save_for_backward(ts, rs, lg, mv, skin)
        ctx.meta = (H, Bp, V, ctl_ts.shape, ctl_rs.shape, log_ctl.shape)
        return skin[:, :, :, None]

    @staticmethod
    def backward(ctx, g):
        ts, rs, lg, mv, skin = ctx.saved_tensors
        H, Bp, V, s_ts, s_rs, s_lg = ctx.meta
        g = g.float().reshape(H, Bp, V).contiguous()
        g_ts, g_rs, g_lg = torch.empty_like(ts), torch.empty_like(rs), torch.empty_like(lg)
        _cabi.skin_weights_backward(ts, rs, lg, mv, skin, g, g_ts, g_rs, g_lg, H, Bp, V)
        return g_ts.reshape(s_ts), g_rs.reshape(s_rs), g_lg.reshape(s_lg), None, None


def skin_project(verts, Rmat, Tmat, nmesh, skin=None, pp=None, fl=None, tocam=True, depth_window=False, rest_ts=None,
                 n_hypo=1, view_eye_z=None):
    """Fused G1+G2+G3.  verts [N or 1,V,3]; Rmat [N*nmesh,3,3]; Tmat [N*nmesh,(1,)3];
    skin [N or 1,nmesh-1,V(,1)]; pp [N,2] and fl [N(,1)] per FRAME (already repeated over hypotheses).
    rest_ts [n_hypo*(nmesh-1),3]: Rmat / Tmat are the RAW part transforms and are composed about these joint centres
    inside the kernel (mesh_net.py:278-281) -- frames laid out [.., n_hypo].
    view_eye_z (a float): also return the render-space vertices LASR hands the rasterizer, (x, -y, (z + e) - e) [N,V,3]
    (eye offset, y flip, look_at of an eye on the z axis: mesh_net.py:81-83, transform.py:41-48) as a fourth value.
    Returns (cam [N,V,3], px [N,V,4] or None, (near, far) 0-dim device tensors or None[, view])."""
    cam, px, nf, view = SkinProjectFunction.apply(verts, Rmat, Tmat, skin, pp, fl, nmesh, tocam, depth_window, rest_ts,
                                                  n_hypo, view_eye_z)
    out = (cam, (px if pp is not None else None), ((nf[0], nf[1]) if depth_window else None))
    return out + (view,) if view_eye_z is not None else out


def obj_to_cam(verts, Rmat, Tmat, nmesh, skin, tocam=True):
    """nnutils/geom_utils.py:45-71 on CUDA tensors."""
    cam, _, _ = skin_project(verts, Rmat, Tmat, nmesh, skin, None, None, tocam, False)
    return cam


class PinholeFunction(Function):
    """nnutils/geom_utils.py:27-34 on CUDA tensors: a skin_project call with the identity pose."""

    @staticmethod
    def forward(ctx, verts, pp, fl):
        N, V, C = verts.shape
        dev = verts.device
        eye = torch.eye(3, device=dev).repeat(N, 1, 1)
        zero = torch.zeros(N, 3, device=dev)
        xyz = verts[:, :, :3].detach().float().contiguous()
        cam = torch.empty(N, V, 3, device=dev)
        px = torch.empty(N, V, 4, device=dev)
        ws = _cabi.lbs_workspace(N, V, 1, dev)
        ppc, flc = pp.detach().float().reshape(N, 2).contiguous(), fl.detach().float().reshape(N).contiguous()
        _cabi.lbs_project_forward(xyz, eye, zero, None, ppc, flc, cam, px, None, ws, N, V, 1, False, False, False)
        ctx.save_for_backward(xyz, eye, zero, ppc, flc)
        ctx.shapes = (verts.shape, pp.shape, fl.shape)
        return torch.cat([px[:, :, :3], verts[:, :, 3:]], dim=-1) if C > 3 else px[:, :, :3]

    @staticmethod
    def backward(ctx, g):
        xyz, eye, zero, pp, fl = ctx.saved_tensors
        vshape, ppshape, flshape = ctx.shapes
        N, V, C = vshape
        dev = xyz.device
        g = g.float()
        g_px = torch.cat([g[:, :, :3], torch.zeros(N, V, 1, device=dev)], dim=-1).contiguous()
        gv = torch.empty(N, V, 3, device=dev)
        gR, gT = torch.empty(N, 3, 3, device=dev), torch.empty(N, 3, device=dev)
        gpp, gfl = torch.empty(N, 2, device=dev), torch.empty(N, device=dev)
        ws = _cabi.lbs_workspace(N, V, 1, dev)
        _cabi.lbs_project_backward(xyz, eye, zero, None, pp, fl, None, g_px, gv, gR, gT, None, gpp, gfl, ws, N, V, 1,
                                   False, False, False)
        gverts = torch.cat([gv, g[:, :, 3:]], dim=-1) if C > 3 else gv
        return gverts, gpp.reshape(ppshape), gfl.reshape(flshape)


def pinhole_cam(verts, pp, fl):
    n_hypo = verts.shape[0] // pp.shape[0]
    pp_f = pp[:, None].repeat(1, n_hypo, 1).view(-1, 2)
    fl_f = fl.reshape(-1)
    return PinholeFunction.apply(verts, pp_f, fl_f)
