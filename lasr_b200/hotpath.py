"""The render hot path of ``LASR.forward`` (nnutils/mesh_net.py:255-447) as one call.

Pose in, three image losses out, with every stage served by this package:
skinning weights (S0) -> fused LBS + camera + projection + depth window (G1-G3) -> forward /
backward flow renders and the texture render through the indexed soft rasterizer (R1-R4) ->
fused silhouette / flow / texture-L1 losses (L1-L3).  It is what a maintainer would call from
``LASR.forward`` in place of lines 255-447, and what the tests use to check the whole chain against
an oracle-backed CPU composition.  Tensor layout follows the reference: a batch holds ``bs`` frame
pairs as [first frames; second frames] (2*bs items), each with ``H`` hypotheses and ``B`` bones.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import geom_utils as gu
from . import losses as L
from . import soft_renderer as sr


@dataclass
class HotPathOptions:
    img_size: int = 256
    n_hypo: int = 1
    n_bones: int = 1
    sigval: float = 1e-4
    l1tex_wt: float = 1.0
    opt_tex: bool = True
    fused: bool = True          # CUDA only: one three-set render + fused image head instead of three renders
    report_flow: bool = True    # fused path: also materialise flow_rd / bgmask (reporting; costs one write)
    packed_grad: bool = os.environ.get("LASR_B200_PACKED_GRAD", "1") != "0"   # fused path: the image gradient travels packed per pixel between the head's and the
                                # rasterizer's backward kernels; the rendered image is then a reporting output only


class RenderHotPath(torch.nn.Module):
    """Holds the three SoftRenderers LASR builds for this region (mesh_net.py:136-145)."""

    def __init__(self, opts: HotPathOptions):
        super().__init__()
        self.opts = opts
        mk = lambda: sr.SoftRenderer(image_size=opts.img_size, sigma_val=opts.sigval, gamma_val=1e-2,
                                     camera_mode="look_at", perspective=False, aggr_func_rgb="softmax",
                                     light_mode="vertex", light_intensity_ambient=1.0,
                                     light_intensity_directionals=0.0)
        self.renderer_softflf, self.renderer_softflb, self.renderer_softtex = mk(), mk(), mk()

    def forward(self, pred_v, tex, faces, Rmat, Tmat, ppoint, scale, batch, bones=None, shared_faces=None):
        """pred_v [2bs*H,V,3]; tex [2bs*H,V,3]; faces [2bs,F,3] int32; Rmat [2bs*H*B,3,3] (already
        transposed as mesh_net.py:259); Tmat [2bs*H*B,3]; ppoint [2bs,2]; scale [2bs*H,1];
        batch = dict(masks [2bs,S,S], imgs [2bs,3,S,S], flow [2bs,3,S,S], occ [2bs,S,S]);
        bones = dict(ctl_ts, ctl_rs, log_ctl, rest_ts) when B > 1.
        Returns dict(total, mask_loss, flow_loss, texture_loss, + renders)."""
        o = self.opts
        H, B, S = o.n_hypo, o.n_bones, o.img_size
        if shared_faces is None and faces.is_cuda:
            # LASR renders one mesh: every frame carries the same topology (ext_nnutils/mesh_net.py:171-185).
            # Verified once per faces tensor (cached); pass shared_faces= to skip even that.
            shared_faces = sr.functional.shared_topology(faces)
        n2 = ppoint.shape[0]          # 2 * bs
        bs = n2 // 2
        # (the fused route renders through the indexed kernels: one topology for every frame -- LASR's case; a batch of
        #  different meshes takes the three-render route, which gathers faces per frame as the reference does)
        fused_route = o.fused and pred_v.is_cuda and o.opt_tex and shared_faces is not None
        if pred_v.is_cuda:
            # 1') bone composition + LBS + camera + projection + depth window in ONE kernel (lasr_b200_lbs_pose_forward /
            #     lasr_b200_lbs_project_forward): the raw part transforms go in, the joint centres ride along.  Both CUDA
            #     routes pose the mesh this way, so they rasterize bit-identical vertices.
            from . import native_geom
            skin = rest = None
            if B > 1:
                skin = gu.skin_weights(bones["ctl_ts"], bones["ctl_rs"], bones["log_ctl"],
                                       pred_v.view(n2, H, -1, 3)[0].detach(), H)
                if H > 1:                      # the kernel reads one shared [1,B-1,V] table only when there is one hypothesis
                    skin = skin.repeat(n2, 1, 1, 1)
                rest = bones["rest_ts"]
            pp_rep = ppoint if H == 1 else ppoint[:, None].repeat(1, H, 1).view(n2 * H, 2)
            eye_z = self._plain_view_eye_z(self.renderer_softtex) if fused_route else None
            posed = native_geom.skin_project(pred_v, Rmat, Tmat, B, skin, pp_rep, scale.view(-1), tocam=True,
                                             depth_window=True, rest_ts=rest, n_hypo=H, view_eye_z=eye_z)
            cam, verts_px, (near, far) = posed[:3]
            if fused_route:
                self.renderer_softtex.rasterizer.near, self.renderer_softtex.rasterizer.far = near, far
                self.renderer_softtex.rasterizer.sigma_val = o.sigval
                return self._forward_fused(cam, verts_px, tex, faces, ppoint, scale, batch, shared_faces, near, far,
                                           view=posed[3] if eye_z is not None else None)
            verts_fl = torch.cat([cam, torch.ones_like(cam[:, :, :1])], dim=-1)
        else:
            skin, Rmat, Tmat = self.bone_transforms(pred_v, Rmat, Tmat, bones, n2)
            # 1) LBS + camera (mesh_net.py:298-311)
            cam = gu.obj_to_cam(pred_v, Rmat, Tmat[:, None, :], B, H, skin)
            verts_fl = torch.cat([cam, torch.ones_like(cam[:, :, :1])], dim=-1)
            verts_px = gu.pinhole_cam(verts_fl, ppoint, scale)
            near, far = gu.depth_window(verts_px[:, :, 2])
        verts_pos0 = verts_fl.view(n2, H, -1, 4)[:bs].reshape(bs * H, -1, 4)
        verts_pos1 = verts_fl.view(n2, H, -1, 4)[bs:].reshape(bs * H, -1, 4)
        for r in (self.renderer_softflf, self.renderer_softflb, self.renderer_softtex):
            r.rasterizer.near, r.rasterizer.far = near, far
            r.rasterizer.sigma_val = o.sigval
        # 2) flow renders (mesh_net.py:318-335)
        px4 = verts_px.view(n2, H, -1, 4)
        rep = lambda t, lo, hi: t[:, None][lo:hi].repeat(1, H, 1).view(-1, t.shape[-1])
        flow_fw, bg_fw, fg_fw = gu.render_flow_soft_2(
            self.renderer_softflf, px4[:bs].reshape(bs * H, -1, 4), faces[:bs], verts_pos0, verts_pos1,
            rep(ppoint, 0, bs), rep(ppoint, bs, n2), scale.view(n2, H, 1)[:bs].reshape(-1, 1),
            scale.view(n2, H, 1)[bs:].reshape(-1, 1), shared_faces=shared_faces)
        flow_bw, bg_bw, fg_bw = gu.render_flow_soft_2(
            self.renderer_softflb, px4[bs:].reshape(bs * H, -1, 4), faces[bs:], verts_pos1, verts_pos0,
            rep(ppoint, bs, n2), rep(ppoint, 0, bs), scale.view(n2, H, 1)[bs:].reshape(-1, 1),
            scale.view(n2, H, 1)[:bs].reshape(-1, 1), shared_faces=shared_faces)
        bgmask = torch.cat([bg_fw, bg_bw], 0)
        flow_rd = torch.cat([flow_fw, flow_bw], 0)
        # 3) texture / silhouette render (mesh_net.py:348-363)
        offset = gu._eye_offset(self.renderer_softtex, verts_px.device)
        pre = verts_px[:, :, :3] + offset
        pre = torch.cat([pre[:, :, :1], -pre[:, :, 1:2], pre[:, :, 2:]], dim=-1)
        self.renderer_softtex.rasterizer.background_color = [1, 1, 1]
        faces_h = faces[:, None].repeat(1, H, 1, 1).view(-1, faces.shape[1], 3)
        texture_render = self.renderer_softtex.render_mesh(sr.Mesh(pre, faces_h, textures=tex, texture_type="vertex",
                                                                   shared_faces=shared_faces))
        mask_pred = texture_render[:, -1]
        rgb = texture_render[:, :3]
        # 4) losses (mesh_net.py:372-447, LPIPS term excluded)
        mp = mask_pred.view(n2, H, S, S)
        fr = flow_rd.view(n2, H, S, S, 2)
        bgm = bgmask.view(n2, H, S, S)
        tr = rgb.reshape(n2, H, 3, S, S)
        if mp.is_cuda and o.opt_tex:
            mask_loss, flow_loss, texture_loss, sub = L.image_losses(mp, batch["masks"], batch["occ"], fr, bgm,
                                                                     batch["flow"], tr, mp, batch["imgs"], o.l1tex_wt)
            total = sub["total"]          # (mask + flow) + texture, out of the loss kernels' own launch
        else:
            mask_loss, _ = L.mask_loss(mp, batch["masks"], batch["occ"])
            flow_loss, _, _ = L.flow_loss(fr, bgm, batch["flow"], batch["occ"], batch["masks"])
            texture_loss = (L.texture_l1_loss(tr, mp, batch["imgs"], batch["masks"], batch["occ"], o.l1tex_wt)[0]
                            if o.opt_tex else mask_loss * 0)
            total = mask_loss + flow_loss + (texture_loss if o.opt_tex else 0)
        return dict(total=total, mask_loss=mask_loss, flow_loss=flow_loss, texture_loss=texture_loss,
                    mask_pred=mask_pred, texture_render=rgb, flow_rd=flow_rd, bgmask=bgmask, verts_px=verts_px,
                    near=near, far=far)

    def bone_transforms(self, pred_v, Rmat, Tmat, bones, n2, repeat_skin=True):
        """Skinning weights and the part transforms composed about their rest positions (mesh_net.py:261-281).
        Returns (skin or None, Rmat [n2*H*B,3,3], Tmat [n2*H*B,3]) ready for ``obj_to_cam``."""
        H, B = self.opts.n_hypo, self.opts.n_bones
        if B <= 1:
            return None, Rmat, Tmat
        skin = gu.skin_weights(bones["ctl_ts"], bones["ctl_rs"], bones["log_ctl"],
                               pred_v.view(n2, H, -1, 3)[0].detach(), H)
        if repeat_skin:
            skin = skin.repeat(n2, 1, 1, 1)
        rest_ts = bones["rest_ts"][:, None, :, None].repeat(n2, 1, 1, 1).view(-1, B - 1, 3, 1)
        R4 = Rmat.view(-1, B, 3, 3)
        T4 = Tmat.view(-1, B, 3, 1)
        T_parts = -R4[:, 1:].matmul(rest_ts) + T4[:, 1:] + rest_ts
        Tmat = torch.cat([T4[:, :1], T_parts], 1).view(-1, 3)
        Rmat = torch.cat([R4[:, :1], R4[:, 1:].permute(0, 1, 3, 2)], 1).view(-1, 3, 3)
        return skin, Rmat, Tmat

    @staticmethod
    def _plain_view_eye_z(r):
        """eye_z when LASR's renderer settings make lighting and camera a fixed map of the projected vertices -- ambient
        light 1 / directional 0 (textures x 1), look_at from an eye ON the negative z axis, orthographic, scale 1 (mesh_net.py:132-149):
        +eye, y flip, look_at(eye) is then (x, -y, (z + e) - e), which the LBS kernel writes itself; None otherwise."""
        t = r.transform.transformer
        eye = getattr(t, "_eye", None)
        ok = (r.lighting.is_identity() and r.transform.camera_mode == "look_at" and not t.perspective
              and t.viewing_scale == 1 and isinstance(eye, (list, tuple)) and len(eye) == 3
              and float(eye[0]) == 0.0 and float(eye[1]) == 0.0 and float(eye[2]) < 0.0      # (+z eyes mirror x)
              and not r.rasterizer.anti_aliasing)
        return float(eye[2]) if ok else None

    def _forward_fused(self, cam, verts_px, tex, faces, ppoint, scale, batch, shared_faces, near, far, view=None):
        """Steps 2-4 of ``forward`` with the mesh rasterized ONCE per frame.  The reference renders every
        posed frame three times with the same geometry and renderer settings but different vertex
        attributes -- flow (own + paired camera-space positions, mesh_net.py:77-87, 318-335) and texture
        (:348-363); here the three attribute sets ride one geometry pass (render_indexed with
        [3,N,V,3] colours) and one fused head turns the 10-plane image into the three losses.  Same
        values as the three-render path; the vertex gradient is the sum the reference's autograd forms."""
        o = self.opts
        H, S = o.n_hypo, o.img_size
        n2 = ppoint.shape[0]
        bs = n2 // 2
        nh = n2 * H
        r = self.renderer_softtex
        pre = None
        if view is None:
            from .soft_renderer.functional.geometry import _const
            offset = gu._eye_offset(r, verts_px.device)
            pre = (verts_px[:, :, :3] + offset) * _const([1.0, -1.0, 1.0], verts_px.device)   # y flip (:83), exact
        own = cam                                            # camera-space positions [n2*H,V,3]
        texc = tex.reshape(nh, -1, 3).to(own.dtype)
        faces_h = faces if H == 1 else faces[:, None].repeat(1, H, 1, 1).view(-1, faces.shape[1], 3)
        link = sr.functional.PackedGradLink() if (o.packed_grad and torch.is_grad_enabled()) else None
        bg9 = [0, 0, 0, 0, 0, 0, 1, 1, 1]
        rep = (lambda t: t) if H == 1 else (lambda t: t[:, None].repeat(1, H, 1).view(nh, 2))   # one hypothesis: no copies
        pp_own = rep(ppoint)
        f_own = scale.view(n2, H, -1)[:, :, 0].reshape(nh)
        # frame i pairs with frame i +- bs (mesh_net.py:318-335): entry (n + bs*H) mod n2*H of the same tensors
        # (needs the one shared topology `forward` looked up; meshes that differ per frame take the gathered path below)
        shifted = (view is not None and shared_faces is not None
                   and (link is not None or not torch.is_grad_enabled()))
        if shifted:
            # the LBS kernel already wrote what lighting (x 1) and the camera make of the mesh, and the pair's positions
            # and intrinsics are the own ones read half a batch further on: no rolled copies, no stacked colour tensor
            rz = r.rasterizer
            image = sr.functional.render_sets(view, (own, own, texc), (0, bs * H, 0), shared_faces, rz.image_size, bg9, near, far,
                                              rz.fill_back, rz.eps, rz.sigma_val, rz.dist_func, rz.dist_eps, rz.gamma_val,
                                              rz.aggr_func_rgb, rz.aggr_func_alpha, grad_link=link)
            pp_pair, f_pair, pair_shift = pp_own, f_own, bs * H
        else:
            pair = torch.roll(own.view(n2, H, -1, 3), bs, 0).reshape(nh, -1, 3)
            sets = torch.stack([own, pair, texc], 0)
            r.rasterizer.background_color = bg9
            r.rasterizer.detached_sets = 1          # own positions enter the flow only under .detach() (:102-103)
            r.rasterizer.grad_link = link
            try:
                if view is not None:
                    r.set_texture_mode("vertex")
                    image = r.rasterizer(sr.Mesh(view, faces_h, textures=sets, texture_type="vertex", shared_faces=shared_faces))
                else:
                    image = r.render_mesh(sr.Mesh(pre, faces_h, textures=sets, texture_type="vertex", shared_faces=shared_faces))
            finally:                                # the renderer is shared state: never leave the K = 3 settings behind
                r.rasterizer.background_color = [1, 1, 1]
                r.rasterizer.detached_sets = 0
                r.rasterizer.grad_link = None
            pp_pair = rep(torch.roll(ppoint, bs, 0))
            f_pair = torch.roll(scale.view(n2, H, -1)[:, :, 0], bs, 0).reshape(nh)
            pair_shift = 0
        mask_loss, flow_loss, texture_loss, sub, flow_rd, bgmask = L.fused_head(
            image, pp_own, pp_pair, f_own, f_pair, batch["imgs"], batch["masks"], batch["occ"], batch["flow"], n2, H,
            o.l1tex_wt, want_flow=o.report_flow, grad_link=link, pair_shift=pair_shift)
        total = sub.pop("total")            # (mask + flow) + texture, out of the head's own launch
        shown = image.detach() if link is not None else image     # packed gradients: the image is for reporting only
        return dict(total=total, mask_loss=mask_loss, flow_loss=flow_loss, texture_loss=texture_loss,
                    mask_pred=shown[:, 9], texture_render=shown[:, 6:9], flow_rd=flow_rd, bgmask=bgmask,
                    verts_px=verts_px, near=near, far=far)
