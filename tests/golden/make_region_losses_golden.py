"""Golden vectors for the WHOLE render region from the REFERENCE's own ``LASR.forward`` lines running on top of this
repository's drop-in ``soft_renderer`` package.

Executes, unmodified (see the two notes), ``nnutils/mesh_net.py:259-447`` -- pose -> flow renders (forward / backward,
through the reference's own ``render_flow_soft_2``, lines 75-104) -> texture / silhouette render -> mask, flow and
texture-L1 losses -- on seeded CPU tensors, with

  * ``sr`` = ``lasr_b200.soft_renderer`` (``SoftRenderer`` objects built with LASR's arguments, ``sr.Mesh``), whose
    rasterizer call is routed to the CPU oracle (itself pinned on the reference kernels) because there is no GPU here;
  * ``obj_to_cam`` / ``pinhole_cam`` = the reference's own nnutils/geom_utils.py functions;
  * stand-ins for ``self`` / ``opts``, ``Tensor.cuda`` the identity, LPIPS stubbed to zero (out of scope), ``kornia``'s
    quaternion conversion = this repository's (kornia is not installed);
  * one ``.clone()`` added to line 286 (in-place transpose through an aliased view, see make_region_golden.py).

It is the drop-in claim in one test vector: the reference's region code, calling our package, produces these losses;
``RenderHotPath`` must produce the same from the same inputs.  Build container only (needs /root/reference), no GPU.

    python tests/golden/make_region_losses_golden.py      ->  tests/golden/region_losses_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "region_losses_ref.npz")


def block(lines, first, last, indent):
    out = []
    for l in lines[first - 1:last]:
        if l.startswith(" " * indent):
            out.append(l[indent:])
        else:                         # commented-out timing lines at column 0, blank lines
            assert l.lstrip().startswith("#") or not l.strip(), l
            out.append("\n")
    return "".join(out)


def setup():
    """Load the reference's functions and source lines once; returns the environment run_case needs."""
    import lasr_b200.soft_renderer as sr
    import lasr_b200.soft_renderer.rasterizer as rast
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    from test_host_glue import _OracleRasterize

    def oracle_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                         fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                         aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))
    rast.srf.soft_rasterize = oracle_rasterize

    spec = importlib.util.spec_from_file_location("ref_geom_utils", os.path.join(REF, "nnutils/geom_utils.py"))
    sys.modules.setdefault("soft_renderer", types.ModuleType("soft_renderer"))
    gu = importlib.util.module_from_spec(spec); spec.loader.exec_module(gu)
    with open(os.path.join(REF, "nnutils/mesh_net.py")) as fh:
        lines = fh.readlines()
    flow_fn = block(lines, 75, 104, 0)
    region = block(lines, 259, 447, 8)
    assert flow_fn.startswith("def render_flow_soft_2(") and region.startswith("Rmat = quat.view(-1,3,3).permute(0,2,1)")
    assert region.rstrip().endswith("self.total_loss += self.texture_loss")
    alias = "Rmat[:,1:] = Rmat[:,1:].permute(0,1,3,2)"
    assert region.count(alias) == 1
    region = region.replace(alias, alias + ".clone()")
    torch.Tensor.cuda = lambda self, *a, **k: self
    return types.SimpleNamespace(sr=sr, gu=gu, flow_fn=flow_fn, region=region, q2r=quaternion_to_rotation_matrix)


def run_case(env, bs, H, B, S, seed):
    """Execute the reference's region lines for one configuration; returns (inputs, outputs) as dicts of numpy arrays."""
    from lasr_b200 import synth
    sr, quaternion_to_rotation_matrix = env.sr, env.q2r
    mk = lambda S: sr.SoftRenderer(image_size=S, sigma_val=1e-4, gamma_val=1e-2, camera_mode="look_at", perspective=False,
                                   aggr_func_rgb="softmax", light_mode="vertex", light_intensity_ambient=1.0,
                                   light_intensity_directionals=0.0)       # nnutils/mesh_net.py:136-145
    v, f = synth.icosphere(1, 0.4)
    V = len(v)
    g = torch.Generator().manual_seed(seed)
    n2, N = 2 * bs, 2 * bs * H
    q = torch.randn(N * B, 4, generator=g) * 0.1; q[:, 3] = 1
    quat = quaternion_to_rotation_matrix(q).reshape(N * B, 9)
    trans = torch.randn(N * B, 2, generator=g) * 0.02
    depth = torch.randn(N * B, 1, generator=g) * 0.02
    depth.view(N, B, 1)[:, 0] += 10
    pred_v = torch.from_numpy(v)[None].repeat(N, 1, 1) + torch.randn(N, V, 3, generator=g) * 0.01
    tex = torch.sigmoid(torch.randn(N, V, 3, generator=g))
    faces = torch.from_numpy(f)[None].repeat(n2, 1, 1)
    ppoint = torch.randn(n2, 2, generator=g) * 0.05
    scale = 9.5 + torch.rand(n2, H, generator=g)          # [2bs, H]: what the reference's code predictor hands over
    occ = torch.randn(n2, S, S, generator=g)
    self = types.SimpleNamespace(
        renderer_soft=mk(S), renderer_softflf=mk(S), renderer_softflb=mk(S), renderer_softtex=mk(S), optim_idx=0, iters=1,
        texture_type="vertex", masks=(torch.rand(n2, S, S, generator=g) > 0.4).float(), imgs=torch.rand(n2, 3, S, S, generator=g),
        flow=torch.randn(n2, 3, S, S, generator=g) * 0.05, occ=occ,
        ptex_loss=types.SimpleNamespace(forward_pair=lambda a, b: torch.zeros(a.shape[0])))
    if B > 1:
        self.ctl_ts = torch.from_numpy(v)[:H * (B - 1)].clone()
        self.rest_ts = self.ctl_ts + 0.01
        self.ctl_rs = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(H * (B - 1), 1)
        self.log_ctl = 1 + torch.randn(H * (B - 1), 3, generator=g) * 0.3
    opts = types.SimpleNamespace(n_hypo=H, n_bones=B, sigval=1e-4, opt_tex="yes", img_size=S, l1tex_wt=1.0)
    ins = dict(quat=quat, trans=trans, depth=depth, pred_v=pred_v, tex=tex, faces=faces, ppoint=ppoint, scale=scale,
               masks=self.masks, imgs=self.imgs, flow=self.flow, occ=occ)
    if B > 1:
        ins.update(ctl_ts=self.ctl_ts, rest_ts=self.rest_ts, ctl_rs=self.ctl_rs, log_ctl=self.log_ctl)
    ins = {k: t.clone().numpy() for k, t in ins.items()}          # before the lines run: they write into `quat` through a view
    ns = dict(self=self, opts=opts, torch=torch, np=np, sr=sr, local_batch_size=bs, quat=quat, trans=trans, depth=depth,
              pred_v=pred_v, tex=tex, faces=faces, ppoint=ppoint, scale=scale, obj_to_cam=env.gu.obj_to_cam,
              pinhole_cam=env.gu.pinhole_cam, citylabs=[[10 * k, 20, 30] for k in range(64)],
              kornia=types.SimpleNamespace(quaternion_to_rotation_matrix=quaternion_to_rotation_matrix))
    exec(env.flow_fn, ns)
    exec(env.region, ns)
    outs = {k: np.asarray(getattr(self, k).detach().numpy())
            for k in ("mask_loss", "flow_rd_loss", "texture_loss", "total_loss", "mask_loss_sub", "flow_rd_loss_sub", "texture_loss_sub")}
    outs["mask_pred"] = self.mask_pred.detach().numpy()
    outs["flow_rd"] = self.flow_rd.detach().numpy()
    return ins, outs


def main():
    env = setup()
    out = {}
    for case, (bs, H, B, S, seed) in {"b1": (1, 1, 1, 16, 0), "b3": (1, 1, 3, 16, 1), "b3h2": (1, 2, 3, 12, 2)}.items():
        ins, outs = run_case(env, bs, H, B, S, seed)
        for k, a in ins.items():
            out[f"{case}/in/{k}"] = a
        out[f"{case}/meta"] = np.asarray([bs, H, B, S], np.int64)
        for k, a in outs.items():
            out[f"{case}/out/{k}"] = a
        print(case, {k: round(float(outs[k]), 6) for k in ("mask_loss", "flow_rd_loss", "texture_loss", "total_loss")})
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
