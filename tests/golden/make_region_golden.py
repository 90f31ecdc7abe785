"""Golden vectors for the geometry half of the render region from the REFERENCE's own code.

Executes, unmodified, the source lines of ``LASR.forward`` that pose the mesh (nnutils/mesh_net.py:259-311: bone
transform composition, skinning weights, linear-blend skinning, pinhole projection, the batch-global depth window)
on seeded CPU tensors, with the reference's own ``obj_to_cam`` / ``pinhole_cam`` (nnutils/geom_utils.py, imported by
path).  Stand-ins: ``self`` / ``opts`` namespaces carrying what the lines read, ``Tensor.cuda`` the identity, the three
renderers as plain namespaces (the lines only set ``rasterizer.near / far``), ``citylabs`` a colour list, and ``kornia``
(0.5.3, not installed here) a namespace whose ``quaternion_to_rotation_matrix`` is this repository's XYZW conversion --
so everything except that one conversion is the reference's arithmetic (plus one `.clone()` that removes an
aliasing hazard of the CPU execution, see the comment in main()).  Build container only, no GPU.

    python tests/golden/make_region_golden.py        ->  tests/golden/region_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "region_ref.npz")


def main():
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    sys.modules.setdefault("soft_renderer", types.ModuleType("soft_renderer"))
    spec = importlib.util.spec_from_file_location("ref_geom_utils", os.path.join(REF, "nnutils/geom_utils.py"))
    gu = importlib.util.module_from_spec(spec); spec.loader.exec_module(gu)
    with open(os.path.join(REF, "nnutils/mesh_net.py")) as fh:
        lines = fh.readlines()[258:311]
    # the block is indented by 8 columns, except two commented-out timing lines that start at column 0
    src = "".join(l[8:] if l.startswith(" " * 8) else ("\n" if l.lstrip().startswith("#") or not l.strip() else l) for l in lines)
    assert src.startswith("Rmat = quat.view(-1,3,3).permute(0,2,1)") and "self.renderer_softtex.rasterizer.far=" in src
    # ONE token is added to the reference's text.  Line 286 transposes the part rotations in place through a view of the
    # same storage (`Rmat[:,1:] = Rmat[:,1:].permute(0,1,3,2)`): source and destination of the copy alias.  On the GPU the
    # elementwise copy reads a whole 3x3 block before it writes it and the result is the transpose the author meant; on
    # the CPU the copy runs sequentially and leaves a symmetrised matrix (checked with torch 2.11).  The golden records the
    # intended (GPU) behaviour: the right-hand side is cloned first.
    alias = "Rmat[:,1:] = Rmat[:,1:].permute(0,1,3,2)"
    assert src.count(alias) == 1
    src = src.replace(alias, alias + ".clone()")
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    for case, (bs, H, B, V, seed) in {"b1": (2, 1, 1, 30, 0), "b4": (2, 1, 4, 40, 1), "b3h2": (1, 2, 3, 25, 2)}.items():
        g = torch.Generator().manual_seed(seed)
        n2 = 2 * bs
        N = n2 * H
        q = torch.randn(N * B, 4, generator=g) * 0.15; q[:, 3] = 1
        quat = quaternion_to_rotation_matrix(q).reshape(N * B, 9)                 # rotation matrices, flattened
        trans = torch.randn(N * B, 2, generator=g) * 0.02
        depth = torch.randn(N * B, 1, generator=g) * 0.02
        depth.view(N, B, 1)[:, 0] += 10
        pred_v = torch.randn(N, V, 3, generator=g) * 0.3
        ppoint = torch.randn(n2, 2, generator=g) * 0.05
        scale = 9.5 + torch.rand(N, 1, generator=g)
        mk = lambda: types.SimpleNamespace(rasterizer=types.SimpleNamespace(near=None, far=None, sigma_val=None))
        self = types.SimpleNamespace(renderer_soft=mk(), renderer_softflf=mk(), renderer_softflb=mk(), renderer_softtex=mk(),
                                     optim_idx=0)
        if B > 1:
            self.ctl_ts = torch.randn(H * (B - 1), 3, generator=g) * 0.3
            self.rest_ts = self.ctl_ts + 0.01
            cr = torch.randn(H * (B - 1), 4, generator=g) * 0.2; cr[:, 3] = 1
            self.ctl_rs = cr
            self.log_ctl = 1 + torch.randn(H * (B - 1), 3, generator=g) * 0.3
        opts = types.SimpleNamespace(n_hypo=H, n_bones=B, sigval=1e-4)
        ns = dict(self=self, opts=opts, torch=torch, np=np, local_batch_size=bs, quat=quat, trans=trans, depth=depth,
                  pred_v=pred_v, ppoint=ppoint, scale=scale, obj_to_cam=gu.obj_to_cam, pinhole_cam=gu.pinhole_cam,
                  kornia=types.SimpleNamespace(quaternion_to_rotation_matrix=quaternion_to_rotation_matrix),
                  citylabs=[[10 * k, 20, 30] for k in range(64)])
        ins = dict(quat=quat, trans=trans, depth=depth, pred_v=pred_v, ppoint=ppoint, scale=scale)
        if B > 1:
            ins.update(ctl_ts=self.ctl_ts, rest_ts=self.rest_ts, ctl_rs=self.ctl_rs, log_ctl=self.log_ctl)
        for k, v in ins.items():                       # before the lines run: they write into `quat` through a view
            out[f"{case}/in/{k}"] = v.clone().numpy()
        exec(src, ns)
        out[f"{case}/meta"] = np.asarray([bs, H, B, V], np.int64)
        res = dict(Rmat=ns["Rmat"], Tmat=ns["Tmat"], deform_v=self.deform_v, verts_fl=ns["verts_fl"],
                   verts_pos0=ns["verts_pos0"], verts_pos1=ns["verts_pos1"],
                   near=self.renderer_softtex.rasterizer.near, far=self.renderer_softtex.rasterizer.far)
        if B > 1:
            res["skin"] = ns["skin"]
        for k, v in res.items():
            out[f"{case}/out/{k}"] = torch.as_tensor(v).detach().numpy()
        print(case, tuple(ns["verts_fl"].shape), float(res["near"]), float(res["far"]))
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
