"""Golden vectors for the flow render's pre- and post-processing (R1) from the REFERENCE's own code.

``render_flow_soft_2`` (nnutils/mesh_net.py:75-104) is a module-level function of a module that cannot be imported
without the CUDA extensions.  This script (build container only; no GPU) reads exactly those source lines from the
reference file and executes them unmodified with a stand-in renderer: ``sr.Mesh`` records what it is given, and
``render_mesh`` returns a seeded "rendered" tensor [2n,4,S,S] -- so the outputs are what the reference's lines make
of a given render: the vertices / attributes handed to the rasterizer, and flow / background mask / foreground mask.
``render_flow_soft_3`` (nnutils/geom_utils.py:73-95; used by scripts/eval_badja.py:226 with a 'hard' renderer) is
executed the same way from its own source lines (cases "s3a", "s3b").
Stored in tests/golden/flow_ref.npz.        python tests/golden/make_flow_golden.py
"""
import os
import textwrap
import types

import numpy as np
import torch

SRC = "/root/reference/nnutils/mesh_net.py"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "flow_ref.npz")


def main():
    with open(SRC) as fh:
        src = textwrap.dedent("".join(fh.readlines()[74:104]))
    assert src.startswith("def render_flow_soft_2(") and "return flow_fw, bgmask, fgmask" in src
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    for case, (bs, H, V, F, S, seed) in {"h1": (3, 1, 20, 30, 16, 0), "h2": (2, 2, 12, 18, 12, 1)}.items():
        g = torch.Generator().manual_seed(seed)
        n = bs * H
        verts = torch.randn(n, V, 4, generator=g)
        faces = torch.randint(0, V, (bs, F, 3), generator=g, dtype=torch.int32)
        pos0 = torch.randn(n, V, 4, generator=g) + torch.tensor([0, 0, 10.0, 0])
        pos1 = torch.randn(n, V, 4, generator=g) + torch.tensor([0, 0, 10.0, 0])
        pp0, pp1 = torch.randn(n, 2, generator=g) * 0.1, torch.randn(n, 2, generator=g) * 0.1
        f0, f1 = 9 + torch.rand(n, 1, generator=g), 9 + torch.rand(n, 1, generator=g)
        rendered = torch.rand(2 * n, 4, S, S, generator=g)
        rendered[:, 2] += 9.0                                  # rendered camera-space z ...
        rendered[0, 2, :3] = 0.0                               # ... except background rows (z < 1e-9)
        rendered[n + 1, 2, :, -2:] = 0.0
        seen = {}

        def Mesh(vertices, faces_, textures=None, texture_type="surface"):
            seen.update(vertices=vertices.clone(), faces=faces_.clone(), textures=textures.clone(), texture_type=texture_type)
            return seen
        renderer = types.SimpleNamespace(
            transform=types.SimpleNamespace(transformer=types.SimpleNamespace(_eye=[0, 0, -2.732051])),
            render_mesh=lambda mesh: rendered.clone())
        ns = dict(torch=torch, np=np, sr=types.SimpleNamespace(Mesh=Mesh))
        exec(src, ns)
        flow, bg, fgm = ns["render_flow_soft_2"](renderer, verts, faces, pos0, pos1, pp0, pp1, f0, f1)
        for k, v in dict(verts=verts, faces=faces, pos0=pos0, pos1=pos1, pp0=pp0, pp1=pp1, f0=f0, f1=f1, rendered=rendered).items():
            out[f"{case}/in/{k}"] = v.numpy()
        out[f"{case}/mesh/vertices"] = seen["vertices"].numpy(); out[f"{case}/mesh/faces"] = seen["faces"].numpy()
        out[f"{case}/mesh/textures"] = seen["textures"].numpy()
        out[f"{case}/out/flow"] = flow.numpy(); out[f"{case}/out/bgmask"] = bg.numpy(); out[f"{case}/out/fgmask"] = fgm.numpy()
        print(case, flow.shape, int(bg.sum()), seen["texture_type"])
    with open("/root/reference/nnutils/geom_utils.py") as fh:
        src3 = textwrap.dedent("".join(fh.readlines()[72:95]))
    assert src3.startswith("def render_flow_soft_3(") and src3.rstrip().endswith("return flow_fw, bgmask, fgmask")
    for case, (n, V, F, S, seed) in {"s3a": (1, 20, 30, 16, 5), "s3b": (3, 12, 18, 10, 6)}.items():
        g = torch.Generator().manual_seed(seed)
        verts = torch.randn(n, V, 4, generator=g)
        target = torch.randn(n, V, 4, generator=g)
        faces = torch.randint(0, V, (n, F, 3), generator=g, dtype=torch.int32)
        rendered = torch.rand(n, 4, S, S, generator=g) * 2 - 1
        rendered[:, 2] += 9.0
        rendered[0, 2, 2:5, 1:7] = 0.0                          # background pixels (z < 1e-9)
        seen = {}

        def Mesh(vertices, faces_, textures=None, texture_type="surface"):
            seen.update(vertices=vertices.clone(), faces=faces_.clone(), textures=textures.clone(), texture_type=texture_type)
            return seen
        renderer = types.SimpleNamespace(
            transform=types.SimpleNamespace(transformer=types.SimpleNamespace(_eye=[0, 0, -2.732051])),
            render_mesh=lambda mesh: rendered.clone())
        ns = dict(torch=torch, np=np, sr=types.SimpleNamespace(Mesh=Mesh))
        exec(src3, ns)
        vin = verts.clone()
        flow, bg, fgm = ns["render_flow_soft_3"](renderer, verts, target, faces)
        assert torch.equal(vin, verts)
        for k, v in dict(verts=vin, target=target, faces=faces, rendered=rendered).items():
            out[f"{case}/in/{k}"] = v.numpy()
        out[f"{case}/mesh/vertices"] = seen["vertices"].numpy(); out[f"{case}/mesh/faces"] = seen["faces"].numpy()
        out[f"{case}/mesh/textures"] = seen["textures"].numpy()
        out[f"{case}/out/flow"] = flow.numpy(); out[f"{case}/out/bgmask"] = bg.numpy(); out[f"{case}/out/fgmask"] = fgm.numpy()
        print(case, flow.shape, int(bg.sum()), seen["texture_type"])
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
