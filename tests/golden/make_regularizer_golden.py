"""Golden vectors for the mesh regularisers, minted from the REFERENCE's own Python.

Imports (read-only, this container only) nnutils/loss_utils.py (ARAPLoss) and
third_party/ext_nnutils/loss_utils.py (LaplacianLoss, FlattenLoss) from /root/reference and records
their outputs AND autograd gradients on seeded inputs over small closed meshes.  ARAPLoss.forward
allocates with a hard-coded ``.cuda()`` (loss_utils.py:50-51); there is no GPU here, so the generator
runs it with ``torch.Tensor.cuda`` patched to the identity -- the arithmetic is untouched.

    python tests/golden/make_regularizer_golden.py     ->  tests/golden/regularizer_ref.npz
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "regularizer_ref.npz")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    from lasr_b200 import synth
    sys.modules.setdefault("soft_renderer", types.ModuleType("soft_renderer"))   # nnutils/geom_utils.py:24
    pkg = types.ModuleType("nnutils")
    pkg.__path__ = [os.path.join(REF, "nnutils")]
    sys.modules["nnutils"] = pkg
    ref_lu = importlib.import_module("nnutils.loss_utils")
    ext_lu = load(os.path.join(REF, "third_party/ext_nnutils/loss_utils.py"), "ref_ext_loss_utils")
    torch.Tensor.cuda = lambda self, *a, **k: self           # see module docstring

    out = {}
    for tag, level in (("ico1", 1), ("ico2", 2)):
        v, f = synth.icosphere(level)
        g = torch.Generator().manual_seed(level)
        V = len(v)
        N = 3
        verts = torch.from_numpy(v).float()
        faces = torch.from_numpy(f).long()
        x = (verts[None] + torch.randn(N, V, 3, generator=g) * 0.08).double()
        dx = (verts[None] * 1.1 + torch.randn(N, V, 3, generator=g) * 0.08).double()
        gout = torch.randn(N, generator=g).double()
        out[f"{tag}_verts"], out[f"{tag}_faces"] = v.astype(np.float32), f.astype(np.int32)
        out[f"{tag}_x"], out[f"{tag}_dx"], out[f"{tag}_gout"] = x.numpy(), dx.numpy(), gout.numpy()

        lap = ext_lu.LaplacianLoss(verts, faces).double()
        xa = x.clone().requires_grad_(True)
        l = lap(xa)
        (l * gout).sum().backward()
        out[f"{tag}_laplacian"], out[f"{tag}_laplacian_gx"] = l.detach().numpy(), xa.grad.numpy()

        fl = ext_lu.FlattenLoss(faces)
        xa = x.clone().requires_grad_(True)
        l = fl(xa)
        (l * gout).sum().backward()
        out[f"{tag}_flatten"], out[f"{tag}_flatten_gx"] = l.detach().numpy(), xa.grad.numpy()

        arap = ref_lu.ARAPLoss(verts, faces)            # float32 buffers + float32 zeros inside forward
        xa = x.float().clone().requires_grad_(True)
        da = dx.float().clone().requires_grad_(True)
        l = arap(da, xa)
        (l * gout.float()).sum().backward()
        out[f"{tag}_arap"], out[f"{tag}_arap_gx"], out[f"{tag}_arap_gdx"] = (l.detach().numpy(), xa.grad.numpy(),
                                                                              da.grad.numpy())
    np.savez_compressed(DST, **out)
    print("wrote", DST, {k: v.shape for k, v in out.items() if "ico1" in k})


if __name__ == "__main__":
    main()
