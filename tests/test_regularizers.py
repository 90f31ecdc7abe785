"""Mesh regularisers: oracle and host path against golden vectors minted from the
reference's own Python (tests/golden/make_regularizer_golden.py), CUDA kernels against both."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "regularizer_ref.npz"))
TAGS = ("ico1", "ico2")


def _mods(tag, device="cpu"):
    from lasr_b200 import loss_utils as LU
    f = torch.from_numpy(GOLD[f"{tag}_faces"])
    v = torch.from_numpy(GOLD[f"{tag}_verts"])
    return {"laplacian": LU.LaplacianLoss(v, f).to(device), "flatten": LU.FlattenLoss(f).to(device),
            "arap": LU.ARAPLoss(v, f).to(device)}


def _run(mods, tag, device, dtype):
    x = torch.from_numpy(GOLD[f"{tag}_x"]).to(device, dtype)
    dx = torch.from_numpy(GOLD[f"{tag}_dx"]).to(device, dtype)
    go = torch.from_numpy(GOLD[f"{tag}_gout"]).to(device, dtype)
    out = {}
    for name, mod in mods.items():
        xa, da = x.clone().requires_grad_(True), dx.clone().requires_grad_(True)
        loss = mod(da, xa) if name == "arap" else mod(xa)
        (loss * go).sum().backward()
        out[name] = (loss.detach().cpu().double().numpy(), xa.grad.cpu().double().numpy(),
                     da.grad.cpu().double().numpy() if name == "arap" else None)
    return out


def _check(out, tag, rtol, flatten_grad_rtol=None):
    for name, (loss, gx, gdx) in out.items():
        ref = GOLD[f"{tag}_{name}"]
        assert np.abs(loss - ref).max() <= rtol * np.abs(ref).max(), (name, loss, ref)
        rg = GOLD[f"{tag}_{name}_gx"]
        tol = flatten_grad_rtol if (name == "flatten" and flatten_grad_rtol) else rtol
        assert np.linalg.norm(gx - rg) <= tol * np.linalg.norm(rg), (name, "grad_x", np.linalg.norm(gx - rg) / np.linalg.norm(rg))
        if gdx is not None:
            rd = GOLD[f"{tag}_arap_gdx"]
            assert np.linalg.norm(gdx - rd) <= rtol * np.linalg.norm(rd), (name, "grad_dx")


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_python(tag):
    from oracle import regularizers as R
    f, x, dx = GOLD[f"{tag}_faces"], GOLD[f"{tag}_x"], GOLD[f"{tag}_dx"]
    for name, val in (("laplacian", R.laplacian_loss(x, f)), ("flatten", R.flatten_loss(x, f)),
                      ("arap", R.arap_loss(dx, x, f))):
        ref = GOLD[f"{tag}_{name}"]
        assert np.abs(val - ref).max() <= 5e-6 * np.abs(ref).max(), name     # the reference's buffers are fp32


@pytest.mark.parametrize("tag", TAGS)
def test_host_path_matches_reference_python(tag):
    _check(_run(_mods(tag), tag, "cpu", torch.float64), tag, 5e-6)


def test_topology_tables_match_the_dense_matrices():
    from lasr_b200 import loss_utils as LU, synth
    from oracle import regularizers as R
    v, f = synth.bumpy_sphere(2, 0.4)
    off, nbr = LU.vertex_adjacency(len(v), f)
    A = R.dense_adjacency(len(v), f)
    assert off[-1] == int(A.sum()) == len(nbr)
    for a in range(len(v)):
        assert np.array_equal(nbr[off[a]:off[a + 1]], np.nonzero(A[a])[0])
    e = LU.dihedral_edges(f)
    assert np.array_equal(e, R.flatten_edges(f))
    ioff, inc = LU.corner_incidence(len(v), e)
    assert ioff[-1] == 4 * len(e) and np.array_equal(np.sort(inc), np.arange(4 * len(e)))
    for vtx in (0, 5, len(v) - 1):
        assert all(e.reshape(-1)[k] == vtx for k in inc[ioff[vtx]:ioff[vtx + 1]])
    with pytest.raises(ValueError):
        LU.dihedral_edges(f[:-3])                      # open mesh: an edge with one face


def test_average_flag_and_shapes():
    mods = _mods("ico1")
    x = torch.from_numpy(GOLD["ico1_x"])
    from lasr_b200 import loss_utils as LU
    f, v = torch.from_numpy(GOLD["ico1_faces"]), torch.from_numpy(GOLD["ico1_verts"])
    assert mods["laplacian"](x).shape == (3,)
    assert torch.allclose(LU.LaplacianLoss(v, f, average=True)(x), mods["laplacian"](x).sum() / 3)
    assert torch.allclose(LU.FlattenLoss(f, average=True)(x), mods["flatten"](x).sum() / 3)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_kernels_match_reference_python(tag):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the flatten gradient is ill-conditioned on the near-regular ico2 sphere: the reference's own expression
    # evaluated by PyTorch in fp32 is 1.2e-4 away from its fp64 value there (4e-7 on ico1), hence 5e-4
    _check(_run(_mods(tag, "cuda"), tag, "cuda", torch.float32), tag, 2e-5, flatten_grad_rtol=5e-4)


@pytest.mark.gpu
def test_kernels_match_host_path_on_spot_and_are_deterministic():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lasr_b200 import loss_utils as LU, synth
    v, f = synth.load_spot()
    g = torch.Generator().manual_seed(0)
    N, V = 6, len(v)
    vt, ft = torch.from_numpy(v), torch.from_numpy(f)
    x = vt[None] + torch.randn(N, V, 3, generator=g) * 0.01
    dx = vt[None] * 1.05 + torch.randn(N, V, 3, generator=g) * 0.01
    go = torch.randn(N, generator=g)
    mods = {"laplacian": LU.LaplacianLoss(vt, ft), "flatten": LU.FlattenLoss(ft), "arap": LU.ARAPLoss(vt, ft)}

    def run(dev, dt):
        res = {}
        for name, m in mods.items():
            m = m.to(dev)
            xa, da = x.to(dev, dt).requires_grad_(True), dx.to(dev, dt).requires_grad_(True)
            loss = m(da, xa) if name == "arap" else m(xa)
            (loss * go.to(dev, dt)).sum().backward()
            res[name] = [loss.detach().cpu().double(), xa.grad.cpu().double()] + ([da.grad.cpu().double()] if name == "arap" else [])
        return res

    ref = run("cpu", torch.float64)
    a, b = run("cuda", torch.float32), run("cuda", torch.float32)
    for name in mods:
        for t_gpu, t_gpu2, t_ref in zip(a[name], b[name], ref[name]):
            assert torch.equal(t_gpu, t_gpu2), name                                    # bit-reproducible
            assert (t_gpu - t_ref).norm() <= 3e-5 * t_ref.norm() + 1e-9, (name, float((t_gpu - t_ref).norm()), float(t_ref.norm()))


@pytest.mark.gpu
def test_size_independent_properties_at_full_size():
    """Config-D topology (icosphere-5, 10 242 vertices) x 32 frames -- where the reference's ARAP would need
    32 x 10242^2 x 4 B = 13 GB per dense intermediate: invariances instead of an oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lasr_b200 import loss_utils as LU, synth
    v, f = synth.icosphere(5)
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(1)
    vt, ft = torch.from_numpy(v).float(), torch.from_numpy(f)
    N = 32
    x = (vt[None] * 0.5 + torch.randn(N, len(v), 3, generator=g) * 0.003).to(dev)
    lap, arap, flat = LU.LaplacianLoss(vt, ft).to(dev), LU.ARAPLoss(vt, ft).to(dev), LU.FlattenLoss(ft).to(dev)
    q = torch.randn(4, generator=g); q = q / q.norm()
    w, a, b, c = q.tolist()
    R = torch.tensor([[1 - 2 * (b * b + c * c), 2 * (a * b - c * w), 2 * (a * c + b * w)],
                      [2 * (a * b + c * w), 1 - 2 * (a * a + c * c), 2 * (b * c - a * w)],
                      [2 * (a * c - b * w), 2 * (b * c + a * w), 1 - 2 * (a * a + b * b)]], device=dev)
    moved = x @ R.T + torch.tensor([0.3, -0.2, 1.0], device=dev)
    assert float(arap(x, x).abs().max()) == 0.0                                    # identical shapes
    assert float(arap(moved, x).abs().max()) < 1e-6                                 # rigid motion keeps edge lengths
    assert torch.allclose(lap(2 * x), 4 * lap(x), rtol=1e-5)                        # quadratic
    assert torch.allclose(lap(moved), lap(x), rtol=2e-3, atol=1e-7)                 # translation / rotation invariant
    assert torch.allclose(flat(moved), flat(x), rtol=2e-3, atol=1e-4)               # dihedral angles are rigid invariants
    assert torch.allclose(arap(1.5 * x, x), arap(x, 1.5 * x), rtol=1e-5)            # |.| is symmetric in (dx, x)


REF_TREE = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TREE, "nnutils", "loss_utils.py")),
                    reason="reference tree not present (GPU box)")
def test_unusual_topologies_behave_like_the_reference_classes(monkeypatch):
    """Open meshes, isolated vertices, two components, a duplicated face: the host path of the three regularisers
    against the reference's own classes imported live (build container only).  Laplacian and ARAP agree everywhere;
    FlattenLoss needs a closed manifold in both (the reference fails inside its edge bookkeeping, here a ValueError
    says why)."""
    import importlib
    import importlib.util
    import sys
    import types
    from lasr_b200 import loss_utils as LU, synth
    monkeypatch.setitem(sys.modules, "soft_renderer", types.ModuleType("soft_renderer"))      # nnutils/geom_utils.py:24
    pkg = types.ModuleType("nnutils"); pkg.__path__ = [os.path.join(REF_TREE, "nnutils")]
    monkeypatch.setitem(sys.modules, "nnutils", pkg)
    for k in [k for k in sys.modules if k.startswith("nnutils.")]:
        monkeypatch.delitem(sys.modules, k)
    ref_lu = importlib.import_module("nnutils.loss_utils")
    spec = importlib.util.spec_from_file_location("ref_ext_lu", os.path.join(REF_TREE, "third_party/ext_nnutils/loss_utils.py"))
    ext_lu = importlib.util.module_from_spec(spec); spec.loader.exec_module(ext_lu)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)                      # ARAPLoss allocates with .cuda()
    v, f = synth.icosphere(1)
    v2, f2 = synth.bumpy_sphere(2)
    meshes = {"closed": (v, f, True), "open": (v, f[12:], False), "strip": (v, f[:20], False), "bumpy2": (v2, f2, True),
              "two components": (np.concatenate([v, v + 3]), np.concatenate([f, f + len(v)]), True),
              "duplicate face": (v, np.concatenate([f, f[:1]]), False)}
    for name, (vv, ff, closed) in meshes.items():
        g = torch.Generator().manual_seed(1)
        V = len(vv)
        verts, faces = torch.from_numpy(vv).float(), torch.from_numpy(ff.astype(np.int64))
        x = (verts[None] + torch.randn(2, V, 3, generator=g) * 0.08).double()
        dx = (verts[None] * 1.1 + torch.randn(2, V, 3, generator=g) * 0.08).double()
        pairs = (("laplacian", lambda: ext_lu.LaplacianLoss(verts, faces).double()(x.clone()), lambda: LU.LaplacianLoss(verts, faces).double()(x.clone())),
                 ("arap", lambda: ref_lu.ARAPLoss(verts, faces).double()(dx.clone(), x.clone()), lambda: LU.ARAPLoss(verts, faces).double()(dx.clone(), x.clone())))
        for what, ref, ours in pairs:
            r, o = ref().detach().numpy(), ours().detach().numpy()
            assert np.abs(r - o).max() <= 5e-7 * np.abs(r).max(), (name, what)          # the reference keeps float32 weights
        if closed:
            r = ext_lu.FlattenLoss(faces).double()(x.clone()).detach().numpy()
            o = LU.FlattenLoss(faces).double()(x.clone()).detach().numpy()
            assert np.abs(r - o).max() <= 1e-12 * np.abs(r).max(), name
        else:
            with pytest.raises(Exception):
                ext_lu.FlattenLoss(faces).double()(x.clone())
            with pytest.raises(ValueError):
                LU.FlattenLoss(faces)
