"""Pin the CPU oracle against outputs of the REFERENCE kernels themselves.

tests/golden/softras_ref_*.npz were produced on a B200 by tests/golden/make_softras_golden.py
from the unmodified reference SoftRas CUDA extension (oracle/_ref).  The oracle (with nvcc's FMA
association, its default) must reproduce them within the north-star tolerance: RGBA 1e-4
relative with ZERO outlier pixels, gradients 1e-3."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

import cases
import parity

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "softras_ref_*.npz")))


def _name(path):
    return os.path.basename(path)[len("softras_ref_"):-4]


def test_golden_fixtures_present():
    assert len(FILES) >= 8, "golden vectors missing: run tests/golden/make_softras_golden.py on a GPU box"
    assert set(_name(f) for f in FILES) == set(cases.GOLDEN_CASES)


@pytest.mark.parametrize("path", FILES, ids=[_name(f) for f in FILES])
def test_oracle_reproduces_reference(path, oracle):
    name = _name(path)
    g = np.load(path)
    c = cases.GOLDEN_CASES[name]()
    h = hashlib.sha256(c["face_vertices"].tobytes() + c["textures"].tobytes()
                       + c["grad_soft_colors"].tobytes()).hexdigest()
    assert h == str(g["inputs_sha256"]), "case inputs drifted from the ones the golden vectors were minted on"
    assert json.loads(str(g["settings"])) == json.loads(json.dumps(c["settings"]))
    oracle.set_contract(True)
    r = oracle.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32)
    gf, gt = oracle.soft_rasterize_backward(r, c["grad_soft_colors"])
    ir = parity.image_report(r["soft_colors"], g["soft_colors"], rel=1e-4)
    assert ir["n_bad"] == 0, (name, ir)
    gfr = parity.grad_report(gf.reshape(g["grad_faces"].shape), g["grad_faces"])
    gtr = parity.grad_report(gt.reshape(g["grad_textures"].shape), g["grad_textures"])
    if gfr["scale"] > 1e-20:
        assert gfr["rel_l2"] < 1e-3 and gfr["max_over_scale"] < 1e-3, (name, gfr)
    assert gtr["rel_l2"] < 1e-3 and gtr["max_over_scale"] < 1e-3, (name, gtr)


def test_brute_force_scan_equals_accelerated(oracle):
    c = cases.lasr_case("bumpy2", 2, 48)
    a = oracle.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32, accel=True)
    b = oracle.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32, accel=False)
    assert np.array_equal(a["soft_colors"], b["soft_colors"]) and np.array_equal(a["aggrs_info"], b["aggrs_info"])
    assert a["pairs"] == b["pairs"]
    ga = oracle.soft_rasterize_backward(a, c["grad_soft_colors"], accel=True)
    gb = oracle.soft_rasterize_backward(b, c["grad_soft_colors"], accel=False)
    assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])


def test_oracle_texture_gradient_matches_finite_differences(oracle):
    """The rendered colour is linear in the textures, so the texture gradient must equal central
    differences of the forward exactly (fp64 instantiation).  The VERTEX gradient of SoftRas is an
    approximation by design (the reference back-propagates through the fragment probability and the
    face depths only, not through the clipped barycentrics: soft_rasterize_cuda_kernel.cu:631-655),
    so it is pinned against the reference kernels' outputs, not against calculus."""
    c = cases.lasr_case("bumpy1", 1, 24, sigma=3e-3)
    s = dict(c["settings"])
    fv = c["face_vertices"].astype(np.float64)
    tex = c["textures"].astype(np.float64)
    g = np.random.RandomState(3).randn(1, 4, 24, 24)
    r = oracle.soft_rasterize_forward(fv, tex, s, np.float64)
    _, gt = oracle.soft_rasterize_backward(r, g)

    def loss(tex_):
        return float((oracle.soft_rasterize_forward(fv, tex_, s, np.float64)["soft_colors"] * g).sum())

    for flat in np.argsort(-np.abs(gt).reshape(-1))[:4]:
        j = np.unravel_index(flat, gt.shape)
        p, m = tex.copy(), tex.copy()
        p[j] += 1e-3
        m[j] -= 1e-3
        fd = (loss(p) - loss(m)) / 2e-3
        assert abs(fd - gt[j]) <= 1e-8 * max(1.0, abs(gt[j])) + 1e-10, (j, fd, gt[j])


def test_reference_driver_sets_up_buffers_like_the_reference_function(monkeypatch):
    """oracle/ref_gpu.py drives the reference's compiled kernels (golden minting, `bench.py --impl reference`) behind a
    restatement of the buffer set-up in the reference's SoftRasterizeFunction.  tests/golden/function_glue_ref.npz holds
    every argument the reference's own Python (functional/soft_rasterize.py, run unmodified with a recording stand-in for
    the pybind module) hands to forward_/backward_soft_rasterize; the restatement must hand over the same."""
    import importlib.util
    import json
    import torch
    from oracle import ref_gpu
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_function_glue_golden", os.path.join(here, "golden", "make_function_glue_golden.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "golden", "function_glue_ref.npz"))
    for name in gen.CASES:
        s = json.loads(str(g[f"{name}/settings"]))
        assert s == json.loads(json.dumps(gen.CASES[name]))
        rec = gen.Recorder(1)
        monkeypatch.setattr(ref_gpu, "ext", lambda: rec)
        fv = torch.from_numpy(g[f"{name}/in/face_vertices"]).requires_grad_(True)
        tex = torch.from_numpy(g[f"{name}/in/textures"]).requires_grad_(True)
        img = ref_gpu.soft_rasterize(fv, tex, **s)
        img.backward(torch.from_numpy(g[f"{name}/in/grad"]))
        for tag, names in (("fwd", gen.FWD_T), ("bwd", gen.BWD_T)):
            tensors, scalars = rec.calls[tag]
            for n, t in zip(names, tensors):
                ref = g[f"{name}/{tag}/{n}"]
                assert t.dtype == torch.float32 and tuple(t.shape) == ref.shape and t.is_contiguous(), (name, tag, n)
                assert np.array_equal(t.detach().numpy(), ref), (name, tag, n)
            want = json.loads(str(g[f"{name}/{tag}/scalars"]))
            assert len(scalars) == len(want) == len(gen.SCALARS)
            for k, a, b in zip(gen.SCALARS, scalars, want):
                assert type(a) is type(b) or not isinstance(b, bool), (name, tag, k)
                assert float(a) == float(b), (name, tag, k, a, b)
        assert np.array_equal(img.detach().numpy(), g[f"{name}/out/image"])
        assert np.array_equal(fv.grad.numpy(), g[f"{name}/out/grad_faces"])
        assert np.array_equal(tex.grad.numpy(), g[f"{name}/out/grad_textures"])


def test_reference_driver_render_chain_matches_the_reference_classes(monkeypatch):
    """`ref_gpu.render_mesh_lasr` (the `--impl reference` arm's restatement of SoftRenderer.render_mesh with LASR's
    settings) against what the reference's own classes hand to the rasterizer (tests/golden/modules_ref.npz, case
    'lasr')."""
    import torch
    from oracle import ref_gpu
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modules_ref.npz"))
    rec = {}

    def soft_rasterize(face_vertices, textures, **settings):
        rec.update(fv=face_vertices, tex=textures, settings=settings)
        return torch.from_numpy(g["lasr/rec/render"])
    monkeypatch.setattr(ref_gpu, "soft_rasterize", soft_rasterize)
    eye = torch.tensor([0.0, 0.0, -(1.0 / np.tan(np.radians(30.0)) + 1.0)], dtype=torch.float32)
    ref_gpu.render_mesh_lasr(torch.from_numpy(g["lasr/in/verts"]), torch.from_numpy(g["lasr/in/faces"]),
                             torch.from_numpy(g["lasr/in/textures"]), eye, dict(image_size=8))
    np.testing.assert_allclose(rec["fv"].numpy(), g["lasr/rec/face_vertices"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(rec["tex"].numpy(), g["lasr/rec/face_textures"], rtol=1e-6, atol=1e-7)
