"""Chamfer nearest neighbours and the texture-atlas kernels.

CPU: the numpy oracles against independent restatements and, once minted, the golden vectors produced by the
reference's own kernels (tests/golden/{chamfer,texture_io}_ref.npz).  GPU: the CUDA kernels through the C ABI /
the reference-shaped Python mirrors against the oracles (indices and texel selections bit-exact, distances
bit-exact, gradients to float rounding) and against the golden vectors.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aux_cases  # noqa: E402
from oracle import chamfer as OC  # noqa: E402
from oracle import texture_io as OT  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _golden(name):
    p = os.path.join(GOLD, name)
    return np.load(p) if os.path.exists(p) else None


# ------------------------------------------------------------------------------------------------ CPU
def test_chamfer_oracle_matches_float64_brute_force():
    for name, (a, b, g1, g2) in aux_cases.chamfer_cases().items():
        d1, d2, i1, i2 = OC.chamfer_forward(a, b)
        for bi in range(a.shape[0]):
            D = ((a[bi][:, None, :].astype(np.float64) - b[bi][None, :, :].astype(np.float64)) ** 2).sum(-1)
            np.testing.assert_allclose(d1[bi], D.min(1), rtol=2e-6, atol=1e-12, err_msg=name)
            np.testing.assert_allclose(d2[bi], D.min(0), rtol=2e-6, atol=1e-12, err_msg=name)
            # the chosen neighbour is a minimiser (to float32 rounding of the distance)
            assert (D[np.arange(D.shape[0]), i1[bi]] <= D.min(1) * (1 + 1e-5) + 1e-12).all(), name
        if name == "ties":      # integer lattice: distances exact, so the lowest index must win
            D = ((a[0][:, None, :] - b[0][None, :, :]) ** 2).sum(-1)
            assert np.array_equal(i1[0], D.argmin(1)) and np.array_equal(i2[0], D.argmin(0))


def test_chamfer_oracle_gradient_is_the_derivative():
    a, b, g1, g2 = aux_cases.chamfer_cases()["small"]
    d1, d2, i1, i2 = OC.chamfer_forward(a, b)
    ga, gb = OC.chamfer_backward(a, b, g1, g2, i1, i2)
    ta = torch.tensor(a, dtype=torch.float64, requires_grad=True); tb = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    D = ((ta[:, :, None, :] - tb[:, None, :, :]) ** 2).sum(-1)
    loss = (D.min(2)[0] * torch.tensor(g1, dtype=torch.float64)).sum() + (D.min(1)[0] * torch.tensor(g2, dtype=torch.float64)).sum()
    loss.backward()
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_chamfer_oracle_matches_reference_kernels_golden():
    g = _golden("chamfer_ref.npz")
    if g is None:
        pytest.skip("tests/golden/chamfer_ref.npz not minted yet (tests/golden/make_aux_golden.py on a GPU box)")
    for name, (a, b, g1, g2) in aux_cases.chamfer_cases().items():
        d1, d2, i1, i2 = OC.chamfer_forward(a, b)
        assert np.array_equal(i1, g[f"{name}/idx1"]) and np.array_equal(i2, g[f"{name}/idx2"]), name
        assert np.array_equal(d1, g[f"{name}/dist1"]) and np.array_equal(d2, g[f"{name}/dist2"]), name   # bit-exact
        ga, gb = OC.chamfer_backward(a, b, g1, g2, i1, i2)
        np.testing.assert_allclose(ga, g[f"{name}/grad1"], rtol=2e-5, atol=2e-6, err_msg=name)   # float atomics
        np.testing.assert_allclose(gb, g[f"{name}/grad2"], rtol=2e-5, atol=2e-6, err_msg=name)


def test_load_textures_oracle_is_bilinear_sampling():
    image, faces, textures, is_update = aux_cases.load_textures_cases()["r4"]
    out = OT.load_textures(image, faces, textures, is_update)
    assert np.array_equal(out[is_update == 0], textures[is_update == 0])      # untouched faces
    # independent float64 restatement of the sampling positions
    R = 4; H, W = image.shape[:2]
    for fn in np.where(is_update)[0][:5]:
        for t in range(R * R):
            wy, wx = t // R, t % R
            if wx + wy < R:
                w0, w1 = (wx + 1 / 3) / R, (wy + 1 / 3) / R
            else:
                w0, w1 = ((R - 1 - wx) + 2 / 3) / R, ((R - 1 - wy) + 2 / 3) / R
            w = np.array([w0, w1, 1 - w0 - w1])
            px, py = (faces[fn, :, 0] * w).sum() * (W - 1), (faces[fn, :, 1] * w).sum() * (H - 1)
            x0, y0 = int(px), int(py); fx, fy = px - x0, py - y0
            ref = (image[y0, x0] * (1 - fx) * (1 - fy) + image[y0 + 1, x0] * (1 - fx) * fy
                   + image[y0, x0 + 1] * fx * (1 - fy) + image[y0 + 1, x0 + 1] * fx * fy)
            np.testing.assert_allclose(out[fn, t], ref, atol=5e-5)


def test_create_texture_image_oracle_paints_every_cell():
    faces, textures, image = aux_cases.create_texture_image_cases()["f10"]
    out = OT.create_texture_image(faces, textures, image)
    res = 8; tile_width = 4
    for fn in range(10):
        r, c = fn // tile_width, fn % tile_width
        cell = out[r * res:(r + 1) * res, c * res:(c + 1) * res].reshape(-1, 3)
        tex = textures[fn]
        assert all(any(np.array_equal(px, t) for t in tex) for px in cell), fn     # every pixel is one of the face's texels
    assert np.array_equal(out[2 * res:, 2 * res:], image[2 * res:, 2 * res:])      # cells beyond the last face untouched


def test_texture_oracles_match_reference_kernels_golden():
    g = _golden("texture_io_ref.npz")
    if g is None:
        pytest.skip("tests/golden/texture_io_ref.npz not minted yet (tests/golden/make_aux_golden.py on a GPU box)")
    for name, (image, faces, textures, is_update) in aux_cases.load_textures_cases().items():
        assert np.array_equal(OT.load_textures(image, faces, textures, is_update), g[f"load/{name}"]), name
    for name, (faces, textures, image) in aux_cases.create_texture_image_cases().items():
        assert np.array_equal(OT.create_texture_image(faces, textures, image), g[f"create/{name}"]), name


def test_obj_round_trip_parsers_need_no_gpu(tmp_path):
    from lasr_b200.soft_renderer.functional import obj_io
    mtl = tmp_path / "m.mtl"
    mtl.write_text("newmtl a\nKd 0.1 0.2 0.3\nmap_Kd tex.png\nnewmtl b\nKd 1 0 0\n")
    colors, files = obj_io.load_mtl(str(mtl))
    assert files == {"a": "tex.png"} and np.allclose(colors["a"], [0.1, 0.2, 0.3]) and np.allclose(colors["b"], [1, 0, 0])
    assert obj_io._uv_index("3/7/2") == 7 and obj_io._uv_index("3//2") == 0 and obj_io._uv_index("3") == 0


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_chamfer_kernels_match_oracle_and_golden():
    from lasr_b200.chamfer3D import dist_chamfer_3D as C
    g = _golden("chamfer_ref.npz")
    dev = torch.device("cuda", 0)
    mod = C.chamfer_3DDist()
    for name, (a, b, g1, g2) in aux_cases.chamfer_cases().items():
        A = torch.from_numpy(a).to(dev).requires_grad_(True); Bt = torch.from_numpy(b).to(dev).requires_grad_(True)
        d1, d2, i1, i2 = mod(A, Bt)
        assert i1.dtype == torch.int32 and i2.dtype == torch.int32
        (d1 * torch.from_numpy(g1).to(dev)).sum().add((d2 * torch.from_numpy(g2).to(dev)).sum()).backward()
        od1, od2, oi1, oi2 = OC.chamfer_forward(a, b)
        assert np.array_equal(i1.cpu().numpy(), oi1) and np.array_equal(i2.cpu().numpy(), oi2), name
        assert np.array_equal(d1.detach().cpu().numpy(), od1) and np.array_equal(d2.detach().cpu().numpy(), od2), name
        oga, ogb = OC.chamfer_backward(a, b, g1, g2, oi1, oi2)
        np.testing.assert_allclose(A.grad.cpu().numpy(), oga, rtol=2e-5, atol=2e-6, err_msg=name)
        np.testing.assert_allclose(Bt.grad.cpu().numpy(), ogb, rtol=2e-5, atol=2e-6, err_msg=name)
        if g is not None:
            assert np.array_equal(i1.cpu().numpy(), g[f"{name}/idx1"]) and np.array_equal(d1.detach().cpu().numpy(), g[f"{name}/dist1"])
            assert np.array_equal(i2.cpu().numpy(), g[f"{name}/idx2"]) and np.array_equal(d2.detach().cpu().numpy(), g[f"{name}/dist2"])


@pytest.mark.gpu
def test_chamfer_lasr_call_site_and_large_clouds():
    """nnutils/mesh_net.py:474-478: nearest mirrored vertex, indices only; plus a cloud big enough for the
    two-queries-per-thread instantiation, checked through idempotence properties."""
    from lasr_b200.chamfer3D import dist_chamfer_3D as C
    dev = torch.device("cuda", 0)
    a, b, _, _ = aux_cases.chamfer_cases()["mirror"]
    _, _, idx1, _ = C.chamfer_3DDist()(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
    tex = torch.rand(a.shape[1], 3, device=dev)
    assert tex[idx1[0].long()].shape == tex.shape
    g = torch.Generator(device="cpu").manual_seed(5)
    P = torch.rand(1, 80000, 3, generator=g).to(dev)
    perm = torch.randperm(80000, generator=g).to(dev)
    d1, d2, i1, i2 = C.chamfer_3DDist()(P, P[:, perm])
    assert float(d1.max()) == 0.0 and float(d2.max()) == 0.0                  # a cloud against its own permutation
    assert torch.equal(P[0, perm][i1[0].long()], P[0])                        # the match is the same point
    Q = torch.rand(1, 70000, 3, generator=g).to(dev)
    d1, d2, i1, i2 = C.chamfer_3DDist()(P, Q)
    chk = ((P[0, :512, None, :] - Q[0, None, :, :]) ** 2).sum(-1).min(1)
    assert torch.allclose(d1[0, :512], chk[0], rtol=1e-5, atol=1e-9)
    with pytest.raises((TypeError, RuntimeError)):
        C.chamfer_3DDist()(torch.rand(1, 4, 3), torch.rand(1, 4, 3))           # CPU tensors are refused loudly


@pytest.mark.gpu
def test_texture_kernels_match_oracle_and_golden():
    from lasr_b200 import _cabi
    g = _golden("texture_io_ref.npz")
    dev = torch.device("cuda", 0)
    for name, (image, faces, textures, is_update) in aux_cases.load_textures_cases().items():
        t = torch.from_numpy(textures).to(dev).clone()
        H, W = image.shape[:2]
        _cabi.native_call("load_textures", torch.from_numpy(image).to(dev), torch.from_numpy(faces).to(dev),
                          torch.from_numpy(is_update).to(dev), t, faces.shape[0], int(np.sqrt(textures.shape[1])), H, W)
        got = t.cpu().numpy()
        assert np.array_equal(got, OT.load_textures(image, faces, textures, is_update)), name
        if g is not None:
            assert np.array_equal(got, g[f"load/{name}"]), name
    for name, (faces, textures, image) in aux_cases.create_texture_image_cases().items():
        im = torch.from_numpy(image).to(dev).clone()
        _cabi.native_call("create_texture_image", torch.from_numpy(faces).to(dev), torch.from_numpy(textures).to(dev), im,
                          faces.shape[0], int(np.sqrt(textures.shape[1])), image.shape[0], image.shape[1], 1e-5)
        got = im.cpu().numpy()
        assert np.array_equal(got, OT.create_texture_image(faces, textures, image)), name
        if g is not None:
            assert np.array_equal(got, g[f"create/{name}"]), name


@pytest.mark.gpu
def test_obj_export_import_round_trip(tmp_path):
    """save_obj -> load_obj through both texture kernels: geometry exact, texel colours preserved to 8-bit PNG."""
    import lasr_b200.soft_renderer as sr
    spot = np.load(os.path.join(GOLD, "spot_mesh.npz"))
    v = torch.from_numpy(spot["vertices"][:300].astype(np.float32)).cuda()
    f = torch.from_numpy((spot["faces"] % 300)[:200].astype(np.int32)).cuda()
    R = 4
    tex = torch.rand(200, 1, 3).repeat(1, R * R, 1).cuda()                    # one flat colour per face
    mesh = sr.Mesh(v, f, tex, texture_res=R, texture_type="surface")
    path = str(tmp_path / "m.obj")
    mesh.save_obj(path, save_texture=True, texture_res_out=16)
    assert os.path.exists(path) and os.path.exists(path[:-4] + ".png") and os.path.exists(path[:-4] + ".mtl")
    back = sr.Mesh.from_obj(path, load_texture=True, texture_res=R, texture_type="surface")
    assert torch.allclose(back.vertices[0], v, atol=1e-6) and torch.equal(back.faces[0], f)
    assert (back.textures[0] - tex).abs().max() < 3.0 / 255
    v2, f2 = sr.functional.load_obj(path)
    assert v2.shape == v.shape and f2.dtype == torch.int32
