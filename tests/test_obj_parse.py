"""OBJ / MTL parsing of soft_renderer.functional.obj_io on the host (no GPU): hand-computed expectations on a small
polygon soup, and -- where the reference tree exists (the build container) -- the reference's own load_obj /
load_textures (third_party/softras/soft_renderer/functional/load_obj.py) run on the same files, with its CUDA
resampling kernel replaced by a stub that records the arguments it would have been called with."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from lasr_b200.soft_renderer.functional import obj_io

REF = "/root/reference/third_party/softras/soft_renderer/functional/load_obj.py"
SPOT = "/root/reference/database/misc/spot/spot_triangulated.obj"

OBJ = """# a quad, a triangle with normals only, a pentagon, two materials
mtllib soup.mtl
v 0 0 0 1.0 0.5 0.25
v 1 0 0 0.1 0.2 0.3
v 1 1 0 0 0 0
v 0 1 0 1 1 1
v 0.5 1.5 0.5 0.2 0.4 0.6
vt 0.0 0.0
vt 1.0 0.0
vt 1.0 1.0
vt 0.0 1.0
vt 0.5 1.5
usemtl red
f 1/1 2/2 3/3 4/4
usemtl tex
f 1//1 2//1 3//1
f 1/1/1 2/2/1 3/3/1 5/5/1 4/4/1
"""
MTL = "newmtl red\nKd 1.0 0.0 0.0\nnewmtl tex\nKd 0.5 0.5 0.5\nmap_Kd tex.png\n"


@pytest.fixture()
def soup(tmp_path):
    (tmp_path / "soup.obj").write_text(OBJ)
    (tmp_path / "soup.mtl").write_text(MTL)
    return str(tmp_path / "soup.obj")


def test_geometry_and_uv_parsing(soup):
    v, colours, f, mtllib = obj_io.parse_obj(soup)
    assert v.shape == (5, 3) and v.dtype == np.float32 and mtllib == "soup.mtl"
    # quad -> 2 triangles, triangle -> 1, pentagon -> 3 (fans around the first corner), 0-based
    assert f.tolist() == [[0, 1, 2], [0, 2, 3], [0, 1, 2], [0, 1, 2], [0, 2, 4], [0, 4, 3]] and f.dtype == np.int32
    assert np.allclose(np.asarray(colours, np.float32)[0], [1.0, 0.5, 0.25])
    uv, mats = obj_io.parse_obj_uv(soup)
    assert uv.shape == (6, 3, 2) and mats == ["red", "red", "tex", "tex", "tex", "tex"]
    assert np.array_equal(uv[0], [[0, 0], [1, 0], [1, 1]]) and np.array_equal(uv[1], [[0, 0], [1, 1], [0, 1]])
    assert np.array_equal(uv[2], [[0.5, 1.5]] * 3)          # `v//vn` corners: index 0 - 1 = the last vt, as the reference
    colors, files = obj_io.load_mtl(soup[:-4] + ".mtl")
    assert files == {"tex": "tex.png"} and np.allclose(colors["red"], [1, 0, 0])


def _reference_loader(captured):
    """The reference's load_obj module with skimage / the CUDA extension stubbed and `.cuda()` made the identity."""
    from PIL import Image
    sk = types.ModuleType("skimage"); skio = types.ModuleType("skimage.io")
    skio.imread = lambda p: np.asarray(Image.open(p))
    sk.io = skio
    ext = types.ModuleType("soft_renderer.cuda.load_textures")

    def load_textures(image, faces, textures, is_update):
        captured.append((image.clone(), faces.clone(), textures.clone(), is_update.clone()))
        return textures
    ext.load_textures = load_textures
    mods = {"skimage": sk, "skimage.io": skio, "soft_renderer": types.ModuleType("soft_renderer"),
            "soft_renderer.cuda": types.ModuleType("soft_renderer.cuda"), "soft_renderer.cuda.load_textures": ext}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        spec = importlib.util.spec_from_file_location("ref_load_obj", REF)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return m


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("which", ["soup", "spot"])
def test_parsers_agree_with_the_reference_loader(soup, which, monkeypatch, tmp_path):
    from PIL import Image
    path = soup if which == "soup" else SPOT
    if which == "soup":
        Image.fromarray((np.random.RandomState(0).rand(8, 8, 3) * 255).astype(np.uint8)).save(str(tmp_path / "tex.png"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    captured = []
    ref = _reference_loader(captured)
    rv, rf, rt = ref.load_obj(path, load_texture=True, texture_res=2, texture_type="surface")
    v, _, f, mtllib = obj_io.parse_obj(path)
    assert np.array_equal(v, rv.numpy()) and np.array_equal(f, rf.numpy())
    uv, mats = obj_io.parse_obj_uv(path)
    uv = torch.from_numpy(uv); uv[1 < uv] = uv[1 < uv] % 1
    assert len(captured) == 1
    image, faces, textures_in, is_update = captured[0]
    assert torch.equal(uv, faces)                                   # the kernel's `faces` argument
    _, files = obj_io.load_mtl(os.path.join(os.path.dirname(path), mtllib))
    (name, fname), = files.items()
    assert np.array_equal((np.asarray(mats) == name).astype(np.int32), is_update.numpy())
    ours_img = np.ascontiguousarray((obj_io._read_image(os.path.join(os.path.dirname(path), fname)).astype(np.float32) / 255.)[::-1][:, :, :3])
    assert np.array_equal(ours_img, image.numpy())                  # the kernel's `image` argument (flipped, 0..1)
    if which == "soup":                                             # material colours painted before the resampling
        assert torch.equal(textures_in[0], torch.tensor([1.0, 0.0, 0.0]).expand(4, 3))
    vv, ff = ref.load_obj(path, normalization=True)
    monkeypatch.undo()


REF_SAVE = "/root/reference/third_party/softras/soft_renderer/functional/save_obj.py"


@pytest.mark.skipif(not os.path.exists(REF_SAVE), reason="reference tree not present (GPU box)")
def test_writer_produces_the_reference_writers_bytes(tmp_path, monkeypatch):
    """The reference's own save_obj (functional/save_obj.py:39-90, loaded unmodified; skimage and the CUDA module stubbed
    -- neither is reached without 'surface' textures) against obj_io.save_obj: identical .obj bytes with no textures and
    with per-vertex colours.  (The 'surface' branch is checked on the GPU through an export -> import round trip; its
    atlas rows deliberately use floor division, see obj_io.create_texture_image.)"""
    skio = types.ModuleType("skimage.io"); skio.imsave = lambda *a, **k: None
    cuda_pkg = types.ModuleType("soft_renderer.cuda"); cuda_pkg.create_texture_image = types.ModuleType("cti")
    for k, m in {"skimage": types.ModuleType("skimage"), "skimage.io": skio, "soft_renderer": types.ModuleType("soft_renderer"),
                 "soft_renderer.cuda": cuda_pkg, "soft_renderer.cuda.create_texture_image": cuda_pkg.create_texture_image}.items():
        monkeypatch.setitem(sys.modules, k, m)
    spec = importlib.util.spec_from_file_location("ref_save_obj", REF_SAVE)
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    v = torch.randn(11, 3, generator=g) * 3
    f = torch.randint(0, 11, (17, 3), generator=g, dtype=torch.int32)
    c = torch.rand(11, 3, generator=g)
    for tag, kw in (("plain", {}), ("vertex", dict(textures=c, texture_type="vertex"))):
        a, b = str(tmp_path / f"ref_{tag}.obj"), str(tmp_path / f"ours_{tag}.obj")
        ref.save_obj(a, v, f, **kw)
        obj_io.save_obj(b, v, f, **kw)
        ra, rb = open(a).read(), open(b).read()
        assert ra.replace(os.path.basename(a), "X") == rb.replace(os.path.basename(b), "X"), tag
        assert ra.count("\nv ") == 11 and ra.count("\nf ") == 17


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present (GPU box)")
def test_random_obj_files_parse_like_the_reference_loader(tmp_path, monkeypatch):
    """Forty random OBJ files -- polygons of 3 to 6 corners in the four face-token styles (v, v/vt, v//vn, v/vt/vn),
    optional per-vertex colours, blank lines, comments, doubled separators -- through the reference's own load_obj
    (live) and through obj_io: vertices, fan-triangulated faces, per-vertex colours and the normalised variant."""
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    ref = _reference_loader([])
    rng = np.random.RandomState(0)
    for trial in range(40):
        nv, nt = rng.randint(4, 12), rng.randint(3, 8)
        coloured = rng.rand() < 0.4
        lines = ["# fuzz", ""]
        for _ in range(nv):
            lines.append("v %.4f %.4f %.4f" % tuple(rng.randn(3)) + (" %.3f %.3f %.3f" % tuple(rng.rand(3)) if coloured else ""))
            if rng.rand() < 0.2:
                lines.append("")
            if rng.rand() < 0.1:
                lines.append("# comment")
        lines += ["vt %.3f %.3f" % tuple(rng.rand(2) * 1.5) for _ in range(nt)] + ["vn 0 0 1"] * 3
        style = str(rng.choice(["v", "v/vt", "v//vn", "v/vt/vn"]))
        for _ in range(rng.randint(2, 7)):
            idx = rng.choice(nv, min(rng.randint(3, 7), nv), replace=False) + 1
            tok = {"v": lambda j: str(j), "v/vt": lambda j: f"{j}/{rng.randint(1, nt + 1)}", "v//vn": lambda j: f"{j}//1",
                   "v/vt/vn": lambda j: f"{j}/{rng.randint(1, nt + 1)}/2"}[style]
            sep = "  " if rng.rand() < 0.3 else " "
            lines.append("f" + sep + sep.join(tok(j) for j in idx))
        p = str(tmp_path / f"m{trial}.obj")
        with open(p, "w") as fh:
            fh.write("\n".join(lines) + "\n")
        rv, rf = ref.load_obj(p)
        v, _, f, _ = obj_io.parse_obj(p)
        assert np.array_equal(v, rv.numpy()) and np.array_equal(f, rf.numpy()), (trial, style)
        if coloured:
            assert torch.equal(obj_io.load_obj(p, load_texture=True, texture_type="vertex")[2],
                               ref.load_obj(p, load_texture=True, texture_type="vertex")[2]), trial
        assert torch.allclose(obj_io.load_obj(p, normalization=True)[0], ref.load_obj(p, normalization=True)[0], atol=1e-6), trial
