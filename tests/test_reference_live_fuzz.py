"""Randomised comparison of the host-side mirrors with the reference's own Python, imported live from /root/reference
(build container only; skipped where the tree is absent).  The golden vectors pin fixed cases that travel; this sweeps
shapes and arguments the goldens do not: hypothesis / bone / vertex counts, tocam on and off, per-frame eyes, list and
tensor arguments, big- and little-endian colour and grey PFM files."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "nnutils", "geom_utils.py")),
                                reason="reference tree not present (GPU box)")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _close(tag, a, b, tol=1e-6):
    d = float((a - b).abs().max())
    assert d <= tol * max(1.0, float(b.abs().max())), (tag, d)


def test_geometry_mirrors_follow_the_reference_on_random_shapes(monkeypatch):
    import warnings
    from lasr_b200 import geom_utils as gu
    import lasr_b200.soft_renderer.functional as srf
    import lasr_b200.soft_renderer as sr_pkg
    monkeypatch.setitem(sys.modules, "soft_renderer", types.ModuleType("soft_renderer"))      # nnutils/geom_utils.py:24
    rg = _load(os.path.join(REF, "nnutils/geom_utils.py"), "ref_geom_live")
    fdir = os.path.join(REF, "third_party/softras/soft_renderer/functional")
    R = {n: getattr(_load(os.path.join(fdir, n + ".py"), "ref_live_" + n), n)
         for n in ("look_at", "orthogonal", "perspective", "face_vertices", "vertex_normals", "ambient_lighting",
                   "directional_lighting", "get_points_from_angles", "projection")}
    g = torch.Generator().manual_seed(0)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                       # the reference calls torch.cross without dim
        for _ in range(25):
            bs, H, B, V = ri(1, 4), ri(1, 3), ri(1, 5), ri(5, 40)
            N = 2 * bs * H
            verts = torch.randn(N, V, 3, generator=g)
            Rm = gu.quaternion_to_rotation_matrix(torch.randn(N * B, 4, generator=g))
            Tm = torch.randn(N * B, 1, 3, generator=g) + torch.tensor([0, 0, 8.0])
            skin = torch.softmax(torch.randn(N, B - 1, V, 1, generator=g), 1) if B > 1 else None
            for tocam in (True, False):
                _close(("obj_to_cam", N, B, V, tocam), gu.obj_to_cam(verts, Rm, Tm, B, H, skin, tocam=tocam),
                       rg.obj_to_cam(verts, Rm, Tm, B, H, skin, tocam=tocam))
            v4 = torch.cat([verts + torch.tensor([0, 0, 6.0]), torch.ones(N, V, 1)], -1)
            pp, fl = torch.randn(2 * bs, 2, generator=g) * 0.1, 8 + torch.rand(2 * bs, H, generator=g)
            _close("pinhole_cam", gu.pinhole_cam(v4, pp, fl), rg.pinhole_cam(v4.clone(), pp, fl))
            _close("orthographic_cam", gu.orthographic_cam(v4, pp, fl), rg.orthographic_cam(v4.clone(), pp, fl))
            nb, nf = ri(1, 6), ri(4, 30)
            nb, nf = (4 if nb == 3 else nb), (5 if nf == 3 else nf)       # the reference's dim-less torch.cross
            vv = torch.randn(nb, V, 3, generator=g)
            ff = torch.randint(0, V, (nb, nf, 3), generator=g, dtype=torch.int32)
            _close("face_vertices", srf.face_vertices(vv, ff), R["face_vertices"](vv, ff))
            _close("vertex_normals", srf.vertex_normals(vv, ff), R["vertex_normals"](vv, ff), 1e-5)
            eye = [float(x) for x in torch.randn(3, generator=g) + torch.tensor([0, 0, -3.0])]
            _close("look_at (list eye)", srf.look_at(vv, eye), R["look_at"](vv, eye), 1e-5)
            eyes = torch.randn(nb, 3, generator=g) + torch.tensor([0, 0, -3.0])
            _close("look_at (per-frame eyes)", srf.look_at(vv, eyes), R["look_at"](vv, eyes), 1e-5)
            far = vv + torch.tensor([0, 0, 5.0])
            _close("perspective", srf.perspective(far, angle=37.0), R["perspective"](far, angle=37.0))
            _close("orthogonal", srf.orthogonal(vv, 1.3), R["orthogonal"](vv, 1.3))
            light = torch.zeros(nb, V, 3)
            nrm = torch.nn.functional.normalize(torch.randn(nb, V, 3, generator=g), dim=2)
            _close("ambient_lighting", srf.ambient_lighting(light.clone(), 0.4, [1, 0.5, 0.2]),
                   R["ambient_lighting"](light.clone(), 0.4, [1, 0.5, 0.2]))
            _close("directional_lighting", srf.directional_lighting(light.clone(), nrm, 0.7, [1, 0.5, 0.2], [0.3, 0.8, -0.5]),
                   R["directional_lighting"](light.clone(), nrm, 0.7, [1, 0.5, 0.2], [0.3, 0.8, -0.5]))
            d, e, a = torch.rand(nb, generator=g) + 2, torch.rand(nb, generator=g) * 60 - 30, torch.rand(nb, generator=g) * 360
            _close("get_points_from_angles", srf.get_points_from_angles(d, e, a), R["get_points_from_angles"](d, e, a))
            Pm = torch.randn(nb, 3, 4, generator=g) * 0.05 + torch.tensor([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 6]])
            dc = torch.randn(nb, 5, generator=g) * 0.05
            _close("projection", srf.projection(vv, Pm, dc, 256), R["projection"](vv, Pm, dc, 256), 1e-6)
            cam = sr_pkg.Transform(camera_mode="projection", P=Pm, dist_coeffs=dc, orig_size=256)
            _close("Projection module", cam.transformer(vv), R["projection"](vv, Pm, dc, 256), 1e-6)
        assert srf.get_points_from_angles(2.7, 20.0, 40.0) == pytest.approx(R["get_points_from_angles"](2.7, 20.0, 40.0))


def test_pfm_reader_follows_the_reference_reader(tmp_path, monkeypatch):
    from lasr_b200.dataloader.pfm import readPFM
    monkeypatch.setitem(sys.modules, "png", types.ModuleType("png"))          # pypng: imported by util_flow, unused here
    uf = _load(os.path.join(REF, "third_party/ext_utils/util_flow.py"), "ref_util_flow_live")
    rng = np.random.RandomState(0)
    for i, (shape, little) in enumerate([((7, 9, 3), True), ((7, 9, 3), False), ((5, 4), True), ((5, 4), False),
                                         ((1, 1, 3), True), ((33, 2), True)]):
        data = rng.randn(*shape).astype(np.float32)
        p = str(tmp_path / f"f{i}.pfm")
        with open(p, "wb") as fh:
            fh.write(b"PF\n" if data.ndim == 3 else b"Pf\n")
            fh.write(f"{shape[1]} {shape[0]}\n".encode())
            fh.write(b"-1.0\n" if little else b"1.0\n")
            np.flipud(data).astype("<f4" if little else ">f4").tofile(fh)
        a, sa = readPFM(p)
        b, sb = uf.readPFM(p)
        assert np.array_equal(a, b) and sa == sb and a.shape == b.shape, (shape, little)
        assert np.array_equal(a, data)


def test_image_losses_follow_the_reference_lines_on_random_inputs(monkeypatch):
    """nnutils/mesh_net.py:361-363,372-447 executed live (as tests/golden/make_losses_golden.py does for the two stored
    cases) on random batch sizes, hypothesis counts and image sizes, with the degenerate situations the stored cases only
    touch once: all-background flow masks, frames whose occlusion map is entirely zero, empty silhouettes."""
    import textwrap
    from lasr_b200 import losses
    with open(os.path.join(REF, "nnutils/mesh_net.py")) as fh:
        lines = fh.readlines()
    prep, body = textwrap.dedent("".join(lines[360:363])), textwrap.dedent("".join(lines[371:447]))
    assert "img_obs = self.imgs" in prep and "# 1) mask loss" in body and "self.total_loss += self.texture_loss" in body
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    g = torch.Generator().manual_seed(3)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    for trial in range(12):
        bs, H, S = ri(1, 4), ri(1, 4), int([8, 12, 20][ri(0, 3)])
        n2 = 2 * bs
        U = lambda *sh: torch.rand(*sh, generator=g)
        masks = (U(n2, S, S) > 0.45).float()
        occ = torch.randn(n2, S, S, generator=g)
        occ[U(n2, S, S) < 0.2] = 0
        imgs, flow = U(n2, 3, S, S), torch.randn(n2, 3, S, S, generator=g) * 0.05
        render = U(n2 * H, 4, S, S)
        flow_rd = torch.randn(n2 * H, S, S, 2, generator=g) * 0.05
        bgmask = U(n2 * H, S, S) < 0.4
        if trial % 3 == 0:
            bgmask[ri(0, n2 * H)] = True                     # a (frame, hypothesis) whose flow mask is empty
        if trial % 4 == 1:
            masks[ri(0, n2)] = 0                             # an empty observed silhouette
        if trial % 5 == 2:
            occ[ri(0, n2)] = 0                               # no valid pixel at all: the reference's masked mean is NaN
        self = types.SimpleNamespace(imgs=imgs, masks=masks, occ=occ, flow=flow, flow_rd=flow_rd, bgmask=bgmask,
                                     texture_render=render[:, :3], fgmask_tex=render[:, -1], mask_pred=render[:, -1], iters=1,
                                     ptex_loss=types.SimpleNamespace(forward_pair=lambda a, b: torch.zeros(a.shape[0])))
        opts = types.SimpleNamespace(n_hypo=H, img_size=S, opt_tex="yes", l1tex_wt=0.7, n_bones=1)
        ns = dict(self=self, opts=opts, torch=torch, local_batch_size=bs, np=np)
        exec(prep, ns)
        exec(body, ns)
        mp, tex = render[:, 3].reshape(n2, H, S, S), render[:, :3].reshape(n2, H, 3, S, S)
        l1, s1 = losses.mask_loss(mp, masks, occ)
        l2, s2, _ = losses.flow_loss(flow_rd.reshape(n2, H, S, S, 2), bgmask.reshape(n2, H, S, S), flow, occ, masks)
        l3, s3 = losses.texture_l1_loss(tex, mp, imgs, masks, occ, 0.7)
        what = (trial, bs, H, S)
        for ours, ref in ((s1, self.mask_loss_sub), (s2, self.flow_rd_loss_sub), (s3, self.texture_loss_sub)):
            np.testing.assert_allclose(ours.numpy(), ref.detach().numpy(), rtol=3e-5, atol=1e-7, err_msg=str(what))
        tot, ref_tot = float(l1 + l2 + l3), float(self.total_loss)
        assert (np.isnan(tot) and np.isnan(ref_tot)) or abs(tot - ref_tot) < 3e-6, what
        assert np.isnan(ref_tot) == bool((occ.reshape(n2, -1) == 0).all(1).any()), what


@pytest.mark.parametrize("bs,H,B,S,seed", [(2, 1, 1, 12, 10), (2, 2, 3, 10, 11), (3, 1, 2, 10, 12), (2, 1, 4, 12, 13)])
def test_whole_region_follows_the_reference_lines_for_larger_batches(monkeypatch, bs, H, B, S, seed):
    """The stored whole-region vectors (region_losses_ref.npz) all have one frame pair per batch.  Here the reference's
    own LASR.forward lines (nnutils/mesh_net.py:259-447, on top of the drop-in package, oracle as the rasterizer) are
    executed live for batches of two and three pairs, with and without bones and hypotheses, and RenderHotPath must
    reproduce their renders and losses."""
    import importlib.util as iu
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    spec = iu.spec_from_file_location("make_region_losses_golden_live", os.path.join(here, "golden", "make_region_losses_golden.py"))
    gen = iu.module_from_spec(spec); spec.loader.exec_module(gen)
    import lasr_b200.soft_renderer.rasterizer as rast
    from lasr_b200 import hotpath
    monkeypatch.setattr(torch.Tensor, "cuda", torch.Tensor.cuda)                    # setup() patches these two globally:
    monkeypatch.setattr(rast.srf, "soft_rasterize", rast.srf.soft_rasterize)        # restored when the test ends
    monkeypatch.setitem(sys.modules, "soft_renderer", sys.modules.get("soft_renderer", types.ModuleType("soft_renderer")))
    env = gen.setup()
    ins, want = gen.run_case(env, bs, H, B, S, seed)
    I = lambda k: torch.from_numpy(ins[k])
    bones = {k: I(k) for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")} if B > 1 else None
    Rmat = I("quat").view(-1, 3, 3).permute(0, 2, 1).contiguous()                   # mesh_net.py:259
    Tmat = torch.cat([I("trans"), I("depth")], 1)                                   # :260
    batch = dict(masks=I("masks"), imgs=I("imgs"), flow=I("flow"), occ=I("occ"))
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B, l1tex_wt=1.0))
    out = m(I("pred_v"), I("tex"), I("faces"), Rmat, Tmat, I("ppoint"), I("scale"), batch, bones)
    np.testing.assert_allclose(out["mask_pred"].detach().numpy(), want["mask_pred"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["flow_rd"].detach().numpy(), want["flow_rd"], rtol=0, atol=5e-6)
    for ours, ref in (("mask_loss", "mask_loss"), ("flow_loss", "flow_rd_loss"), ("texture_loss", "texture_loss"), ("total", "total_loss")):
        assert abs(float(out[ours]) - float(want[ref])) < 2e-6, (ours, float(out[ours]), float(want[ref]))


def test_dataloader_oracle_follows_the_reference_dataset_on_random_videos(tmp_path, monkeypatch):
    """The reference's own BaseDataset.__getitem__ (dataloader/vidbase.py, imported live) on random synthetic videos --
    frame sizes that are not multiples of anything, objects of any size anywhere (crops leaving the frame on every
    side) -- against oracle/dataloader.py on the same decoded inputs: bit for bit at power-of-two img_size, float32
    rounding at img_size 100 (see the oracle's header)."""
    cv2 = pytest.importorskip("cv2")
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    from oracle import dataloader as OD
    gen = _load(os.path.join(here, "golden", "make_dataloader_golden.py"), "make_dataloader_golden_live")
    for name in ("absl", "absl.flags", "absl.app", "png"):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules["absl"].flags, sys.modules["absl"].app = sys.modules["absl.flags"], sys.modules["absl.app"]
    monkeypatch.syspath_prepend(os.path.join(REF, "third_party"))
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "dataloader" or k.startswith(("dataloader.", "ext_utils"))]:
        monkeypatch.delitem(sys.modules, k)
    from ext_utils import image as image_utils
    monkeypatch.setattr(image_utils, "sample_contour", lambda m: np.zeros((1000, 2)))     # skimage tracer: not on this path
    from ext_utils.util_flow import readPFM
    from dataloader import vidbase
    rng = np.random.RandomState(123)
    keys = ("img", "imgn", "mask", "flow", "flown", "occ", "occn", "cam", "camn", "pps")
    checked = 0
    for trial in range(8):
        H, W = int(rng.choice([48, 63, 80, 97])), int(rng.choice([64, 91, 128, 150]))
        S = [32, 64, 100][trial % 3]
        kw = dict(seed=100 + trial, H=H, W=W, centre=(float(rng.uniform(0, W)), float(rng.uniform(0, H))),
                  step=(float(rng.uniform(-6, 6)), float(rng.uniform(-4, 4))),
                  axes=(float(rng.uniform(3, 0.45 * W)), float(rng.uniform(3, 0.45 * H))))
        d = tmp_path / f"v{trial}"
        d.mkdir()
        L = gen.synth_video(str(d), **kw)
        masks = np.stack([cv2.imread(p, 0) for p in L["mask"]])
        if not all((m > 0).any() for m in masks):
            continue                                             # object left the frame: the reference raises on it
        ds = vidbase.BaseDataset(types.SimpleNamespace(img_size=S))
        ds.imglist, ds.masklist, ds.flowfwlist, ds.flowbwlist, ds.camlist = L["img"], L["mask"], L["fw"], L["bw"], L["cam"]
        ds.dframe, ds.can_frame, ds.baselist, ds.directlist, ds.num_imgs = 1, 0, [0, 1, 1, 2], [1, 1, 0, 0], 4
        gold = {"frames_bgr": np.stack([cv2.imread(p) for p in L["img"]]), "masks": masks,
                "baselist": np.asarray(ds.baselist), "directlist": np.asarray(ds.directlist)}
        for dd in ("fw", "bw"):
            gold[f"flow_{dd}"] = np.stack([np.ascontiguousarray(readPFM(p)[0]) for p in L[dd]])
            gold[f"occ_{dd}"] = np.stack([np.ascontiguousarray(readPFM(p.replace("flo-", "occ-"))[0]) for p in L[dd]])
        for it in (0, 1, 3):
            elem = ds[it]
            r = OD.preprocess_pair(*OD.pair_inputs(gold, it, ""), S)
            for k in keys:
                ref, got = np.asarray(elem[k]), r[k]
                assert got.shape == ref.shape, (trial, it, k)
                if S != 100 or k in ("mask", "cam", "camn", "pps"):
                    tol = 1e-12 if k in ("img", "imgn") else 0.0
                    assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() <= tol, (trial, (H, W, S), it, k)
                else:
                    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=5e-6 * max(1.0, float(np.abs(ref).max())),
                                               err_msg=str((trial, it, k)))
            checked += 1
    assert checked >= 15


def test_module_layer_follows_the_reference_classes_on_random_configurations(monkeypatch):
    """Forty random SoftRenderer configurations -- camera (perspective / orthographic, viewing angle and scale, eyes as
    lists, numpy arrays, per-frame tensors, from angles), lighting (mode, intensities, colours, direction), rasterizer
    scalars and modes, anti-aliasing, fill_back, texture type and resolution, default textures -- through the
    reference's own classes (loaded live, recording rasterizer) and through the mirror: same face vertices, face
    textures, scalars and pooled output.  tests/golden/modules_ref.npz stores eight fixed ones."""
    import warnings
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    gen = _load(os.path.join(here, "golden", "make_modules_golden.py"), "make_modules_golden_live")
    import lasr_b200.soft_renderer as ours_sr
    import lasr_b200.soft_renderer.rasterizer as rast
    from lasr_b200 import synth
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(rast.srf, "soft_rasterize", rast.srf.soft_rasterize)            # restored at teardown
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "soft_renderer" or k.startswith("soft_renderer.")}
    try:
        srf = types.ModuleType("soft_renderer.functional")
        fdir = os.path.join(gen.SR, "functional")
        for fn in ("look_at", "look", "perspective", "orthogonal", "face_vertices", "vertex_normals", "ambient_lighting",
                   "directional_lighting", "get_points_from_angles"):
            setattr(srf, fn, getattr(gen.load(os.path.join(fdir, fn + ".py"), "ref_fn_live_" + fn), fn))
        sr = types.ModuleType("soft_renderer")
        sr.functional = srf
        sys.modules["soft_renderer"], sys.modules["soft_renderer.functional"] = sr, srf
        for mod, names in (("mesh", ("Mesh",)), ("lighting", ("AmbientLighting", "DirectionalLighting", "Lighting")),
                           ("transform", ("Projection", "LookAt", "Look", "Transform")), ("rasterizer", ("SoftRasterizer",)),
                           ("renderer", ("SoftRenderer",))):
            m = gen.load(os.path.join(gen.SR, mod + ".py"), "soft_renderer." + mod)
            for n in names:
                setattr(sr, n, getattr(m, n))
        rng = np.random.RandomState(0)
        v, f = synth.icosphere(1, 0.6)
        U = lambda n: [float(x) for x in rng.rand(n)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for trial in range(40):
                N = int(rng.choice([1, 2, 4, 5]))                 # not 3: the reference's dim-less torch.cross
                g = torch.Generator().manual_seed(trial)
                ttype, tres = str(rng.choice(["surface", "vertex"])), int(rng.choice([1, 2, 3]))
                kw = dict(image_size=int(rng.choice([4, 8])), camera_mode="look_at", perspective=bool(rng.rand() < 0.5),
                          viewing_angle=float(rng.uniform(15, 60)), viewing_scale=float(rng.uniform(0.5, 2)),
                          light_mode=str(rng.choice(["surface", "vertex"])), light_intensity_ambient=float(rng.uniform(0, 1)),
                          light_color_ambient=U(3), light_intensity_directionals=float(rng.uniform(0, 1)),
                          light_color_directionals=U(3), light_directions=[float(x) for x in rng.randn(3)],
                          anti_aliasing=bool(rng.rand() < 0.3), fill_back=bool(rng.rand() < 0.5), background_color=U(3),
                          near=float(rng.uniform(0.1, 2)), far=float(rng.uniform(20, 100)),
                          sigma_val=float(10 ** rng.uniform(-6, -3)), gamma_val=float(10 ** rng.uniform(-4, -1)),
                          dist_func=str(rng.choice(["hard", "euclidean", "barycentric"])),
                          aggr_func_rgb=str(rng.choice(["hard", "softmax"])), aggr_func_alpha=str(rng.choice(["hard", "prod", "sum"])))
                eye_kind = str(rng.choice(["none", "list", "np", "tensorN", "angles_f", "angles_t"]))
                if eye_kind == "list":
                    kw["eye"] = [float(x) for x in rng.randn(3) + [0, 0, -3]]
                verts = torch.from_numpy(v)[None].repeat(N, 1, 1) + torch.randn(N, len(v), 3, generator=g) * 0.03
                faces = torch.from_numpy(f)[None].repeat(N, 1, 1)
                tex = None
                if rng.rand() < 0.8:
                    tex = torch.rand(N, len(f), tres ** 2, 3, generator=g) if ttype == "surface" else torch.rand(N, len(v), 3, generator=g)
                recs = []
                for pkg, target in ((sr, srf), (ours_sr, rast.srf)):
                    rec = {}

                    def soft_rasterize(face_vertices, textures_, *args, _rec=rec):
                        _rec["fv"], _rec["ft"], _rec["args"] = face_vertices.clone(), textures_.clone(), args
                        S = args[0]
                        n = face_vertices.shape[0]
                        return torch.arange(n * 4 * S * S, dtype=torch.float32).reshape(n, 4, S, S) / 100
                    target.soft_rasterize = soft_rasterize
                    r = pkg.SoftRenderer(**kw)
                    if eye_kind == "np":
                        r.transform.set_eyes(np.asarray([0.2, -0.1, -2.5], np.float32))
                    elif eye_kind == "tensorN":
                        r.transform.set_eyes(torch.tensor([[0.2, -0.1, -2.5]]).repeat(N, 1) + torch.arange(N)[:, None] * 0.1)
                    elif eye_kind == "angles_f":
                        r.transform.set_eyes_from_angles(2.5, 15.0, 70.0)
                    elif eye_kind == "angles_t":
                        r.transform.set_eyes_from_angles(torch.full((N,), 2.5), torch.linspace(-20, 20, N), torch.linspace(0, 300, N))
                    mesh = pkg.Mesh(verts.clone(), faces.clone(), textures=None if tex is None else tex.clone(),
                                    texture_res=tres, texture_type=ttype)
                    rec["img"] = r.render_mesh(mesh)
                    recs.append(rec)
                a, b = recs
                what = (trial, eye_kind, N, ttype, kw["light_mode"])
                assert torch.allclose(a["fv"], b["fv"], atol=1e-5), what
                assert torch.allclose(a["ft"], b["ft"], atol=1e-6), what
                assert torch.equal(a["img"], b["img"]) and len(a["args"]) == len(b["args"]) == 13, what
                assert all(x == y for x, y in zip(a["args"], b["args"])), what
    finally:
        for k in [k for k in sys.modules if k == "soft_renderer" or k.startswith("soft_renderer.")]:
            del sys.modules[k]
        sys.modules.update({k: m for k, m in saved.items() if m is not None})
