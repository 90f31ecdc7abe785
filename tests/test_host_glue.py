"""Host-side glue (rows S0, G1-G3, R2 of SURVEY.md section 8a) against golden vectors minted from
the reference's own Python (tests/golden/make_geom_golden.py) and against the numpy oracle.
CPU only: these functions are plain PyTorch for CPU tensors (the CUDA kernels that serve CUDA
tensors are checked in the -m gpu tests)."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geom_ref.npz"))
T = lambda k: torch.from_numpy(GOLD[k])


def close(a, b, tol=2e-6):
    a = a.detach().numpy() if torch.is_tensor(a) else np.asarray(a)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


# ---- oracle restatements pinned against the reference's Python --------------------------------
def test_oracle_lbs_and_cameras(oracle):
    assert close(oracle.obj_to_cam(GOLD["lbs_verts"], GOLD["lbs_R"], GOLD["lbs_T"], 5, 1, GOLD["lbs_skin"]), GOLD["lbs_out"])
    assert close(oracle.obj_to_cam(GOLD["lbs_verts"], GOLD["lbs_R"], GOLD["lbs_T"], 5, 1, GOLD["lbs_skin"], tocam=False),
                 GOLD["lbs_out_notocam"])
    assert close(oracle.obj_to_cam(GOLD["lbs_verts"], GOLD["lbs_R"][::5], GOLD["lbs_T"][::5], 1, 1, None), GOLD["rigid_out"])
    assert close(oracle.pinhole_cam(GOLD["cam_v4"], GOLD["cam_pp"], GOLD["cam_fl"]), GOLD["pinhole_out"])


def test_oracle_softras_functionals(oracle):
    eye = [0, 0, oracle.LASR_EYE_Z]
    assert close(oracle.look_at(GOLD["sr_verts"], eye), GOLD["look_at_out"])
    assert close(oracle.look_at(GOLD["sr_verts"], [0.3, -0.2, -2.0]), GOLD["look_at_eye2"])
    assert close(oracle.orthogonal(GOLD["sr_verts"], 2.0), GOLD["orthogonal_s2"])
    assert np.array_equal(oracle.face_vertices(GOLD["sr_verts"], GOLD["sr_faces"]), GOLD["face_vertices_out"])
    assert close(oracle.vertex_normals(GOLD["sr_verts"], GOLD["sr_faces"]), GOLD["vertex_normals_out"], 1e-5)


# ---- product functions (CPU-tensor path) pinned against the same vectors ------------------------
def test_product_geom_utils():
    from lasr_b200 import geom_utils as gu
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R"), T("lbs_T"), 5, 1, T("lbs_skin")), GOLD["lbs_out"])
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R"), T("lbs_T"), 5, 1, T("lbs_skin"), tocam=False),
                 GOLD["lbs_out_notocam"])
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R")[::5].contiguous(), T("lbs_T")[::5].contiguous(), 1, 1, None),
                 GOLD["rigid_out"])
    assert close(gu.pinhole_cam(T("cam_v4"), T("cam_pp"), T("cam_fl")), GOLD["pinhole_out"])
    assert close(gu.orthographic_cam(T("cam_v4"), T("cam_pp"), T("cam_fl")), GOLD["ortho_out"])


def test_product_softras_functionals():
    import lasr_b200.soft_renderer as sr
    srf = sr.functional
    eye = [0, 0, -(1.0 / np.tan(np.radians(30.0)) + 1)]
    assert close(srf.look_at(T("sr_verts"), eye), GOLD["look_at_out"])
    assert close(srf.look_at(T("sr_verts"), [0.3, -0.2, -2.0]), GOLD["look_at_eye2"])
    assert close(srf.orthogonal(T("sr_verts"), 1.0), GOLD["orthogonal_out"])
    assert close(srf.orthogonal(T("sr_verts"), 2.0), GOLD["orthogonal_s2"])
    assert np.array_equal(srf.face_vertices(T("sr_verts"), T("sr_faces")).numpy(), GOLD["face_vertices_out"])
    assert close(srf.vertex_normals(T("sr_verts"), T("sr_faces")), GOLD["vertex_normals_out"], 1e-5)
    light = torch.zeros(4, 40, 3)
    light = srf.ambient_lighting(light, 0.3, (1, 0.5, 0.25))
    light = srf.directional_lighting(light, T("vertex_normals_out"), 0.7, (0.2, 1, 1), (0, 1, 0))
    assert close(light, GOLD["light_out"])


def test_lasr_camera_chain_is_xy_flip_free_identity(oracle):
    """with LASR's settings look_at(eye) after '+eye' and orthogonal(1) is the identity."""
    import lasr_b200.soft_renderer as sr
    r = sr.SoftRenderer(image_size=32, camera_mode="look_at", perspective=False, light_mode="vertex",
                        light_intensity_ambient=1.0, light_intensity_directionals=0.0)
    assert r.lighting.is_identity()
    v = T("sr_verts") + torch.tensor([0, 0, 10.0])
    eye = torch.tensor(r.transform.transformer._eye)
    out = r.transform.transformer(v + eye)
    assert (out - v).abs().max() < 2e-6
    ref = oracle.lasr_render_vertices(v.numpy())
    flip = v.numpy().copy()
    flip[:, :, 1] *= -1
    assert np.abs(ref - flip).max() < 2e-6


def test_mesh_container_semantics():
    import lasr_b200.soft_renderer as sr
    v, f = T("sr_verts"), T("sr_faces")
    m = sr.Mesh(v, f, textures=torch.rand(4, 40, 3), texture_type="vertex")
    assert m.batch_size == 4 and m.num_vertices == 40 and m.num_faces == 60 and m.texture_res == 1
    assert m.face_vertices.shape == (4, 60, 3, 3) and m.face_textures.shape == (4, 60, 3, 3)
    fv0 = m.face_vertices
    m.vertices = v * 2                                     # setter invalidates the cached gather
    assert torch.equal(m.face_vertices, fv0 * 2)
    m.fill_back_()
    assert m.num_faces == 120 and torch.equal(m.faces[:, 60:], f[:, :, [2, 1, 0]])
    m.reset_()
    assert m.num_faces == 60
    ms = sr.Mesh(v[0], f[0])                               # 2-D inputs get a batch dimension
    assert ms.vertices.shape == (1, 40, 3) and ms.textures.shape == (1, 60, 1, 3)
    with pytest.raises(ValueError):
        sr.SoftRenderer(camera_mode="fisheye")
    with pytest.raises(ValueError):
        sr.Lighting(light_mode="pixel")


def test_quaternion_and_skin(oracle):
    from lasr_b200 import geom_utils as gu
    ident = gu.quaternion_to_rotation_matrix(torch.tensor([[0.0, 0.0, 0.0, 1.0]]))
    assert torch.allclose(ident[0], torch.eye(3))
    g = torch.Generator().manual_seed(0)
    q = torch.randn(16, 4, generator=g)
    R = gu.quaternion_to_rotation_matrix(q)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(16, 3, 3), atol=1e-5)
    assert torch.allclose(torch.linalg.det(R), torch.ones(16), atol=1e-5)
    # 90 degrees about z (x,y,z,w) = (0,0,sin45,cos45) maps x -> y
    Rz = gu.quaternion_to_rotation_matrix(torch.tensor([[0.0, 0.0, np.sqrt(0.5), np.sqrt(0.5)]], dtype=torch.float32))[0]
    assert torch.allclose(Rz @ torch.tensor([1.0, 0.0, 0.0]), torch.tensor([0.0, 1.0, 0.0]), atol=1e-6)
    assert np.allclose(oracle.quaternion_to_rotation_matrix(q.numpy()), R.numpy(), atol=1e-6)
    H, B1, V = 2, 4, 30
    ctl_ts, ctl_rs = torch.randn(H * B1, 3, generator=g) * 0.3, torch.randn(H * B1, 4, generator=g)
    log_ctl, mean_v = torch.randn(H * B1, 3, generator=g) * 0.2, torch.randn(H, V, 3, generator=g) * 0.3
    skin = gu.skin_weights(ctl_ts, ctl_rs, log_ctl, mean_v, H)
    assert skin.shape == (H, B1, V, 1) and torch.allclose(skin.sum(1), torch.ones(H, V, 1), atol=1e-6)
    assert np.allclose(oracle.skin_weights(ctl_ts.numpy(), ctl_rs.numpy(), log_ctl.numpy(), mean_v.numpy(), H),
                       skin.numpy(), atol=2e-6)


def test_depth_window(oracle):
    from lasr_b200 import geom_utils as gu
    z = T("pinhole_out")[:, :, 2]
    near, far = gu.depth_window(z)
    n2, f2 = oracle.near_far(z.numpy())
    assert abs(float(near) - n2) < 1e-6 and abs(float(far) - f2) < 1e-6


def test_image_losses_match_loop_restatement(oracle):
    """L1-L3: the vectorised product losses equal the numpy restatement of the reference's Python
    double loops (nnutils/mesh_net.py:372-447)."""
    from lasr_b200 import losses
    g = torch.Generator().manual_seed(3)
    N, H, S = 4, 2, 24
    mask_pred = torch.rand(N, H, S, S, generator=g)
    masks = (torch.rand(N, S, S, generator=g) > 0.5).float()
    occ = torch.randn(N, S, S, generator=g)
    occ[torch.rand(N, S, S, generator=g) < 0.2] = 0
    l, sub = losses.mask_loss(mask_pred, masks, occ)
    lo, subo = oracle.mask_loss(mask_pred.numpy(), masks.numpy(), occ.numpy())
    assert np.allclose(sub.numpy(), subo, rtol=1e-5, atol=1e-7) and abs(float(l) - lo) < 1e-6

    flow_rd = torch.randn(N, H, S, S, 2, generator=g) * 0.05
    bg = torch.rand(N, H, S, S, generator=g) < 0.4
    bg[1, 0] = True                                           # an empty mask for one (frame, hypothesis)
    flow_obs = torch.randn(N, 3, S, S, generator=g) * 0.05
    l, sub, _ = losses.flow_loss(flow_rd, bg, flow_obs, occ, masks)
    lo, subo = oracle.flow_loss(flow_rd.numpy(), bg.numpy(), flow_obs.numpy(), occ.numpy(), masks.numpy())
    assert sub[1, 0] == 0
    assert np.allclose(sub.numpy(), subo, rtol=2e-5, atol=1e-7) and abs(float(l) - lo) < 1e-6

    tex = torch.rand(N, H, 3, S, S, generator=g)
    fg = torch.rand(N, H, S, S, generator=g)
    imgs = torch.rand(N, 3, S, S, generator=g)
    l, sub = losses.texture_l1_loss(tex, fg, imgs, masks, occ, 1.0)
    lo, subo = oracle.texture_l1_loss(tex.numpy(), fg.numpy(), imgs.numpy(), masks.numpy(), occ.numpy(), 1.0)
    assert np.allclose(sub.numpy(), subo, rtol=2e-5, atol=1e-7) and abs(float(l) - lo) < 1e-6


def test_flow_postprocessing_matches_oracle(oracle):
    from lasr_b200 import geom_utils as gu
    g = torch.Generator().manual_seed(5)
    n, S = 2, 16
    p0 = torch.rand(n, S, S, 3, generator=g) + 9
    p1 = torch.rand(n, S, S, 3, generator=g) + 9
    p0[0, :4] = 0                       # background rows (rendered z < 1e-9)
    pp0, pp1 = torch.randn(n, 2, generator=g) * 0.1, torch.randn(n, 2, generator=g) * 0.1
    f0, f1 = 9 + torch.rand(n, 1, generator=g), 9 + torch.rand(n, 1, generator=g)
    fo, bgo = oracle.flow_from_rendered_positions(p0.numpy(), p1.numpy(), pp0.numpy(), pp1.numpy(), f0.numpy(), f1.numpy())
    bg = (p0[..., 2] < 1e-9) | (p1[..., 2] < 1e-9)
    ten = torch.full_like(p0, 10.0)
    q0, q1 = torch.where(bg[..., None], ten, p0), torch.where(bg[..., None], ten, p1)
    flow = gu._reproject(q1, pp1, f1) - gu._reproject(q0, pp0, f0)
    assert np.array_equal(bg.numpy(), bgo) and np.allclose(flow.numpy(), fo, atol=1e-6)


def test_flow_render_matches_the_reference_source_lines(oracle):
    """R1 pinned: tests/golden/flow_ref.npz holds what the reference's own render_flow_soft_2 (nnutils/mesh_net.py:75-104,
    executed unmodified with a stand-in renderer returning a seeded render) hands to the rasterizer and makes of its
    output; the product's function with the same stand-in, and the oracle's re-projection, must reproduce both."""
    import types
    from lasr_b200 import geom_utils as gu
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_ref.npz"))
    T = torch.from_numpy
    for case in ("h1", "h2"):
        I = lambda k: T(g[f"{case}/in/{k}"])
        seen = {}

        def render_mesh(mesh):
            seen["vertices"], seen["faces"], seen["textures"] = mesh.vertices, mesh.faces, mesh.textures
            return I("rendered").clone()
        renderer = types.SimpleNamespace(
            transform=types.SimpleNamespace(transformer=types.SimpleNamespace(_eye=[0, 0, -2.732051])), render_mesh=render_mesh)
        flow, bg, fgm = gu.render_flow_soft_2(renderer, I("verts"), I("faces"), I("pos0"), I("pos1"), I("pp0"), I("pp1"),
                                              I("f0"), I("f1"))
        assert np.array_equal(seen["vertices"].numpy(), g[f"{case}/mesh/vertices"])          # eye offset + y flip, exact
        assert np.array_equal(seen["faces"].numpy(), g[f"{case}/mesh/faces"])
        assert np.array_equal(seen["textures"].numpy(), g[f"{case}/mesh/textures"])
        assert np.array_equal(bg.numpy(), g[f"{case}/out/bgmask"]) and bg.any()
        assert np.array_equal(fgm.numpy(), g[f"{case}/out/fgmask"])
        np.testing.assert_allclose(flow.numpy(), g[f"{case}/out/flow"], rtol=1e-6, atol=1e-6)
        n = g[f"{case}/in/verts"].shape[0]
        r = g[f"{case}/in/rendered"]
        fo, bgo = oracle.flow_from_rendered_positions(r[:n, :3].transpose(0, 2, 3, 1), r[n:, :3].transpose(0, 2, 3, 1),
                                                      g[f"{case}/in/pp0"], g[f"{case}/in/pp1"], g[f"{case}/in/f0"], g[f"{case}/in/f1"])
        assert np.array_equal(bgo, g[f"{case}/out/bgmask"])
        np.testing.assert_allclose(fo, g[f"{case}/out/flow"], rtol=1e-6, atol=1e-6)


def test_install_as_soft_renderer_exposes_what_lasr_uses():
    """nnutils/mesh_net.py:47,132-149,306-316,354-356 and nnutils/geom_utils.py:24 import and poke these."""
    import importlib
    import sys
    import lasr_b200
    saved = {k: sys.modules.get(k) for k in ("soft_renderer", "soft_renderer.functional")}
    try:
        lasr_b200.install_as_soft_renderer()
        sr = importlib.import_module("soft_renderer")
        r = sr.SoftRenderer(image_size=32, sigma_val=1e-4, gamma_val=1e-2, camera_mode="look_at", perspective=False,
                            aggr_func_rgb="softmax", light_mode="vertex", light_intensity_ambient=1.0,
                            light_intensity_directionals=0.0)
        r.rasterizer.near, r.rasterizer.far = 1.0, 2.0
        r.rasterizer.sigma_val = 1e-5
        r.rasterizer.background_color = [1, 1, 1]
        assert len(r.transform.transformer._eye) == 3 and abs(r.transform.transformer._eye[2] + 2.7320508) < 1e-6
        r.set_sigma(1e-4); r.set_gamma(1e-3); r.set_texture_mode("vertex")
        assert r.rasterizer.sigma_val == 1e-4 and r.rasterizer.gamma_val == 1e-3 and r.lighting.light_mode == "vertex"
        assert hasattr(sr, "Mesh") and hasattr(sr.functional, "soft_rasterize") and hasattr(sr.functional, "face_vertices")
        with pytest.raises(NotImplementedError):
            sr.functional.voxelization(None)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _losses_golden():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_ref.npz"))
    for case in ("h1", "h2"):
        bs, H, S = (int(v) for v in g[f"{case}/meta"])
        n2 = 2 * bs
        I = lambda k: g[f"{case}/in/{k}"]
        args = dict(mask_pred=I("render")[:, 3].reshape(n2, H, S, S), masks=I("masks"), occ=I("occ"),
                    flow_rd=I("flow_rd").reshape(n2, H, S, S, 2), bgmask=I("bgmask").reshape(n2, H, S, S), flow=I("flow"),
                    tex=I("render")[:, :3].reshape(n2, H, 3, S, S), imgs=I("imgs"), wt=float(g[f"{case}/l1tex_wt"]))
        yield case, args, {k: g[f"{case}/out/{k}"] for k in ("mask_loss_sub", "flow_rd_loss_sub", "texture_loss_sub",
                                                              "mask_loss", "flow_rd_loss", "texture_loss", "total_loss")}


def test_image_losses_match_the_reference_source_lines(oracle):
    """L1-L3 pinned: tests/golden/losses_ref.npz holds what the reference's own lines (nnutils/mesh_net.py:361-363,
    372-447, executed unmodified on CPU tensors, LPIPS stubbed to zero) compute; the oracle's loop restatement and the
    product's vectorised host path must both reproduce it."""
    from lasr_b200 import losses
    for case, a, want in _losses_golden():
        lo, subo = oracle.mask_loss(a["mask_pred"], a["masks"], a["occ"])
        assert np.allclose(subo, want["mask_loss_sub"], rtol=1e-5, atol=1e-7) and abs(lo - float(want["mask_loss"])) < 1e-6
        lo, subo = oracle.flow_loss(a["flow_rd"], a["bgmask"], a["flow"], a["occ"], a["masks"])
        assert np.allclose(subo, want["flow_rd_loss_sub"], rtol=2e-5, atol=1e-7) and abs(lo - float(want["flow_rd_loss"])) < 1e-6
        assert (want["flow_rd_loss_sub"] == 0).any()            # the empty-mask branch (:412) is in the golden
        lo, subo = oracle.texture_l1_loss(a["tex"], a["mask_pred"], a["imgs"], a["masks"], a["occ"], a["wt"])
        assert np.allclose(subo, want["texture_loss_sub"], rtol=2e-5, atol=1e-7) and abs(lo - float(want["texture_loss"])) < 1e-6
        T = torch.from_numpy
        l1, s1 = losses.mask_loss(T(a["mask_pred"]), T(a["masks"]), T(a["occ"]))
        l2, s2, _ = losses.flow_loss(T(a["flow_rd"]), T(a["bgmask"]), T(a["flow"]), T(a["occ"]), T(a["masks"]))
        l3, s3 = losses.texture_l1_loss(T(a["tex"]), T(a["mask_pred"]), T(a["imgs"]), T(a["masks"]), T(a["occ"]), a["wt"])
        assert np.allclose(s1.numpy(), want["mask_loss_sub"], rtol=1e-5, atol=1e-7)
        assert np.allclose(s2.numpy(), want["flow_rd_loss_sub"], rtol=2e-5, atol=1e-7)
        assert np.allclose(s3.numpy(), want["texture_loss_sub"], rtol=2e-5, atol=1e-7)
        assert abs(float(l1 + l2 + l3) - float(want["total_loss"])) < 2e-6, case


class _OracleRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fv, tex, settings):
        from oracle import oracle as O
        s = {k: (float(v) if torch.is_tensor(v) else v) for k, v in settings.items()}
        r = O.soft_rasterize_forward(fv.detach().numpy(), tex.detach().numpy(), s, np.float32)
        ctx.r, ctx.shapes = r, (fv.shape, tex.shape)
        return torch.from_numpy(r["soft_colors"])

    @staticmethod
    def backward(ctx, g):
        from oracle import oracle as O
        gf, gt = O.soft_rasterize_backward(ctx.r, g.contiguous().numpy())
        return torch.from_numpy(gf).reshape(ctx.shapes[0]), torch.from_numpy(gt).reshape(ctx.shapes[1]), None


def test_hot_path_region_runs_on_cpu_with_oracle_rasterizer(monkeypatch):
    """The whole mesh_net.py:255-447 region (skin -> LBS -> projection -> 3 renders -> 3 losses) on CPU
    tensors with the oracle as the rasterizer: finite losses, gradients reach every input, and the
    losses equal the loop-for-loop restatement applied to the same renders."""
    from lasr_b200 import hotpath, synth
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    import lasr_b200.soft_renderer.rasterizer as rast
    from oracle import oracle as O

    def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
             aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))

    monkeypatch.setattr(rast.srf, "soft_rasterize", fake)
    g = torch.Generator().manual_seed(0)
    H, B, S, bs = 1, 3, 24, 1
    v, f = synth.icosphere(1, 0.4)
    N = 2 * bs * H
    pred_v = torch.from_numpy(v)[None].repeat(N, 1, 1).requires_grad_(True)
    tex = torch.sigmoid(torch.randn(N, len(v), 3, generator=g)).requires_grad_(True)
    faces = torch.from_numpy(f)[None].repeat(2 * bs, 1, 1)
    q = torch.randn(N * B, 4, generator=g) * 0.1
    q[:, 3] = 1
    Rmat = quaternion_to_rotation_matrix(q).permute(0, 2, 1).contiguous().requires_grad_(True)
    Tm = torch.randn(N * B, 3, generator=g) * 0.02
    Tm.view(N, B, 3)[:, 0, 2] += 10
    Tmat = Tm.requires_grad_(True)
    ppoint = (torch.randn(2 * bs, 2, generator=g) * 0.05).requires_grad_(True)
    scale = (9.5 + torch.rand(N, 1, generator=g)).requires_grad_(True)
    ctl = torch.from_numpy(v)[:H * (B - 1)].clone()
    bones = dict(ctl_ts=ctl.clone().requires_grad_(True), rest_ts=(ctl + 0.01).requires_grad_(True),
                 ctl_rs=torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(H * (B - 1), 1).requires_grad_(True),
                 log_ctl=(1 + torch.randn(H * (B - 1), 3, generator=g) * 0.3).requires_grad_(True))
    occ = torch.randn(2 * bs, S, S, generator=g)
    batch = dict(masks=(torch.rand(2 * bs, S, S, generator=g) > 0.4).float(), imgs=torch.rand(2 * bs, 3, S, S, generator=g),
                 flow=torch.randn(2 * bs, 3, S, S, generator=g) * 0.05, occ=occ)
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B))
    out = m(pred_v, tex, faces, Rmat, Tmat, ppoint, scale, batch, bones)
    out["total"].backward()
    assert all(torch.isfinite(out[k]) for k in ("total", "mask_loss", "flow_loss", "texture_loss"))
    for name, t in dict(pred_v=pred_v, tex=tex, Rmat=Rmat, Tmat=Tmat, ppoint=ppoint, scale=scale, **bones).items():
        assert t.grad is not None and torch.isfinite(t.grad).all(), name
    assert pred_v.grad.abs().max() > 0 and tex.grad.abs().max() > 0 and Rmat.grad.abs().max() > 0
    n2 = 2 * bs
    lo, _ = O.mask_loss(out["mask_pred"].detach().view(n2, H, S, S).numpy(), batch["masks"].numpy(), occ.numpy())
    assert abs(float(out["mask_loss"]) - lo) < 1e-6
    lo, _ = O.flow_loss(out["flow_rd"].detach().view(n2, H, S, S, 2).numpy(), out["bgmask"].view(n2, H, S, S).numpy(),
                        batch["flow"].numpy(), occ.numpy(), batch["masks"].numpy())
    assert abs(float(out["flow_loss"]) - lo) < 1e-6
    lo, _ = O.texture_l1_loss(out["texture_render"].detach().reshape(n2, H, 3, S, S).numpy(),
                              out["mask_pred"].detach().view(n2, H, S, S).numpy(), batch["imgs"].numpy(),
                              batch["masks"].numpy(), occ.numpy(), 1.0)
    assert abs(float(out["texture_loss"]) - lo) < 1e-6


@pytest.mark.parametrize("case", ["b1", "b4", "b3h2"])
def test_region_posing_matches_the_reference_source_lines(monkeypatch, case):
    """The geometry half of the region pinned: tests/golden/region_ref.npz holds what the reference's own lines
    (nnutils/mesh_net.py:259-311 with its own obj_to_cam / pinhole_cam, executed unmodified on CPU tensors; only
    kornia's quaternion conversion is a stand-in) compute -- skinning weights, composed bone transforms, posed and
    projected vertices, the batch-global depth window.  RenderHotPath (three-render structure, CPU tensors, oracle
    as the rasterizer) must arrive at the same projected vertices and window from the same pose inputs."""
    from lasr_b200 import geom_utils as gu
    from lasr_b200 import hotpath, synth
    import lasr_b200.soft_renderer.rasterizer as rast
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "region_ref.npz"))
    bs, H, B, V = (int(v) for v in g[f"{case}/meta"])
    I = lambda k: torch.from_numpy(g[f"{case}/in/{k}"])
    O_ = lambda k: g[f"{case}/out/{k}"]

    def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
             aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))

    monkeypatch.setattr(rast.srf, "soft_rasterize", fake)
    n2, N, S = 2 * bs, 2 * bs * H, 12
    bones = None
    if B > 1:
        bones = {k: I(k) for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")}
        skin = gu.skin_weights(bones["ctl_ts"], bones["ctl_rs"], bones["log_ctl"], I("pred_v").view(n2, H, V, 3)[0], H)
        np.testing.assert_allclose(skin.numpy(), O_("skin")[:H], rtol=2e-5, atol=1e-7)       # S0 (given the conversion)
    Rmat = I("quat").view(-1, 3, 3).permute(0, 2, 1).contiguous()            # mesh_net.py:259
    Tmat = torch.cat([I("trans"), I("depth")], 1)                            # :260
    rng = np.random.RandomState(0)
    faces = torch.from_numpy(rng.randint(0, V, (1, 12, 3)).astype(np.int32)).repeat(n2, 1, 1)
    gt = torch.Generator().manual_seed(1)
    batch = dict(masks=(torch.rand(n2, S, S, generator=gt) > 0.4).float(), imgs=torch.rand(n2, 3, S, S, generator=gt),
                 flow=torch.randn(n2, 3, S, S, generator=gt) * 0.05, occ=torch.randn(n2, S, S, generator=gt))
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B))
    out = m(I("pred_v"), torch.rand(N, V, 3, generator=gt), faces, Rmat, Tmat, I("ppoint"), I("scale"), batch, bones)
    np.testing.assert_allclose(out["verts_px"].detach().numpy(), O_("verts_fl"), rtol=2e-5, atol=2e-6)
    assert abs(float(out["near"]) - float(O_("near"))) < 1e-5 and abs(float(out["far"]) - float(O_("far"))) < 1e-5


@pytest.mark.parametrize("case", ["b1", "b3", "b3h2"])
def test_whole_region_matches_the_reference_lines_running_on_the_drop_in_package(monkeypatch, case):
    """tests/golden/region_losses_ref.npz: the reference's own LASR.forward lines (nnutils/mesh_net.py:259-447 and its
    render_flow_soft_2, :75-104), executed on CPU tensors ON TOP OF lasr_b200.soft_renderer (oracle as the rasterizer
    backend), give these renders and losses.  RenderHotPath -- the one-call replacement of that region -- must give
    the same from the same inputs: rendered silhouette and flow, and the three losses."""
    from lasr_b200 import hotpath
    import lasr_b200.soft_renderer.rasterizer as rast
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "region_losses_ref.npz"))
    bs, H, B, S = (int(v) for v in g[f"{case}/meta"])
    I = lambda k: torch.from_numpy(g[f"{case}/in/{k}"])
    W = lambda k: g[f"{case}/out/{k}"]

    def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
             aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))

    monkeypatch.setattr(rast.srf, "soft_rasterize", fake)
    bones = {k: I(k) for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")} if B > 1 else None
    Rmat = I("quat").view(-1, 3, 3).permute(0, 2, 1).contiguous()            # mesh_net.py:259
    Tmat = torch.cat([I("trans"), I("depth")], 1)                            # :260
    batch = dict(masks=I("masks"), imgs=I("imgs"), flow=I("flow"), occ=I("occ"))
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B, l1tex_wt=1.0))
    out = m(I("pred_v"), I("tex"), I("faces"), Rmat, Tmat, I("ppoint"), I("scale"), batch, bones)
    np.testing.assert_allclose(out["mask_pred"].detach().numpy(), W("mask_pred"), rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["flow_rd"].detach().numpy(), W("flow_rd"), rtol=0, atol=5e-6)
    for ours, ref in (("mask_loss", "mask_loss"), ("flow_loss", "flow_rd_loss"), ("texture_loss", "texture_loss"), ("total", "total_loss")):
        assert abs(float(out[ours]) - float(W(ref))) < 2e-6, (ours, float(out[ours]), float(W(ref)))


def test_package_surface_matches_reference_exports():
    """third_party/softras/soft_renderer/__init__.py:1-15 exports these names; LASR reaches the first three
    (nnutils/mesh_net.py, lasr/lasr.py), the rest keep `import soft_renderer as sr` code importable."""
    import lasr_b200.soft_renderer as sr
    for name in ("Mesh", "SoftRenderer", "Renderer", "Transform", "Projection", "LookAt", "Look", "Lighting",
                 "AmbientLighting", "DirectionalLighting", "Rasterizer", "SoftRasterizer", "LaplacianLoss",
                 "FlattenLoss", "functional", "losses"):
        assert hasattr(sr, name), name
    assert sr.losses.LaplacianLoss is sr.LaplacianLoss
    with pytest.raises(NotImplementedError):
        sr.Renderer()


def test_synthetic_workload_is_the_reference_generators(monkeypatch):
    """tests/golden/render_syn_ref.npz is what the reference's own scripts/render_syn.py (run whole and unmodified on
    the drop-in package, tests/golden/make_render_syn_golden.py) writes for three frames of Spot: camera files, posed
    meshes, 512^2 silhouettes and flow files.  bench.py's workload (lasr_b200.synth: Spot fixture, camera orbit,
    projection) must be that generator's geometry, and the package's flow render -- `render_flow_soft_2`, the trainer's
    copy of the script's `render_flow` -- must reproduce its silhouettes and flow files from it."""
    import lasr_b200.soft_renderer as sr
    import lasr_b200.soft_renderer.rasterizer as rast
    from lasr_b200 import geom_utils as gu, synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_syn_ref.npz"))
    n, S = int(g["nframes"]), int(g["img_size"])
    v, f = synth.load_spot()
    R, T, pp, fl = synth.spot_cameras(n)
    px, cam = synth.project_frames(v, R, T, pp, fl)
    for i in range(n):
        c = g[f"cam{i}"]                       # [focal, tx, ty, qw, qx, qy, qz, depth]  (render_syn.py:150-157,193)
        assert c[0] == fl[i, 0] and c[7] == T[i, 0, 2] and (c[1:3] == T[i, 0, :2]).all()
        q_conj = torch.tensor([[-c[4], -c[5], -c[6], c[3]]])                     # render_syn.py:162: obj -> cam rotation
        np.testing.assert_allclose(gu.quaternion_to_rotation_matrix(q_conj)[0].numpy(), R[i], atol=2e-7)
        np.testing.assert_allclose(cam[i], g[f"mesh{i}"], rtol=0, atol=2e-6)     # posed vertices (the exported meshes)
        np.testing.assert_allclose(px[i, :, :2], g[f"verts_px{i}"][:, :2], rtol=0, atol=2e-6)   # projected x, y

    def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
             aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))
    monkeypatch.setattr(rast.srf, "soft_rasterize", fake)
    renderer = sr.SoftRenderer(image_size=S, sigma_val=1e-12, camera_mode="look_at", perspective=False,
                               aggr_func_rgb="hard", light_mode="vertex", light_intensity_ambient=1.,
                               light_intensity_directionals=0.)                  # render_syn.py:139-141
    eye_z = float(renderer.transform.transformer._eye[2])
    z = cam[:, :, 2]                                                             # render_syn.py:174: z to [-0.5, 0.5]
    zn = (z - z.min(1, keepdims=True)) / (z.max(1, keepdims=True) - z.min(1, keepdims=True)) - 0.5
    T_ = torch.from_numpy
    verts = T_(np.concatenate([px[:, :, :2], zn[:, :, None]], -1).astype(np.float32))
    np.testing.assert_allclose(verts.numpy(), np.stack([g[f"verts_px{i}"][:, :3] for i in range(n)]), atol=3e-6)
    verts[:, :, 2] -= 2 * eye_z                    # the script subtracts the eye offset where the trainer adds it (:44 vs mesh_net.py:81)
    pos = T_(cam)
    faces = T_(f)[None]
    zero, foc = torch.zeros(1, 2), torch.full((1, 1), 10.0)
    for i in range(1, n):
        for tag, a, b, j in (("FlowFW", i - 1, i, i - 1), ("FlowBW", i, i - 1, i)):   # render_syn.py:203-211
            flow, bg, fgm = gu.render_flow_soft_2(renderer, verts[a:a + 1], faces, pos[a:a + 1], pos[b:b + 1], zero, zero, foc, foc)
            out = np.concatenate([(flow / 2 * (S - 1)).numpy()[0], 1 - bg.float().numpy()[0][:, :, None]], -1)
            ref, ref_sum = g[f"{tag}{j}"], g[f"{tag}{j}_sum"]
            assert np.array_equal(out[::4, ::4, 2], ref[:, :, 2]) and out[:, :, 2].sum() == ref_sum[2]
            # flow in pixels: 'hard' colour aggregation takes the nearest face's value, so where synth's pose and the
            # script's differ by float rounding a pixel on a crease reads the neighbouring face (<= 0.01 px off) and one
            # on an occlusion edge the other surface (pixels off)
            d = np.abs(out[::4, ::4, :2] - ref[:, :, :2])
            assert (d > 1e-4 + 1e-5 * np.abs(ref[:, :, :2])).mean() < 1e-3 and (d > 0.05).sum() <= 4, ((d > 1e-4).sum(), d.max())
            np.testing.assert_allclose(out.astype(np.float64).sum((0, 1)), ref_sum, rtol=2e-4)
            mask = np.clip(np.rint(128 * fgm[0].numpy()), 0, 255).astype(np.uint8)               # render_syn.py:191
            # the silhouette edge is hard (sigma 1e-12) and synth's pose differs from the script's by float rounding
            # (cv2.Rodrigues -> quaternion -> matrix there): a few edge pixels of 262 144 may flip
            assert (mask != g[f"mask{a}"]).sum() <= 8, (mask != g[f"mask{a}"]).sum()


def test_flow_render_3_matches_the_reference_source_lines():
    """`render_flow_soft_3` (nnutils/geom_utils.py:73-95, the evaluation scripts' flow render: rendered target positions
    against the pixel grid) pinned the same way: tests/golden/flow_ref.npz cases s3a / s3b are the reference's own
    lines executed with a stand-in renderer."""
    import types
    from lasr_b200 import geom_utils as gu
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flow_ref.npz"))
    T = torch.from_numpy
    for case in ("s3a", "s3b"):
        I = lambda k: T(g[f"{case}/in/{k}"])
        seen = {}

        def render_mesh(mesh):
            seen["vertices"], seen["faces"], seen["textures"] = mesh.vertices, mesh.faces, mesh.textures
            return I("rendered").clone()
        renderer = types.SimpleNamespace(
            transform=types.SimpleNamespace(transformer=types.SimpleNamespace(_eye=[0, 0, -2.732051])), render_mesh=render_mesh)
        flow, bg, fgm = gu.render_flow_soft_3(renderer, I("verts"), I("target"), I("faces"))
        assert np.array_equal(seen["vertices"].numpy(), g[f"{case}/mesh/vertices"])
        assert np.array_equal(seen["faces"].numpy(), g[f"{case}/mesh/faces"])
        assert np.array_equal(seen["textures"].numpy(), g[f"{case}/mesh/textures"])
        assert np.array_equal(bg.numpy(), g[f"{case}/out/bgmask"]) and bg.any()
        assert np.array_equal(fgm.numpy(), g[f"{case}/out/fgmask"])
        np.testing.assert_allclose(flow.numpy(), g[f"{case}/out/flow"], rtol=1e-6, atol=1e-6)


def test_module_layer_matches_the_reference_classes(monkeypatch):
    """R2 at the module level: tests/golden/modules_ref.npz records what the reference's own SoftRenderer / Lighting /
    Transform / SoftRasterizer / Mesh classes (loaded unmodified, tests/golden/make_modules_golden.py) hand to the
    rasterizer and return, for cameras, lights, texture types and setters LASR's one configuration does not exercise.
    The mirror must hand over the same face vertices, face textures and scalars, and pool the render the same way."""
    import json
    import lasr_b200.soft_renderer as sr
    import lasr_b200.soft_renderer.rasterizer as rast
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modules_ref.npz"))
    names = sorted({k.split("/")[0] for k in g.files if not k.startswith("mesh_ops")})
    assert len(names) == 8
    arg_names = ("image_size", "background_color", "near", "far", "fill_back", "eps", "sigma_val", "dist_func",
                 "dist_eps", "gamma_val", "aggr_func_rgb", "aggr_func_alpha", "texture_type")
    for name in names:
        meta = json.loads(str(g[f"{name}/meta"]))
        rec = {}

        def soft_rasterize(face_vertices, textures, *args):
            rec["face_vertices"], rec["face_textures"], rec["args"] = face_vertices, textures, dict(zip(arg_names, args))
            return torch.from_numpy(g[f"{name}/rec/render"]).clone()
        monkeypatch.setattr(rast.srf, "soft_rasterize", soft_rasterize)
        r = sr.SoftRenderer(**meta["kwargs"])
        for fn, a in meta["calls"]:
            a = [torch.tensor(x) if isinstance(x, list) and fn == "set_eyes_from_angles" else x for x in a]
            (getattr(r, fn) if hasattr(r, fn) else getattr(r.transform, fn))(*a)
        verts, faces = torch.from_numpy(g[f"{name}/in/verts"]), torch.from_numpy(g[f"{name}/in/faces"])
        tex = torch.from_numpy(g[f"{name}/in/textures"]) if f"{name}/in/textures" in g.files else None
        if meta["via_forward"]:
            img = r(verts, faces, tex, texture_type=meta["texture_type"])
        else:
            img = r.render_mesh(sr.Mesh(verts, faces, textures=tex, texture_res=meta["texture_res"],
                                        texture_type=meta["texture_type"]))
        np.testing.assert_allclose(rec["face_vertices"].numpy(), g[f"{name}/rec/face_vertices"], rtol=1e-6, atol=1e-6,
                                   err_msg=name)
        np.testing.assert_allclose(rec["face_textures"].numpy(), g[f"{name}/rec/face_textures"], rtol=1e-6, atol=1e-7,
                                   err_msg=name)
        for k, v in meta["args"].items():
            got = rec["args"][k]
            assert (list(got) == list(v)) if isinstance(v, list) else (got == v), (name, k, got, v)
        np.testing.assert_allclose(img.numpy(), g[f"{name}/out/image"], rtol=1e-6, atol=1e-7, err_msg=name)


def test_mesh_bookkeeping_matches_the_reference_class():
    """Mesh.fill_back_ / reset_ / the vertices setter and the cached derived tensors against the reference's own Mesh
    (mesh.py:67-146; tests/golden/modules_ref.npz, mesh_ops_*) -- including its habit of doubling per-vertex textures
    along the vertex axis when filling back."""
    import lasr_b200.soft_renderer as sr
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modules_ref.npz"))
    T = torch.from_numpy
    for ttype in ("surface", "vertex"):
        pre = f"mesh_ops_{ttype}"
        mesh = sr.Mesh(T(g[f"{pre}/in/verts"]), T(g[f"{pre}/in/faces"]), textures=T(g[f"{pre}/in/textures"]), texture_res=2,
                       texture_type=ttype)
        mesh.surface_normals
        mesh.fill_back_()
        mesh.fill_back_()
        assert np.array_equal(mesh.faces.numpy(), g[f"{pre}/filled/faces"])
        assert np.array_equal(mesh.textures.numpy(), g[f"{pre}/filled/textures"])
        assert np.array_equal(mesh.face_vertices.numpy(), g[f"{pre}/filled/face_vertices"])
        np.testing.assert_allclose(mesh.surface_normals.numpy(), g[f"{pre}/filled/surface_normals"], atol=1e-6)
        np.testing.assert_allclose(mesh.vertex_normals.numpy(), g[f"{pre}/filled/vertex_normals"], atol=1e-6)
        if ttype == "surface":
            assert np.array_equal(mesh.face_textures.numpy(), g[f"{pre}/filled/face_textures"])
        mesh.vertices = mesh.vertices * 2.0
        assert np.array_equal(mesh.face_vertices.numpy(), g[f"{pre}/scaled/face_vertices"])
        mesh.reset_()
        assert np.array_equal(mesh.faces.numpy(), g[f"{pre}/reset/faces"])
        assert np.array_equal(mesh.textures.numpy(), g[f"{pre}/reset/textures"])
        assert np.array_equal(mesh.face_vertices.numpy(), g[f"{pre}/reset/face_vertices"])
        import json
        meta = json.loads(str(g[f"{pre}/meta"]))
        assert mesh.texture_res == meta["texture_res"] and mesh.batch_size == meta["batch_size"]


def _full_region_inputs():
    """Inputs and golden of tests/golden/region_full_ref.npz (Spot, 26 bones, 256^2, two frame pairs): the inputs are
    regenerated from the seed by the generator script's own ``inputs()``."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_region_full_golden", os.path.join(here, "golden", "make_region_full_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = np.load(os.path.join(here, "golden", "region_full_ref.npz"))
    bs, H, B, S, seed = (int(v) for v in g["meta"])
    return gen.inputs(seed, bs, H, B, S), g, (bs, H, B, S)


def test_full_size_region_on_cpu_matches_the_reference_lines_with_gradients(monkeypatch):
    """tests/golden/region_full_ref.npz: the reference's own LASR.forward lines, with autograd, at LASR's articulated
    shape (Spot, 26 bones, 256^2).  RenderHotPath on CPU tensors (three renders, oracle as the rasterizer) must give
    the same losses, silhouette and -- through its own autograd -- the same gradient for every input."""
    from lasr_b200 import hotpath
    import lasr_b200.soft_renderer.rasterizer as rast
    x, g, (bs, H, B, S) = _full_region_inputs()

    def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
             eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
             aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))

    monkeypatch.setattr(rast.srf, "soft_rasterize", fake)
    L = {k: x[k].clone().requires_grad_(True) for k in ("quat", "trans", "depth", "pred_v", "tex", "ppoint", "scale",
                                                        "ctl_ts", "rest_ts", "ctl_rs", "log_ctl")}
    Rmat = L["quat"].view(-1, 3, 3).permute(0, 2, 1).contiguous()            # mesh_net.py:259
    Tmat = torch.cat([L["trans"], L["depth"]], 1)                            # :260
    batch = dict(masks=x["masks"], imgs=x["imgs"], flow=x["flow"], occ=x["occ"])
    bones = {k: L[k] for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")}
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B, l1tex_wt=1.0))
    out = m(L["pred_v"], L["tex"], x["faces"], Rmat, Tmat, L["ppoint"], L["scale"], batch, bones)
    out["total"].backward()
    for ours, ref in (("mask_loss", "mask_loss"), ("flow_loss", "flow_rd_loss"), ("texture_loss", "texture_loss"), ("total", "total_loss")):
        assert abs(float(out[ours]) - float(g[f"out/{ref}"])) < 5e-6, (ours, float(out[ours]), float(g[f"out/{ref}"]))
    sil = g["out/mask_pred_u16"].astype(np.float32) / 65535.0
    d = np.abs(out["mask_pred"].detach().numpy() - sil)
    assert (d > 1e-3).mean() < 1e-4, float((d > 1e-3).mean())            # isolated flips at edge-on faces only
    for k, t in L.items():
        ref = torch.from_numpy(g[f"grad/{k}"])
        err = float((t.grad - ref).norm() / ref.norm())
        assert err < 1e-3, (k, err)
