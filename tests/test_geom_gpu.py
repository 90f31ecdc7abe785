"""GPU parity of the fused LBS + projection kernels against the plain
PyTorch expressions of lasr_b200.geom_utils on CPU tensors -- which tests/test_host_glue.py pins to
golden vectors minted from the reference's own nnutils/geom_utils.py -- forward and backward."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geom_ref.npz"))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def rand_rot(n, g):
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    return quaternion_to_rotation_matrix(torch.randn(n, 4, generator=g))


def make_inputs(N, V, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    verts = torch.randn(N, V, 3, generator=g) * 0.3
    R = rand_rot(N * B, g)
    T = torch.randn(N * B, 1, 3, generator=g) * 0.05
    T.view(N, B, 1, 3)[:, 0, 0, 2] += 10.0
    skin = torch.softmax(torch.randn(N, max(B - 1, 1), V, 1, generator=g) * 2, dim=1) if B > 1 else None
    pp = torch.randn(N, 2, generator=g) * 0.1
    fl = 8 + torch.rand(N, 1, generator=g) * 4
    return verts, R, T, skin, pp, fl


def close(a, b, tol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())


def test_golden_vectors_from_reference(dev):
    from lasr_b200 import geom_utils as gu
    T = lambda k: torch.from_numpy(GOLD[k]).to(dev)
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R"), T("lbs_T"), 5, 1, T("lbs_skin")), torch.from_numpy(GOLD["lbs_out"]), 2e-6)
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R"), T("lbs_T"), 5, 1, T("lbs_skin"), tocam=False),
                 torch.from_numpy(GOLD["lbs_out_notocam"]), 2e-6)
    assert close(gu.obj_to_cam(T("lbs_verts"), T("lbs_R")[::5].contiguous(), T("lbs_T")[::5].contiguous(), 1, 1, None),
                 torch.from_numpy(GOLD["rigid_out"]), 2e-6)
    assert close(gu.pinhole_cam(T("cam_v4"), T("cam_pp"), T("cam_fl")), torch.from_numpy(GOLD["pinhole_out"]), 2e-6)


@pytest.mark.parametrize("N,V,B", [(3, 70, 1), (4, 300, 5), (6, 2930, 26)])
def test_obj_to_cam_and_pinhole_forward_backward(dev, N, V, B):
    from lasr_b200 import geom_utils as gu
    verts, R, T, skin, pp, fl = make_inputs(N, V, B)
    cpu = [t.clone().requires_grad_(True) if t is not None else None for t in (verts, R, T, skin, pp, fl)]
    gpu = [t.clone().to(dev).requires_grad_(True) if t is not None else None for t in (verts, R, T, skin, pp, fl)]
    outs = []
    for v_, R_, T_, s_, pp_, fl_ in (cpu, gpu):
        cam = gu.obj_to_cam(v_, R_, T_, B, 1, s_)
        v4 = torch.cat([cam, torch.ones_like(cam[:, :, :1])], -1)
        px = gu.pinhole_cam(v4, pp_, fl_)
        outs.append((cam, px))
    assert close(outs[1][0], outs[0][0], 2e-6) and close(outs[1][1], outs[0][1], 2e-6)
    g = torch.Generator().manual_seed(9)
    gc, gp = torch.randn(N, V, 3, generator=g), torch.randn(N, V, 4, generator=g)
    (outs[0][0] * gc).sum().add((outs[0][1] * gp).sum()).backward()
    (outs[1][0] * gc.to(dev)).sum().add((outs[1][1] * gp.to(dev)).sum()).backward()
    for a, b, nm in zip(gpu, cpu, ("verts", "R", "T", "skin", "pp", "fl")):
        if a is None:
            continue
        assert close(a.grad, b.grad, 5e-5), nm


def test_fused_skin_project_shared_mesh_and_weights(dev):
    """LASR's layout: one mean shape and one set of skinning weights for every frame."""
    from lasr_b200 import geom_utils as gu
    from lasr_b200 import native_geom
    N, V, B = 8, 500, 9
    verts, R, T, skin, pp, fl = make_inputs(N, V, B, seed=3)
    v1, s1 = verts[:1].clone(), skin[:1].clone()
    cpu = [t.clone().requires_grad_(True) for t in (v1, R, T, s1, pp, fl)]
    cam_ref = gu.obj_to_cam(cpu[0].repeat(N, 1, 1), cpu[1], cpu[2], B, 1, cpu[3].repeat(N, 1, 1, 1))
    px_ref = gu.pinhole_cam(torch.cat([cam_ref, torch.ones_like(cam_ref[:, :, :1])], -1), cpu[4], cpu[5])
    gpu = [t.clone().to(dev).requires_grad_(True) for t in (v1, R, T, s1, pp, fl)]
    cam, px, (near, far) = native_geom.skin_project(gpu[0], gpu[1], gpu[2], B, gpu[3], gpu[4], gpu[5].reshape(-1),
                                                    depth_window=True)
    assert close(cam, cam_ref, 2e-6) and close(px, px_ref, 2e-6)
    n_ref, f_ref = gu.depth_window(px_ref[:, :, 2])
    assert abs(float(near) - float(n_ref)) < 1e-5 and abs(float(far) - float(f_ref)) < 1e-5
    g = torch.Generator().manual_seed(4)
    gc, gp = torch.randn(N, V, 3, generator=g), torch.randn(N, V, 4, generator=g)
    ((cam_ref * gc).sum() + (px_ref * gp).sum()).backward()
    ((cam * gc.to(dev)).sum() + (px * gp.to(dev)).sum()).backward()
    for a, b, nm in zip(gpu, cpu, ("verts", "R", "T", "skin", "pp", "fl")):
        assert close(a.grad, b.grad, 5e-5), nm
    # deterministic
    gpu2 = [t.clone().to(dev).requires_grad_(True) for t in (v1, R, T, s1, pp, fl)]
    cam2, px2, _ = native_geom.skin_project(gpu2[0], gpu2[1], gpu2[2], B, gpu2[3], gpu2[4], gpu2[5].reshape(-1),
                                            depth_window=True)
    ((cam2 * gc.to(dev)).sum() + (px2 * gp.to(dev)).sum()).backward()
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(gpu, gpu2))


def test_hypotheses_repeat_of_principal_point(dev):
    """pinhole_cam's pp is per frame-pair and repeated over hypotheses (geom_utils.py:28-29)."""
    from lasr_b200 import geom_utils as gu
    g = torch.Generator().manual_seed(1)
    verts = torch.cat([torch.randn(6, 40, 2, generator=g), 9 + torch.rand(6, 40, 1, generator=g), torch.ones(6, 40, 1)], -1)
    pp, fl = torch.randn(3, 2, generator=g) * 0.1, 9 + torch.rand(6, 1, generator=g)
    assert close(gu.pinhole_cam(verts.to(dev), pp.to(dev), fl.to(dev)), gu.pinhole_cam(verts, pp, fl), 2e-6)


def test_fused_image_losses_match_cpu_expressions(dev):
    """L1-L3 on CUDA (one fused masked-reduction kernel + its backward) == the PyTorch expressions on
    CPU tensors, which tests/test_host_glue.py ties to the loop-for-loop restatement of the reference."""
    from lasr_b200 import losses
    g = torch.Generator().manual_seed(11)
    N, H, S = 5, 2, 96
    mk = lambda *sh: torch.rand(*sh, generator=g)
    mask_pred, fg, tex = mk(N, H, S, S), mk(N, H, S, S), mk(N, H, 3, S, S)
    flow_rd = torch.randn(N, H, S, S, 2, generator=g) * 0.05
    masks = (mk(N, S, S) > 0.5).float()
    occ = torch.randn(N, S, S, generator=g)
    occ[mk(N, S, S) < 0.2] = 0
    bg = mk(N, H, S, S) < 0.4
    bg[1, 0] = True
    flow_obs, imgs = torch.randn(N, 3, S, S, generator=g) * 0.05, mk(N, 3, S, S)

    def run(d):
        t = lambda x: x.clone().to(d)
        mp, fr, tx, fgx = (t(x).requires_grad_(True) for x in (mask_pred, flow_rd, tex, fg))
        lm, sm = losses.mask_loss(mp, t(masks), t(occ))
        lf, sf, _ = losses.flow_loss(fr, t(bg), t(flow_obs), t(occ), t(masks))
        lt, stx = losses.texture_l1_loss(tx, fgx, t(imgs), t(masks), t(occ), 0.7)
        (lm + 2 * lf + 3 * lt).backward()
        return [x.detach().cpu() for x in (lm, lf, lt, sm, sf, stx, mp.grad, fr.grad, tx.grad, fgx.grad)]

    a, b = run(torch.device("cpu")), run(dev)
    for x, y in zip(a, b):
        assert torch.allclose(x, y, rtol=2e-4, atol=1e-7), (x.flatten()[:4], y.flatten()[:4])
    # the single fused call gives the same three numbers
    t = lambda x: x.clone().to(dev)
    lm, lf, lt, _ = losses.image_losses(t(mask_pred), t(masks), t(occ), t(flow_rd), t(bg), t(flow_obs), t(tex), t(fg),
                                        t(imgs), 0.7)
    assert torch.allclose(torch.stack([lm, lf, lt]).cpu(), torch.stack(a[:3]), rtol=2e-4)
    # the scalar step riding the partial-sum launch (lasr_b200_losses_forward_total) == the two separate calls, bit for bit
    from lasr_b200 import _cabi
    c = lambda x: t(x).float().contiguous()
    args = (c(mask_pred), c(flow_rd), t(bg).to(torch.uint8).contiguous(), c(tex), c(fg), c(imgs), c(masks), c(occ), c(flow_obs))
    e = lambda *sh: torch.full(sh, float("nan"), device=dev)
    s1, s2, sub1, sub2, o1, o2 = e(N, H, 7), e(N, H, 7), e(3, N, H), e(3, N, H), e(3), e(4)
    _cabi.losses_forward(*args, s1, N, H, S)
    _cabi.losses_finalize(s1, sub1, o1, N, H, 0.7)
    _cabi.losses_forward(*args, s2, N, H, S, sub2, o2, 0.7)       # + a fourth entry: the sum of the three
    assert torch.equal(s1, s2) and torch.equal(sub1, sub2) and torch.equal(o1, o2[:3]) and bool(torch.isfinite(o2).all())
    assert float(o2[3]) == float((o1[0] + o1[1]) + o1[2])
    # and its adjoint: the gradient of the sum reaches every sum exactly as the same gradient on the three means does
    g1, g2 = e(N, H, 7), e(N, H, 7)
    _cabi.losses_coefficients(s1, torch.tensor([1.5, 1.5, 1.5], device=dev), None, g1, N, H, 0.7)
    _cabi.losses_coefficients(s1, torch.tensor([0.0, 0.0, 0.0, 1.5], device=dev), None, g2, N, H, 0.7)
    assert torch.equal(g1, g2)


@pytest.mark.parametrize("H,Bp,V", [(1, 25, 2930), (2, 7, 642), (1, 1, 100)])
def test_native_skin_weights_match_the_reference_expression(H, Bp, V):
    """S0 kernel (lasr_b200_skin_weights_*) vs the reference's expression (nnutils/mesh_net.py:264-267) run as
    plain PyTorch in float64 on the CPU: weights and the gradients of all three bone parameters."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lasr_b200 import geom_utils as gu
    g = torch.Generator().manual_seed(H * 100 + Bp)
    mean_v = torch.randn(H, V, 3, generator=g) * 0.3
    idx = torch.randint(0, V, (H, Bp), generator=g)
    ctl_ts = torch.stack([mean_v[h, idx[h]] for h in range(H)]).reshape(H * Bp, 3) + 0.01
    ctl_rs = torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(H * Bp, 1) + torch.randn(H * Bp, 4, generator=g) * 0.2
    ctl_rs = ctl_rs * (0.5 + torch.rand(H * Bp, 1, generator=g))          # not unit length
    log_ctl = 1.0 + torch.randn(H * Bp, 3, generator=g) * 0.5
    gout = torch.randn(H, Bp, V, 1, generator=g)

    def run(dev, dt):
        leaves = [x.clone().to(dev, dt).requires_grad_(True) for x in (ctl_ts, ctl_rs, log_ctl)]
        skin = gu.skin_weights(leaves[0], leaves[1], leaves[2], mean_v.to(dev, dt), H)
        (skin * gout.to(dev, dt)).sum().backward()
        return skin.detach().cpu().double(), [x.grad.cpu().double() for x in leaves]

    s_gpu, g_gpu = run(torch.device("cuda", 0), torch.float32)
    s_ref, g_ref = run(torch.device("cpu"), torch.float64)
    assert s_gpu.shape == (H, Bp, V, 1)
    assert float((s_gpu - s_ref).abs().max()) < 2e-6
    assert float((s_gpu.sum(1) - 1).abs().max()) < 1e-5
    for a, b, nm in zip(g_gpu, g_ref, ("ctl_ts", "ctl_rs", "log_ctl")):
        assert (a - b).norm() <= 2e-5 * b.norm() + 1e-7, (nm, float((a - b).norm()), float(b.norm()))
    # deterministic
    s2, g2 = run(torch.device("cuda", 0), torch.float32)
    assert torch.equal(s2, s_gpu) and all(torch.equal(x, y) for x, y in zip(g2, g_gpu))


@pytest.mark.parametrize("H,B,V", [(1, 4, 300), (2, 3, 77), (1, 26, 2930)])
def test_bone_composition_inside_the_lbs_kernel_equals_the_pytorch_composition(dev, H, B, V):
    """lasr_b200_lbs_pose_forward / backward (skin_project(rest_ts=...)): the part transforms are composed about their rest
    positions while the frame's bones are staged in shared memory (mesh_net.py:278-281: T' = -R rest + T + rest, R' = R^T).
    Against the same kernels fed the composition done in PyTorch (RenderHotPath.bone_transforms): positions, depth window
    and the gradient of every input, the joint centres included."""
    from lasr_b200 import hotpath, native_geom
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    g = torch.Generator().manual_seed(B * 7 + H)
    n2 = 4
    N = n2 * H
    leaves = dict(
        verts=torch.randn(N, V, 3, generator=g) * 0.3,
        R=quaternion_to_rotation_matrix(torch.randn(N * B, 4, generator=g)),
        T=torch.randn(N * B, 3, generator=g) * 0.05,
        rest=torch.randn(H * (B - 1), 3, generator=g) * 0.3,
        skin=torch.softmax(torch.randn(N, B - 1, V, 1, generator=g), 1),
        pp=torch.randn(N, 2, generator=g) * 0.05, fl=9.5 + torch.rand(N, generator=g))
    leaves["T"].view(N, B, 3)[:, 0, 2] += 8.0
    gc, gp = torch.randn(N, V, 3, generator=g).to(dev), torch.randn(N, V, 4, generator=g).to(dev)

    def run(in_kernel):
        L = {k: v.clone().to(dev).requires_grad_(True) for k, v in leaves.items()}
        if in_kernel:
            cam, px, (near, far) = native_geom.skin_project(L["verts"], L["R"], L["T"], B, L["skin"], L["pp"], L["fl"],
                                                            depth_window=True, rest_ts=L["rest"], n_hypo=H)
        else:
            m = hotpath.RenderHotPath(hotpath.HotPathOptions(n_hypo=H, n_bones=B))
            bones = dict(rest_ts=L["rest"])
            rest = bones["rest_ts"][:, None, :, None].repeat(n2, 1, 1, 1).view(-1, B - 1, 3, 1)
            R4, T4 = L["R"].view(-1, B, 3, 3), L["T"].view(-1, B, 3, 1)
            Tm = torch.cat([T4[:, :1], -R4[:, 1:].matmul(rest) + T4[:, 1:] + rest], 1).view(-1, 3)
            Rm = torch.cat([R4[:, :1], R4[:, 1:].permute(0, 1, 3, 2)], 1).reshape(-1, 3, 3)
            cam, px, (near, far) = native_geom.skin_project(L["verts"], Rm, Tm, B, L["skin"], L["pp"], L["fl"],
                                                            depth_window=True)
        torch.autograd.backward([cam, px], [gc, gp])
        return cam.detach(), px.detach(), float(near), float(far), {k: v.grad for k, v in L.items()}

    a, b = run(True), run(False)
    assert float((a[0] - b[0]).abs().max()) < 2e-6 and float((a[1] - b[1]).abs().max()) < 2e-5
    assert abs(a[2] - b[2]) < 1e-5 and abs(a[3] - b[3]) < 1e-5
    for k in leaves:
        err = float((a[4][k] - b[4][k]).norm() / b[4][k].norm())
        assert err < 2e-5, (k, err)


def test_frame_sums_of_shared_parameter_gradients_in_one_launch():
    """lasr_b200_lbs_sum_frames: per-frame gradients of the shared mesh, the shared skinning weights and the joint centres
    (frames of one hypothesis) summed in one launch == the three PyTorch reductions it replaces."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lasr_b200 import _cabi
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    for N, V, B, H in ((8, 642, 26, 1), (6, 100, 4, 2), (64, 2930, 26, 1), (3, 7, 2, 3)):
        gv = torch.randn(N, V, 3, generator=g).to(dev)
        gs = torch.randn(N, B - 1, V, generator=g).to(dev)
        gr = torch.randn(N, B - 1, 3, generator=g).to(dev)
        sv, ss, sr = _cabi.lbs_sum_frames(gv, gs, gr, N, V, B, H)
        tol = dict(rtol=1e-5, atol=1e-5)
        assert torch.allclose(sv, gv.double().sum(0).float(), **tol)
        assert torch.allclose(ss, gs.double().sum(0).float(), **tol)
        assert torch.allclose(sr.reshape(H * (B - 1), 3), gr.double().view(N // H, H * (B - 1), 3).sum(0).float(), **tol)
        only_v = _cabi.lbs_sum_frames(gv, None, None, N, V, B, H)
        assert only_v[1] is None and only_v[2] is None and torch.equal(only_v[0], sv)
        only_r = _cabi.lbs_sum_frames(None, None, gr, N, V, B, H)
        assert only_r[0] is None and torch.equal(only_r[2], sr)
