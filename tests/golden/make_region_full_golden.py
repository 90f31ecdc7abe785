"""Full-size golden for the WHOLE render region: the reference's own ``LASR.forward`` lines (nnutils/mesh_net.py:259-447
and its ``render_flow_soft_2``, :75-104) executed with autograd at LASR's articulated shape -- Spot (2930 V / 5856 F),
26 bones, 256 x 256, H = 1, two frame pairs -- on top of the drop-in ``soft_renderer`` package with the CPU oracle as the
rasterizer (forward AND backward), exactly as tests/golden/make_region_losses_golden.py does for its small cases.

What is stored (tests/golden/region_full_ref.npz):
  * the three losses and their sum;
  * d(total_loss)/d(every input): mean shape, vertex colours, the pose network's outputs (quat / trans / depth /
    ppoint / scale) and the four bone parameter sets -- the reference's autograd through its own lines, the oracle's
    analytic backward inside;
  * the rendered silhouette (16-bit quantised: enough to count pixels that differ by more than 1e-3) and flow;
  * the posed vertices the reference's ``obj_to_cam`` / ``pinhole_cam`` produced (camera space and projected), so a GPU
    test can hand the rasterizer bit-identical geometry and hold the gradients to 1e-3 without the isolated pixel
    flips a last-bit difference in the skinning arithmetic causes at edge-on faces.
Inputs are regenerated from the seed by ``inputs(seed)`` below (imported by the test), not stored.

Build container only (needs /root/reference), no GPU.   python tests/golden/make_region_full_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "region_full_ref.npz")
BS, H, B, S, SEED = 2, 1, 26, 256, 11


def inputs(seed=SEED, bs=BS, n_hypo=H, n_bones=B, img_size=S):
    """Seeded inputs of the region in the reference's own variable names (what LASR.forward holds at line 259)."""
    from lasr_b200 import synth
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    v, f = synth.load_spot()
    V = len(v)
    g = torch.Generator().manual_seed(seed)
    n2, N = 2 * bs, 2 * bs * n_hypo
    Rb, Tb, _, _ = synth.spot_cameras(N, seed)                           # the render_syn.py orbit for the body
    q = torch.randn(N * n_bones, 4, generator=g) * 0.06
    q[:, 3] = 1
    R = quaternion_to_rotation_matrix(q).view(N, n_bones, 3, 3)
    R[:, 0] = torch.from_numpy(Rb).permute(0, 2, 1)                      # line 259 transposes: hand over R^T
    quat = R.reshape(N * n_bones, 9).contiguous()
    trans = torch.randn(N * n_bones, 2, generator=g) * 0.01
    depth = torch.randn(N * n_bones, 1, generator=g) * 0.01
    trans.view(N, n_bones, 2)[:, 0] = torch.from_numpy(Tb[:, 0, :2])
    depth.view(N, n_bones, 1)[:, 0] = torch.from_numpy(Tb[:, 0, 2:])
    pred_v = torch.from_numpy(v)[None].repeat(N, 1, 1)
    tex = torch.from_numpy(synth.vertex_colors(V, 0))[None].repeat(N, 1, 1)
    faces = torch.from_numpy(f)[None].repeat(n2, 1, 1)
    ppoint = torch.randn(n2, 2, generator=g) * 0.03
    scale = 9.5 + torch.rand(n2, n_hypo, generator=g)
    idx = torch.randperm(V, generator=g)[:n_hypo * (n_bones - 1)]
    ctl = torch.from_numpy(v)[idx]
    occ = torch.randn(n2, img_size, img_size, generator=g)
    occ[torch.rand(n2, img_size, img_size, generator=g) < 0.1] = 0
    # observations: a coarse disc silhouette (so the masked means see object and background) and noise images / flow
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, img_size), torch.linspace(-1, 1, img_size), indexing="ij")
    masks = ((xx ** 2 + yy ** 2) < 0.45)[None].repeat(n2, 1, 1).float()
    return dict(quat=quat, trans=trans, depth=depth, pred_v=pred_v, tex=tex, faces=faces, ppoint=ppoint, scale=scale,
                ctl_ts=ctl.clone(), rest_ts=ctl.clone() + 0.005,
                ctl_rs=torch.tensor([[0.0, 0.0, 0.0, 1.0]]).repeat(n_hypo * (n_bones - 1), 1)
                + torch.randn(n_hypo * (n_bones - 1), 4, generator=g) * 0.03,
                log_ctl=1.0 + torch.randn(n_hypo * (n_bones - 1), 3, generator=g) * 0.3,
                masks=masks, imgs=torch.rand(n2, 3, img_size, img_size, generator=g),
                flow=torch.randn(n2, 3, img_size, img_size, generator=g) * 0.03, occ=occ)


LEAVES = ("quat", "trans", "depth", "pred_v", "tex", "ppoint", "scale", "ctl_ts", "rest_ts", "ctl_rs", "log_ctl")


def main():
    import make_region_losses_golden as small
    env = small.setup()
    sr = env.sr
    mk = lambda: sr.SoftRenderer(image_size=S, sigma_val=1e-4, gamma_val=1e-2, camera_mode="look_at", perspective=False,
                                 aggr_func_rgb="softmax", light_mode="vertex", light_intensity_ambient=1.0,
                                 light_intensity_directionals=0.0)       # nnutils/mesh_net.py:136-145
    x = inputs()
    leaves = {k: x[k].clone().requires_grad_(True) for k in LEAVES}
    posed = {}

    def obj_to_cam(verts, *a, **k):                                       # record what reaches the rasterizer
        out = env.gu.obj_to_cam(verts, *a, **k)
        if verts.shape[1] == x["pred_v"].shape[1] and k.get("tocam", True):
            posed["cam"] = out.detach().clone()
        return out

    def pinhole_cam(verts, *a, **k):
        out = env.gu.pinhole_cam(verts, *a, **k)
        if verts.shape[1] == x["pred_v"].shape[1]:
            posed["px"] = out.detach().clone()
        return out

    self = types.SimpleNamespace(
        renderer_soft=mk(), renderer_softflf=mk(), renderer_softflb=mk(), renderer_softtex=mk(), optim_idx=0, iters=1,
        texture_type="vertex", masks=x["masks"], imgs=x["imgs"], flow=x["flow"], occ=x["occ"],
        ptex_loss=types.SimpleNamespace(forward_pair=lambda a, b: torch.zeros(a.shape[0])),
        ctl_ts=leaves["ctl_ts"], rest_ts=leaves["rest_ts"], ctl_rs=leaves["ctl_rs"], log_ctl=leaves["log_ctl"])
    opts = types.SimpleNamespace(n_hypo=H, n_bones=B, sigval=1e-4, opt_tex="yes", img_size=S, l1tex_wt=1.0)
    ns = dict(self=self, opts=opts, torch=torch, np=np, sr=sr, local_batch_size=BS,
              quat=leaves["quat"] * 1.0,            # the lines write into `quat` through a view: not a leaf
              trans=leaves["trans"], depth=leaves["depth"], pred_v=leaves["pred_v"], tex=leaves["tex"], faces=x["faces"],
              ppoint=leaves["ppoint"], scale=leaves["scale"], obj_to_cam=obj_to_cam, pinhole_cam=pinhole_cam,
              citylabs=[[10 * k, 20, 30] for k in range(64)],
              kornia=types.SimpleNamespace(quaternion_to_rotation_matrix=env.q2r))
    exec(env.flow_fn, ns)
    exec(env.region, ns)
    self.total_loss.backward()
    out = {"meta": np.asarray([BS, H, B, S, SEED], np.int64)}
    for k in ("mask_loss", "flow_rd_loss", "texture_loss", "total_loss"):
        out[f"out/{k}"] = np.asarray(getattr(self, k).detach().numpy())
    for k in LEAVES:
        out[f"grad/{k}"] = leaves[k].grad.numpy().astype(np.float32)
    out["out/mask_pred_u16"] = np.round(self.mask_pred.detach().numpy().clip(0, 1) * 65535.0).astype(np.uint16)
    out["out/flow_rd_f16"] = self.flow_rd.detach().numpy().astype(np.float16)
    out["out/bgmask"] = np.packbits(self.bgmask.detach().numpy())
    out["posed/cam"] = posed["cam"].numpy()
    out["posed/px"] = posed["px"][:, :, :3].numpy()
    print({k: round(float(out[f"out/{k}"]), 6) for k in ("mask_loss", "flow_rd_loss", "texture_loss", "total_loss")})
    print({k: float(np.linalg.norm(out[f"grad/{k}"])) for k in LEAVES})
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
