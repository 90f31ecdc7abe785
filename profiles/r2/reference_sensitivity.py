"""How sensitive is the REFERENCE's own arithmetic to a last-bit change of its inputs?  Runs the whole region on CPU tensors
(RenderHotPath's host path with the CPU oracle as the rasterizer: reproduces the reference's LASR.forward lines to 1e-6,
tests/test_host_glue.py) on the inputs of tests/golden/region_full_ref.npz, once as they are and once with the mean shape
moved by ONE float32 ulp per coordinate (random sign), and reports how many silhouette pixels change by more than 1e-3
and how far every gradient moves.  This is the noise floor any end-to-end comparison of two implementations of the
skinning arithmetic sits on (edge-on sliver faces flip between inside and outside)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_host_glue import _full_region_inputs, _OracleRasterize
from lasr_b200 import hotpath
import lasr_b200.soft_renderer.rasterizer as rast


def fake(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100, fill_back=True,
         eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
         aggr_func_alpha="prod", texture_type="surface"):
    return _OracleRasterize.apply(face_vertices, textures, dict(
        image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
        eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
        aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))


rast.srf.soft_rasterize = fake
x, g, (bs, H, B, S) = _full_region_inputs()
names = ("quat", "trans", "depth", "pred_v", "tex", "ppoint", "scale", "ctl_ts", "rest_ts", "ctl_rs", "log_ctl")


def run(pred_v):
    L = {k: x[k].clone().requires_grad_(True) for k in names}
    L["pred_v"] = pred_v.clone().requires_grad_(True)
    Rmat = L["quat"].view(-1, 3, 3).permute(0, 2, 1).contiguous()
    Tmat = torch.cat([L["trans"], L["depth"]], 1)
    m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B, l1tex_wt=1.0))
    out = m(L["pred_v"], L["tex"], x["faces"], Rmat, Tmat, L["ppoint"], L["scale"],
            dict(masks=x["masks"], imgs=x["imgs"], flow=x["flow"], occ=x["occ"]),
            {k: L[k] for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")})
    out["total"].backward()
    return out["mask_pred"].detach().numpy(), {k: L[k].grad for k in names}, float(out["total"])


base = run(x["pred_v"])
rng = np.random.RandomState(0)
v = x["pred_v"].numpy()
sign = rng.choice([-1.0, 1.0], size=v.shape).astype(np.float32)
moved = np.nextafter(v, v + sign)                     # one ulp, random direction
pert = run(torch.from_numpy(moved))
d = np.abs(base[0] - pert[0])
print(json.dumps({"perturbation": "mean shape moved by one float32 ulp per coordinate",
                  "flip_pixels": int((d > 1e-3).sum()), "flip_frac": float((d > 1e-3).mean()), "max_sil_change": float(d.max()),
                  "loss_change": abs(base[2] - pert[2]) / abs(base[2]),
                  "grad_change": {k: float((base[1][k] - pert[1][k]).norm() / base[1][k].norm()) for k in names}}))
