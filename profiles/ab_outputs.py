"""Dump forward images + gradients of a stress set (all tests/cases.py cases, plus triangle soups rich in
slivers, huge and sub-pixel faces at several sigmas) to an .npz -- run once per library build
(LASR_B200_LIB=...) and compare the files bit for bit: an optimisation that only culls work must not change
a single output value.  usage: python profiles/ab_outputs.py out.npz ; python profiles/ab_outputs.py --compare a.npz b.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    worst = {"img": 0.0, "gf": 0.0, "gt": 0.0}
    for k in bad:
        d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64))
        scale = max(float(np.nanmax(np.abs(a[k]))), 1e-30)
        worst[k.rsplit("/", 1)[1]] = max(worst[k.rsplit("/", 1)[1]], float(np.nanmax(d)) / scale)
        print("DIFF", k, "max", np.nanmax(d), "count", int((d > 0).sum()), "max/scale", float(np.nanmax(d)) / scale)
    print(f"{len(a.files)} arrays compared, {len(bad)} differ; images differing: {sum(k.endswith('/img') for k in bad)}; "
          f"worst max|d|/max|x|: img {worst['img']:.2e} gf {worst['gf']:.2e} gt {worst['gt']:.2e}")
    sys.exit(1 if bad else 0)

import torch
import cases
import lasr_b200.soft_renderer as sr

dev = torch.device("cuda", 0)
out = {}


def run(tag, c):
    fv = torch.from_numpy(c["face_vertices"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(c["textures"]).to(dev).requires_grad_(True)
    img = sr.functional.soft_rasterize(fv, tex, **c["settings"])
    img.backward(torch.from_numpy(c["grad_soft_colors"]).to(dev))
    out[tag + "/img"], out[tag + "/gf"], out[tag + "/gt"] = img.detach().cpu().numpy(), fv.grad.cpu().numpy(), tex.grad.cpu().numpy()


for name, mk in {**cases.GOLDEN_CASES, **cases.LIVE_CASES}.items():
    run(name, mk())

rng = np.random.RandomState(7)
for k, (F, S, sigma, kind) in enumerate([(400, 96, 1e-4, "sliver"), (400, 96, 1e-5, "sliver"), (300, 128, 1e-3, "mixed"),
                                         (60, 64, 1e-4, "huge"), (2000, 160, 1e-4, "tiny"), (500, 96, 3e-4, "mixed")]):
    N = 2
    c0 = rng.uniform(-1.1, 1.1, (N, F, 1, 2))
    if kind == "sliver":      # long thin triangles, many nearly degenerate, some exactly degenerate
        d = rng.normal(size=(N, F, 1, 2)); d /= np.linalg.norm(d, axis=-1, keepdims=True)
        L = rng.uniform(0.05, 1.0, (N, F, 1, 1)); w = 10.0 ** rng.uniform(-7, -2, (N, F, 1, 1))
        nrm = np.concatenate([-d[..., 1:], d[..., :1]], -1)
        xy = np.concatenate([c0, c0 + d * L, c0 + d * L * rng.uniform(0.2, 0.8, (N, F, 1, 1)) + nrm * w], 2)
        xy[:, ::37, 2] = xy[:, ::37, 1]                      # exactly degenerate
    elif kind == "huge":
        xy = c0 + rng.uniform(-1.5, 1.5, (N, F, 3, 2))
    elif kind == "tiny":
        xy = c0 + rng.uniform(-0.01, 0.01, (N, F, 3, 2))
    else:
        xy = c0 + rng.uniform(-0.3, 0.3, (N, F, 3, 2)) * 10.0 ** rng.uniform(-2, 0, (N, F, 1, 1))
    z = rng.uniform(2.0, 8.0, (N, F, 3, 1))
    fv = np.concatenate([xy, z], -1).astype(np.float32)
    s = dict(image_size=S, background_color=[0.2, 0.4, 0.6], near=1.0, far=10.0, fill_back=True, eps=1e-3, sigma_val=sigma,
             dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax", aggr_func_alpha="prod",
             texture_type="vertex")
    c = dict(face_vertices=fv, textures=rng.rand(N, F, 3, 3).astype(np.float32), settings=s,
             grad_soft_colors=rng.normal(size=(N, 4, S, S)).astype(np.float32))
    run(f"soup{k}_{kind}", c)
    if kind in ("sliver", "mixed"):                          # the generic (IEEE) path and the other distance modes
        for dist in ("barycentric", "hard"):
            c2 = dict(c, settings=dict(s, dist_func=dist, aggr_func_alpha="sum"))
            run(f"soup{k}_{kind}_{dist}", c2)
np.savez_compressed(sys.argv[1], **out)
print("wrote", sys.argv[1], len(out), "arrays")
