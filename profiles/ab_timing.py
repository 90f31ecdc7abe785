"""A/B kernel timing of one library build (select it with LASR_B200_LIB=...): config A (Spot 5856 F, 256^2, 64 frames)
through SoftRasterizeFunction (forward, backward) and through render_indexed with K = 1 and K = 3 colour sets.
usage: LASR_B200_LIB=path/to/variant.so python profiles/ab_timing.py [tag] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from lasr_b200 import synth
import lasr_b200.soft_renderer as sr

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("LASR_B200_LIB", "default"))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda", 0)


def timed(fn, reps=reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):     # best of 3 batches: A/B differences of a few % are what we are after
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


res = {}
for mesh, N, S in (("spot", 64, 256), ("ico3", 64, 256)):
    c = cases.lasr_case(mesh, N, S)
    fv = torch.from_numpy(c["face_vertices"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(c["textures"]).to(dev).requires_grad_(True)
    g = torch.from_numpy(c["grad_soft_colors"]).to(dev)
    fwd = timed(lambda: sr.functional.soft_rasterize(fv.detach(), tex.detach(), **c["settings"]))

    def both():
        fv.grad = tex.grad = None
        sr.functional.soft_rasterize(fv, tex, **c["settings"]).backward(g)
    fb = timed(both)
    res[f"{mesh}_fwd"], res[f"{mesh}_fb"] = fwd, fb

N, S = 64, 256
v_np, f_np = synth.load_spot()
px, cam, near, far = synth.render_inputs(v_np, f_np, N)
verts = torch.from_numpy(px).to(dev)
tri = torch.from_numpy(f_np).int().to(dev)
col = torch.from_numpy(synth.vertex_colors(len(v_np)))[None].repeat(N, 1, 1).to(dev)
kw = dict(image_size=S, near=near, far=far, fill_back=True, sigma_val=1e-4, gamma_val=1e-2)
for K in (1, 3):
    cc0 = col if K == 1 else torch.stack([col] * K)
    bg = [0.0] * (3 * K)
    v = verts.clone().requires_grad_(True)
    cc = cc0.clone().requires_grad_(True)
    g = torch.randn(N, 3 * K + 1, S, S, device=dev)
    res[f"idxK{K}_fwd"] = timed(lambda: sr.functional.render_indexed(v.detach(), cc.detach(), tri, background_color=bg, **kw))

    def both():
        v.grad = cc.grad = None
        sr.functional.render_indexed(v, cc, tri, background_color=bg, **kw).backward(g)
    res[f"idxK{K}_fb"] = timed(both)
print(f"AB {tag}: " + "  ".join(f"{k} {v:.3f}" for k, v in res.items()))
