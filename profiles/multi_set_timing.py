"""Kernel-level cost of carrying K colour sets through one geometry pass (render_indexed, Spot mesh,
256^2, 64 frames) against K separate single-set renders.  usage: python profiles/multi_set_timing.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lasr_b200 import synth
import lasr_b200.soft_renderer as sr

dev = torch.device("cuda", 0)
N, S = 64, 256
v_np, f_np = synth.load_spot()
px, cam, near, far = synth.render_inputs(v_np, f_np, N)     # vertices as the rasterizer sees them in LASR
verts = torch.from_numpy(px).to(dev)
tri = torch.from_numpy(f_np).int().to(dev)
col = torch.from_numpy(synth.vertex_colors(len(v_np)))[None].repeat(N, 1, 1).to(dev)
kw = dict(image_size=S, near=near, far=far, fill_back=True, sigma_val=1e-4, gamma_val=1e-2)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for K in (1, 2, 3):
    c = col if K == 1 else torch.stack([col] * K)
    bg = [0.0] * (3 * K)
    v = verts.clone().requires_grad_(True)
    cc = c.clone().requires_grad_(True)
    g = torch.randn(N, 3 * K + 1, S, S, device=dev)
    fwd = timed(lambda: sr.functional.render_indexed(v.detach(), cc.detach(), tri, background_color=bg, **kw))

    def both():
        v.grad = cc.grad = None
        sr.functional.render_indexed(v, cc, tri, background_color=bg, **kw).backward(g)
    fb = timed(both)
    print(f"K={K}: forward {fwd:.3f} ms, forward+backward {fb:.3f} ms, backward ~{fb - fwd:.3f} ms "
          f"({N / fb * 1e3:.0f} frames/s; per colour set {fb / K:.3f} ms)")
