"""Where the end-to-end step of bench.py spends its time (CUDA events around each phase)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from lasr_b200 import _cabi

S, N = 256, 64
dev = torch.device("cuda", 0)
wl = bench.build_workload(N, S, 0)
arm = bench.OursArm(wl, S, dev)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_verts, h_colors = pin(wl["verts_px"]), pin(wl["colors"])
h_faces = pin(np.repeat(wl["faces"][None], N, 0).astype(np.int32))
h_target = pin(np.random.RandomState(7).rand(N, 4, S, S).astype(np.float32))
names = ["h2d", "render_fwd", "loss_fwd", "backward", "reduce+d2h"]
acc = {k: [] for k in names}
wall = []
for it in range(15):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record()
    verts = h_verts.to(dev, non_blocking=True).requires_grad_(True)
    colors = h_colors.to(dev, non_blocking=True).requires_grad_(True)
    faces = h_faces.to(dev, non_blocking=True)
    target = h_target.to(dev, non_blocking=True)
    ev[1].record()
    img = arm.render(verts, faces, colors)
    ev[2].record()
    loss = 0.5 * (img - target).pow(2).mean()
    ev[3].record()
    loss.backward()
    ev[4].record()
    gv = verts.grad.sum(0); gc = colors.grad.sum(0)
    out = (loss.detach().cpu(), gv.cpu(), gc.cpu())
    ev[5].record()
    torch.cuda.synchronize(); wall.append((time.perf_counter() - t0) * 1e3)
    if it >= 5:
        for i, k in enumerate(names):
            acc[k].append(ev[i].elapsed_time(ev[i + 1]))
print({k: round(float(np.mean(v)), 3) for k, v in acc.items()}, "sum", round(sum(float(np.mean(v)) for v in acc.values()), 3),
      "wall", round(float(np.mean(wall[5:])), 3))
