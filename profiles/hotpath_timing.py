"""Throughput of the whole render hot path (lasr_b200.hotpath.RenderHotPath: skin -> LBS -> projection ->
flow fw/bw renders + texture render -> three losses, forward + backward) on one GPU.
usage: python profiles/hotpath_timing.py [bones] [pairs] [graph]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lasr_b200 import hotpath, synth
from lasr_b200.geom_utils import quaternion_to_rotation_matrix
from lasr_b200.pipeline import GraphedStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 26
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
use_graph = (sys.argv[3] == "graph") if len(sys.argv) > 3 else False
use_ref = (sys.argv[3] == "reference") if len(sys.argv) > 3 else False
H, S = 1, 256
dev = torch.device("cuda", 0)
inp, leaves = synth.hot_path_inputs(bs, B, S, dev, n_hypo=H)
faces, batch, bones = inp["faces"], inp["batch"], inp["bones"]
n2 = 2 * bs
N = n2 * H
fused = os.environ.get("FUSED", "1") != "0"            # one three-set render + fused head (default) vs three renders
report_flow = os.environ.get("REPORT_FLOW", "1") != "0"
m = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=S, n_hypo=H, n_bones=B, fused=fused,
                                                 report_flow=report_flow)).to(dev)

def step(*xs):
    xs = [x.detach().requires_grad_(True) for x in xs]
    bn = dict(zip(bones.keys(), xs[6:])) if bones else None
    out = m(xs[0], xs[1], faces, xs[2], xs[3], xs[4], xs[5], batch, bn)
    out["total"].backward()
    return (out["total"].detach(),) + tuple(x.grad for x in xs)

if use_ref:
    from oracle import ref_gpu
    from lasr_b200 import geom_utils as gu
    eye_t = torch.tensor(m.renderer_softtex.transform.transformer._eye, device=dev)

    def step(*xs):          # the same region restated on the reference's kernels and Python (loops, syncs)
        xs = [x.detach().requires_grad_(True) for x in xs]
        Rm, Tm, skin = xs[2], xs[3], None
        if bones:
            bn = dict(zip(bones.keys(), xs[6:]))
            skin = gu.skin_weights(bn["ctl_ts"], bn["ctl_rs"], bn["log_ctl"], xs[0].view(n2, H, -1, 3)[0].detach(), H).repeat(n2, 1, 1, 1)
            rest = bn["rest_ts"][:, None, :, None].repeat(n2, 1, 1, 1).view(-1, B - 1, 3, 1)
            R4, T4 = Rm.view(-1, B, 3, 3), Tm.view(-1, B, 3, 1)
            Tm = torch.cat([T4[:, :1], -R4[:, 1:].matmul(rest) + T4[:, 1:] + rest], 1).view(-1, 3)
            Rm = torch.cat([R4[:, :1], R4[:, 1:].permute(0, 1, 3, 2)], 1).view(-1, 3, 3)
        total, *_ = ref_gpu.lasr_region_reference(xs[0], xs[1], faces, Rm, Tm, xs[4], xs[5], batch, H, B, S, skin, eye_t)
        total.backward()
        return (total.detach(),) + tuple(x.grad for x in xs)

run = GraphedStep(step, leaves) if use_graph else (lambda: step(*leaves))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 15 if not use_ref else 3
t0 = time.perf_counter(); e0.record()
for _ in range(K):
    outs = run()
e1.record(); torch.cuda.synchronize()
ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / K
print(f"hotpath B={B} pairs={bs} frames={N} graph={use_graph} reference={use_ref}: {ms:.3f} ms/step -> {N/ms*1e3:.0f} frames/s  loss={float(outs[0]):.5f}")
