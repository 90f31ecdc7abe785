"""Chamfer nearest neighbours: ours vs the reference's kernels (oracle/_ref/chamfer_3D*.so) at LASR's call shape
(one cloud against its mirror image, nnutils/mesh_net.py:474-477) and at evaluation size.
usage: python profiles/aux_timing.py"""
import glob, importlib.util, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lasr_b200.chamfer3D import dist_chamfer_3D as C


def ref_ext(name):
    c = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + "*.so"))
    if not c:
        return None
    spec = importlib.util.spec_from_file_location(name, c[0]); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ref = ref_ext("chamfer_3D")
ours = C.chamfer_3DDist()
g = torch.Generator().manual_seed(0)
for B, n in ((1, 642), (1, 2930), (1, 10242), (1, 100000), (8, 10242)):
    a = torch.randn(B, n, 3, generator=g).cuda(); b = a * torch.tensor([-1.0, 1, 1]).cuda()
    from lasr_b200 import _cabi
    od1 = torch.empty(B, n).cuda(); od2 = torch.empty(B, n).cuda()
    oi1 = torch.empty(B, n, dtype=torch.int32).cuda(); oi2 = torch.empty(B, n, dtype=torch.int32).cuda()
    # both timed at the extension boundary with preallocated outputs (forward of both directions)
    t = timed(lambda: _cabi.native_call("chamfer_forward", a, b, od1, od2, oi1, oi2, B, n, n), 20 if n > 50000 else 50)
    line = f"chamfer B={B} n=m={n}: ours {t * 1e3:.1f} us"
    if ref is not None:
        d1 = torch.zeros(B, n).cuda(); d2 = torch.zeros(B, n).cuda()
        i1 = torch.zeros(B, n, dtype=torch.int32).cuda(); i2 = torch.zeros(B, n, dtype=torch.int32).cuda()
        tr = timed(lambda: ref.forward(a, b, d1, d2, i1, i2), 5 if n > 50000 else 20)
        o = ours(a, b)
        same = bool(torch.equal(o[2], i1) and torch.equal(o[0], d1))
        line += f", reference kernels {tr * 1e3:.1f} us ({tr / t:.1f}x), identical output: {same}"
    print(line)
