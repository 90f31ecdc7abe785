"""Mesh regularisers on one GPU: the sparse kernels behind lasr_b200.loss_utils against
the reference's dense formulation restated in PyTorch (third_party/ext_nnutils/loss_utils.py:34-152,
nnutils/loss_utils.py:29-64), forward + backward, Spot mesh, LASR's call shapes (mesh_net.py:455-456,496).
usage: python profiles/regularizer_timing.py [frames]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lasr_b200 import loss_utils as LU, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
v, f = synth.load_spot()
V = len(v)
vt, ft = torch.from_numpy(v), torch.from_numpy(f)
g = torch.Generator().manual_seed(0)
x = (vt[None] + torch.randn(N, V, 3, generator=g) * 0.01).to(dev)
dx = (vt[None] * 1.05 + torch.randn(N, V, 3, generator=g) * 0.01).to(dev)
lap, flat, arap = LU.LaplacianLoss(vt, ft).to(dev), LU.FlattenLoss(ft).to(dev), LU.ARAPLoss(vt, ft).to(dev)
E2, E = arap.n_edges, flat.edges.shape[0]

# the reference's dense operands
A = torch.zeros(V, V, device=dev)
fl = ft.long().to(dev)
for a, b in ((0, 1), (1, 2), (2, 0)):
    A[fl[:, a], fl[:, b]] = 1
    A[fl[:, b], fl[:, a]] = 1
L = -A.clone()
L[torch.arange(V), torch.arange(V)] = A.sum(1)
L = L / L.diagonal()[:, None]


def dense_laplacian(x):
    return torch.matmul(L, x).pow(2).sum((1, 2))


def dense_arap(dx, x):
    diffx = torch.zeros(x.shape[0], V, V, device=dev)
    diffdx = torch.zeros(x.shape[0], V, V, device=dev)
    for i in range(3):
        diffdx += (A.matmul(torch.diag_embed(dx[:, :, i])) - dx[:, :, i:i + 1]).pow(2)
        diffx += (A.matmul(torch.diag_embed(x[:, :, i])) - x[:, :, i:i + 1]).pow(2)
    diff = (diffx - diffdx).abs()
    return torch.stack([diff[i][A.bool()].mean() for i in range(x.shape[0])])


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fb(fn, *leaves):
    def run():
        ls = [t.detach().requires_grad_(True) for t in leaves]
        fn(*ls).sum().backward()
        return ls
    return run


rows = []
for name, ours, dense, leaves, nbytes in (
        ("LaplacianLoss", lap, dense_laplacian, (x,), N * V * 12 * 4 + (V + E2) * 4),
        ("FlattenLoss", flat, None, (x,), N * V * 12 * 2 + N * E * 48 * 2 + E * 16),
        ("ARAPLoss", arap, dense_arap, (dx[:N // 2], x[:N // 2]), (N // 2) * V * 12 * 4 + (V + E2) * 4)):
    t = timed(fb(ours, *leaves), 30)
    line = f"{name:14s} frames={leaves[0].shape[0]:3d}  sparse kernels fwd+bwd {t * 1e3:8.1f} us  ({nbytes / t / 1e6:7.1f} GB/s of compulsory bytes)"
    if dense is not None:
        try:
            td = timed(fb(dense, *leaves), 2)
            a = ours(*leaves); b = dense(*leaves)
            line += f"   dense formulation {td:9.2f} ms  ({td / t:8.0f}x)   max rel diff {float(((a - b).abs() / b.abs()).max()):.1e}"
        except torch.cuda.OutOfMemoryError:
            line += "   dense formulation: out of memory"
    print(line)
print(f"V={V} directed edges={E2} dihedral edges={E}; peak allocated {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
