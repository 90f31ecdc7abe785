# quick look: ours vs REF-GPU timing, config A (Spot 5856F, 256^2), N frames
import sys, os, time, numpy as np, torch
ROOT=os.environ.get("GRAFT_REPO_ROOT","/root/repo"); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,"tests"))
import cases
from oracle import ref_gpu
from lasr_b200.soft_renderer import functional as srf
N=int(sys.argv[1]) if len(sys.argv)>1 else 64
c=cases.lasr_case("spot",N,256)
fv=torch.from_numpy(c["face_vertices"]).cuda().requires_grad_(True); tex=torch.from_numpy(c["textures"]).cuda().requires_grad_(True)
g=torch.from_numpy(c["grad_soft_colors"]).cuda()
def step(fn):
    fv.grad=None; tex.grad=None
    img=fn(fv,tex,**c["settings"]); img.backward(g)
for nm,fn,iters in (("ours",srf.soft_rasterize,20),("ref",ref_gpu.soft_rasterize,3)):
    for _ in range(2): step(fn)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): step(fn)
    e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/iters
    print(f"{nm}: {ms:.3f} ms/step N={N} -> {N/ms*1e3:.0f} frames/s")
