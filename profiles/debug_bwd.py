import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, parity
from oracle import oracle as O
import lasr_b200.soft_renderer as sr
c = cases.GOLDEN_CASES["lasr_spot_s128"]()
dev = torch.device("cuda", 0)
fv = torch.from_numpy(c["face_vertices"]).to(dev).requires_grad_(True)
tex = torch.from_numpy(c["textures"]).to(dev).requires_grad_(True)
img = sr.functional.soft_rasterize(fv, tex, **c["settings"]); img.backward(torch.from_numpy(c["grad_soft_colors"]).to(dev))
r = O.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], np.float32)
gf, gt = O.soft_rasterize_backward(r, c["grad_soft_colors"])
ogt = tex.grad.cpu().numpy().reshape(gt.shape); ogf = fv.grad.cpu().numpy().reshape(gf.shape)
print("img", parity.fmt(parity.image_report(img.detach().cpu().numpy(), r["soft_colors"])))
print("gf", parity.fmt(parity.grad_report(ogf, gf))); print("gt", parity.fmt(parity.grad_report(ogt, gt)))
d = np.abs(ogt - gt).reshape(gt.shape[1], -1).max(1); bad = np.argsort(-d)[:8]
fvn = c["face_vertices"][0]
for f in bad:
    x = fvn[f][:,0]; y = fvn[f][:,1]
    print("face", f, "err", d[f], "ours", ogt[0,f].reshape(-1)[:3], "ref", gt[0,f].reshape(-1)[:3], "bbox px", (x.max()-x.min())*64, (y.max()-y.min())*64,
          "gf ours", ogf[0,f][:3], "ref", gf[0,f][:3])
print("n faces with err>1e-6:", (d>1e-6).sum(), "of", len(d))
