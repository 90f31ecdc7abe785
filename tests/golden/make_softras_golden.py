"""Mint golden vectors from the REFERENCE kernels and report the three-way gaps.

Run on a B200 box (needs a GPU):   python tests/golden/make_softras_golden.py [outdir]

For every case in tests/cases.GOLDEN_CASES it runs the unmodified reference SoftRas CUDA
extension (oracle/_ref, see oracle/ref_gpu.py) forward + backward (3x: the reference backward uses
float atomics, the spread is the noise floor), and stores inputs-hash + outputs as
<outdir>/softras_ref_<case>.npz.  It also runs the CPU oracle and this repo's CUDA path on the
same inputs and prints / stores the gaps (report.json), for GOLDEN and LIVE cases.
Copy the .npz files and report.json into tests/golden/ and commit them.
"""
import hashlib
import json
import os
import platform
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import parity  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import ref_gpu  # noqa: E402
from lasr_b200.soft_renderer import functional as srf  # noqa: E402


def run_torch(fn, c):
    dev = "cuda"
    fv = torch.from_numpy(c["face_vertices"]).to(dev).requires_grad_(True)
    tex = torch.from_numpy(c["textures"]).to(dev).requires_grad_(True)
    s = c["settings"]
    img = fn(fv, tex, **s)
    g = torch.from_numpy(c["grad_soft_colors"]).to(dev)
    img.backward(g)
    torch.cuda.synchronize()
    return img.detach().cpu().numpy(), fv.grad.cpu().numpy(), tex.grad.cpu().numpy()


def run_oracle(c, dtype=np.float32):
    r = O.soft_rasterize_forward(c["face_vertices"], c["textures"], c["settings"], dtype)
    gf, gt = O.soft_rasterize_backward(r, c["grad_soft_colors"])
    return r["soft_colors"], gf.reshape(c["face_vertices"].shape), gt.reshape(c["textures"].shape), r


def three_way(name, c, report):
    ref_img, ref_gf, ref_gt = run_torch(ref_gpu.soft_rasterize, c)
    spreads = []
    for _ in range(2):
        _, gf2, _ = run_torch(ref_gpu.soft_rasterize, c)
        spreads.append(np.abs(gf2 - ref_gf).max())
    our_img, our_gf, our_gt = run_torch(lambda a, b, **s: srf.soft_rasterize(a, b, **s), c)
    _, our_gf2, our_gt2 = run_torch(lambda a, b, **s: srf.soft_rasterize(a, b, **s), c)
    t0 = time.time()
    orc_img, orc_gf, orc_gt, orc = run_oracle(c)
    t_orc = time.time() - t0
    rep = dict(
        ours_vs_ref_img=parity.image_report(our_img, ref_img),
        ours_vs_ref_gf=parity.grad_report(our_gf, ref_gf), ours_vs_ref_gt=parity.grad_report(our_gt, ref_gt),
        oracle_vs_ref_img=parity.image_report(orc_img, ref_img),
        oracle_vs_ref_gf=parity.grad_report(orc_gf, ref_gf), oracle_vs_ref_gt=parity.grad_report(orc_gt, ref_gt),
        ours_vs_oracle_img=parity.image_report(our_img, orc_img),
        ours_vs_oracle_gf=parity.grad_report(our_gf, orc_gf), ours_vs_oracle_gt=parity.grad_report(our_gt, orc_gt),
        ref_grad_run_spread=float(max(spreads)), ref_grad_scale=float(np.abs(ref_gf).max()),
        ours_deterministic=bool(np.array_equal(our_gf, our_gf2) and np.array_equal(our_gt, our_gt2)),
        pairs=int(orc["pairs"]), oracle_seconds=t_orc)
    report[name] = rep
    print(f"== {name}  pairs={rep['pairs']}  oracle {t_orc:.2f}s  ref-grad-spread {rep['ref_grad_run_spread']:.3g}"
          f"  ours deterministic={rep['ours_deterministic']}")
    for k in ("ours_vs_ref_img", "ours_vs_ref_gf", "ours_vs_ref_gt", "oracle_vs_ref_img", "oracle_vs_ref_gf",
              "oracle_vs_ref_gt", "ours_vs_oracle_img", "ours_vs_oracle_gf"):
        print(f"   {k:22s} {parity.fmt(rep[k])}")
    return ref_img, ref_gf, ref_gt


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    report_only = "--report-only" in sys.argv      # re-measure the gaps without re-minting the vectors
    out = args[0] if args else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    report = {}
    for name, ctor in cases.GOLDEN_CASES.items():
        c = ctor()
        ref_img, ref_gf, ref_gt = three_way(name, c, report)
        h = hashlib.sha256(c["face_vertices"].tobytes() + c["textures"].tobytes()
                           + c["grad_soft_colors"].tobytes()).hexdigest()
        if report_only:
            continue
        np.savez_compressed(os.path.join(out, f"softras_ref_{name}.npz"), soft_colors=ref_img,
                            grad_faces=ref_gf, grad_textures=ref_gt, inputs_sha256=np.array(h),
                            settings=np.array(json.dumps(c["settings"])))
    for name, ctor in cases.LIVE_CASES.items():
        three_way(name, ctor(), report)
    meta = dict(gpu=torch.cuda.get_device_name(0), torch=torch.__version__, cuda=torch.version.cuda,
                driver=torch.cuda.driver_version() if hasattr(torch.cuda, "driver_version") else None,
                host=platform.processor(), cores=os.cpu_count(), when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()))
    with open(os.path.join(out, "report.json"), "w") as fh:
        json.dump(dict(meta=meta, cases=report), fh, indent=1)
    print("meta", meta)


if __name__ == "__main__":
    main()
