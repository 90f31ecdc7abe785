"""Offline estimate (CPU, numpy) of what row spans of the dilated TRIANGLE would save the backward over sweeping the
dilated BOUNDING BOX: for every face of one Spot frame at 256^2, count the pixels of its window, the pixels within
distance m*r of the triangle (r = sqrt(threshold), m = span margin) and the 32-pixel sweeps either needs.
usage: python profiles/analysis_backward_sweeps.py [sigma]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases

sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
S = 256
c = cases.lasr_case("spot", 1, S, sigma=sigma)
fv = c["face_vertices"].reshape(-1, 3, 3)[:, :, :2].astype(np.float64)      # [F,3,2]
r = np.sqrt(np.log(1 / 1e-4 - 1) * sigma)
coords = (2 * np.arange(S) + 1 - S) / S


def seg_dist2(p, a, b):
    ab = b - a
    t = np.clip(((p - a) @ ab) / max(ab @ ab, 1e-30), 0, 1)
    d = p - (a + t[:, None] * ab)
    return (d * d).sum(1)


def inside(p, t):
    def cr(a, b, q): return (b[0] - a[0]) * (q[:, 1] - a[1]) - (b[1] - a[1]) * (q[:, 0] - a[0])
    s0, s1, s2 = cr(t[0], t[1], p), cr(t[1], t[2], p), cr(t[2], t[0], p)
    return ((s0 >= 0) & (s1 >= 0) & (s2 >= 0)) | ((s0 <= 0) & (s1 <= 0) & (s2 <= 0))


tot = dict(window=0, box=0, sweeps_window=0)
margins = (1.0, 1.05, 1.25)
span = {m: dict(px=0, sweeps=0, rows=0) for m in margins}
nf = 0
for t in fv:
    lo, hi = t.min(0) - r, t.max(0) + r
    xs = np.where((coords >= lo[0]) & (coords <= hi[0]))[0]; ys = np.where((coords >= lo[1]) & (coords <= hi[1]))[0]
    if len(xs) == 0 or len(ys) == 0:
        continue
    nf += 1
    w, h = len(xs) + 1, len(ys) + 1          # the kernel's conservative window is about one pixel wider
    tot["window"] += w * h; tot["sweeps_window"] += -(-w * h // 32); tot["box"] += len(xs) * len(ys)
    gx, gy = np.meshgrid(coords[xs], coords[ys]); p = np.stack([gx.ravel(), gy.ravel()], 1)
    d2 = np.minimum(np.minimum(seg_dist2(p, t[0], t[1]), seg_dist2(p, t[1], t[2])), seg_dist2(p, t[2], t[0]))
    d2[inside(p, t)] = 0
    for m in margins:
        ok = (d2 <= (m * r) ** 2).reshape(len(ys), len(xs))
        # a row span is the hull of the row's accepted pixels (the dilated triangle is convex)
        px = sum((np.where(row)[0].max() - np.where(row)[0].min() + 1) for row in ok if row.any())
        span[m]["px"] += px; span[m]["sweeps"] += -(-px // 32); span[m]["rows"] += int(ok.any(1).sum())
print(f"Spot, 256^2, sigma {sigma:g}: {nf} faces with a non-empty window; dilation r = {r * S / 2:.2f} px")
print(f"  window (what the backward sweeps today): {tot['window'] / nf:6.1f} px/face, {tot['sweeps_window'] / nf:.2f} sweeps/face"
      f" (exact border box {tot['box'] / nf:.1f} px)")
for m in margins:
    s = span[m]
    print(f"  row spans of the triangle dilated by {m:.2f} r: {s['px'] / nf:6.1f} px/face, {s['sweeps'] / nf:.2f} sweeps/face "
          f"({100 * (1 - s['sweeps'] / tot['sweeps_window']):.0f} % fewer sweeps), {s['rows'] / nf:.1f} rows/face")
