"""Parity metrics shared by tests and golden scripts.

Floating-point tolerance of the north star: RGBA within 1e-4 relative, vertex gradients within
1e-3.  The soft rasterizer contains hard decisions (inside/outside of edge-on sliver faces, the
distance cut-off, z-test ties) that a 1-ulp difference can flip for an isolated pixel, so image
parity is stated as: every pixel within tolerance except a counted, bounded number of outliers."""
from __future__ import annotations

import numpy as np


def image_report(a, b, rel=1e-4):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    # NaN: a pixel that is NaN on one side only is an outlier (a plain `d > tol` would let it through); NaN on both sides
    # -- the reference's own arithmetic produces it, e.g. for a vertex at z = 0 -- counts as agreement
    na, nb = np.isnan(a), np.isnan(b)
    d = np.where(na | nb, 0.0, np.abs(np.where(na, 0.0, a) - np.where(nb, 0.0, b)))
    tol = rel * np.maximum(1.0, np.abs(np.where(nb, 0.0, b)))
    bad = (d > tol) | (na != nb)
    return dict(max_abs=float(d.max()) if d.size else 0.0, mean_abs=float(d.mean()) if d.size else 0.0,
                n=int(d.size), n_bad=int(bad.sum()), frac_bad=float(bad.mean()) if d.size else 0.0,
                p9999=float(np.quantile(d, 0.9999)) if d.size else 0.0)


def grad_report(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.abs(b).max()), 1e-30)
    nb = max(float(np.linalg.norm(b)), 1e-30)
    d = np.abs(a - b)
    return dict(max_abs=float(d.max()), scale=scale, max_over_scale=float(d.max() / scale),
                rel_l2=float(np.linalg.norm(a - b) / nb),
                frac_gt_1e3=float((d > 1e-3 * scale).mean()))


def fmt(r):
    return " ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items())
