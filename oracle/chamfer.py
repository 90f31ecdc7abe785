"""oracle/chamfer.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's brute-force nearest-neighbour kernels
(third_party/chamfer3D/chamfer3D.cu:12-129 forward, :157-176 backward) behind chamfer_3DFunction
(third_party/chamfer3D/dist_chamfer_3D.py:27-63).

* forward: for every point of cloud 1 the SQUARED distance to, and the int32 index of, its nearest point in
  cloud 2 (and vice versa).  Ties: the reference scans targets in ascending order with a strict ``<`` inside a
  512-point tile (:36-115) and a strict ``>`` across tiles (:117) -> the lowest index among equal distances,
  which is what ``argmin`` returns.  The distance is evaluated in float32 with nvcc's contraction of
  ``x2*x2+y2*y2+z2*z2`` (fma(z2,z2, fma(x2,x2, y2*y2)); `cuobjdump -sass oracle/_ref/chamfer_kernel.o`),
  emulated exactly through float64 (a float32 product is exact in float64; each fma is one rounding).
* backward: grad p1_j = 2 g1_j (p1_j - p2_idx1[j]) - sum_{k: idx2[k]=j} 2 g2_k (p2_k - p1_j), float64 accumulation
  (the reference scatters with float atomics in arbitrary order; compare with a tolerance).

Parity status: pinned against the reference's own kernels run on a B200 (tests/golden/chamfer_ref.npz, produced by
tests/golden/make_aux_golden.py through oracle/_ref/chamfer_3D*.so).
"""
from __future__ import annotations

import numpy as np


def _fma32(a, b, c):
    """float32 fma(a, b, c) for float32 arrays: a*b is exact in float64; the float64 sum is then rounded
    once more to float32 (double rounding can differ from a true fma only when the float64 sum lies within
    2^-53 relative of a float32 rounding boundary -- not observed on the golden vectors)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def sqdist(p, q):
    """[n,3] x [m,3] float32 -> [n,m] float32 squared distances with the reference's rounding."""
    d = q[None, :, :].astype(np.float32) - p[:, None, :].astype(np.float32)      # x2 = buf - x1  (:33-35)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    return _fma32(z, z, _fma32(x, x, (y * y).astype(np.float32)))


def chamfer_forward(xyz1, xyz2, chunk=2048):
    xyz1 = np.ascontiguousarray(xyz1, np.float32)
    xyz2 = np.ascontiguousarray(xyz2, np.float32)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist1 = np.empty((B, n), np.float32); idx1 = np.empty((B, n), np.int32)
    dist2 = np.empty((B, m), np.float32); idx2 = np.empty((B, m), np.int32)
    for b in range(B):
        for (p, q, dist, idx) in ((xyz1[b], xyz2[b], dist1[b], idx1[b]), (xyz2[b], xyz1[b], dist2[b], idx2[b])):
            for s in range(0, len(p), chunk):
                d = sqdist(p[s:s + chunk], q)
                i = d.argmin(1)                      # first minimum = lowest index among ties
                idx[s:s + chunk] = i
                dist[s:s + chunk] = d[np.arange(len(i)), i]
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    x1 = np.asarray(xyz1, np.float64); x2 = np.asarray(xyz2, np.float64)
    g1 = np.zeros_like(x1); g2 = np.zeros_like(x2)
    for b in range(x1.shape[0]):
        d = 2.0 * np.asarray(graddist1[b], np.float64)[:, None] * (x1[b] - x2[b][idx1[b]])
        g1[b] += d
        np.add.at(g2[b], idx1[b], -d)
        d = 2.0 * np.asarray(graddist2[b], np.float64)[:, None] * (x2[b] - x1[b][idx2[b]])
        g2[b] += d
        np.add.at(g1[b], idx2[b], -d)
    return g1.astype(np.float32), g2.astype(np.float32)
