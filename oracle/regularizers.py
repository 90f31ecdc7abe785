"""CPU oracle of the mesh regularisers -- TEST INFRASTRUCTURE, not product code.

A numpy restatement of the reference's own (dense) algorithms, kept deliberately in the reference's
form so that it can be read against it line by line:

* ``laplacian_loss``  third_party/ext_nnutils/loss_utils.py:34-65  (dense normalised Laplacian, ``L @ x``)
* ``flatten_loss``    third_party/ext_nnutils/loss_utils.py:67-152 (edge list construction + dihedral cosine)
* ``arap_loss``       nnutils/loss_utils.py:29-64                  (dense [V,V] edge differences)

Pinned: tests/test_oracle_golden.py checks every function against tests/golden/regularizer_ref.npz, which
tests/golden/make_regularizer_golden.py minted by running the reference's Python itself.  Only tests/,
__graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import numpy as np


def dense_adjacency(num_vertices, faces):
    """0/1 matrix with an entry for every face edge, both directions (loss_utils.py:36-43 of both files)."""
    A = np.zeros((num_vertices, num_vertices), np.float64)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        A[faces[:, a], faces[:, b]] = 1
        A[faces[:, b], faces[:, a]] = 1
    return A


def laplacian_matrix(num_vertices, faces):
    """ext loss_utils.py:40-57: -1 off-diagonal, degree on the diagonal, rows divided by their diagonal."""
    L = -dense_adjacency(num_vertices, faces)
    L[np.arange(num_vertices), np.arange(num_vertices)] = -L.sum(1)
    for i in range(num_vertices):
        if L[i, i] != 0:
            L[i, :] /= L[i, i]
    return L


def laplacian_loss(x, faces):
    """x [N,V,3] -> [N]: sum over vertices and coordinates of (L x)^2 (ext loss_utils.py:59-65, average=False)."""
    L = laplacian_matrix(x.shape[1], faces)
    y = np.einsum("ij,njc->nic", L, x.astype(np.float64))
    return (y ** 2).sum((1, 2))


def flatten_edges(faces):
    """ext loss_utils.py:73-103: edges from face columns (0,1) and (1,2), each with the opposite vertices of
    its first and second incident face (ascending face id).  Returned sorted (the reference keeps Python-set
    order; the loss is a sum over edges, so the order only permutes fp additions)."""
    faces = np.asarray(faces)
    pairs = np.sort(np.concatenate((faces[:, 0:2], faces[:, 1:3]), axis=0), axis=1)
    edges = sorted(set(map(tuple, pairs.tolist())))
    out = []
    for v0, v1 in edges:
        opp = []
        for face in faces:                      # ascending face id
            if v0 in face and v1 in face:
                opp.append(int([u for u in face if u != v0 and u != v1][0]))
        out.append((v0, v1, opp[0], opp[1]))
    return np.asarray(out, np.int64)


def flatten_loss(x, faces, eps=1e-6):
    """x [N,V,3] -> [N]: sum over edges of (cos(dihedral) + 1)^2 (ext loss_utils.py:110-152, average=False)."""
    e = flatten_edges(faces)
    x = x.astype(np.float64)
    v0, v1, v2, v3 = (x[:, e[:, k]] for k in range(4))

    def side(b):
        a = v1 - v0
        al2, bl2 = (a ** 2).sum(-1), (b ** 2).sum(-1)
        al1, bl1 = np.sqrt(al2 + eps), np.sqrt(bl2 + eps)
        ab = (a * b).sum(-1)
        cos = ab / (al1 * bl1 + eps)
        sin = np.sqrt(1 - cos ** 2 + eps)
        cb = b - a * (ab / (al2 + eps))[..., None]
        return cb, bl1 * sin

    cb1, l1 = side(v2 - v0)
    cb2, l2 = side(v3 - v0)
    cos = (cb1 * cb2).sum(-1) / (l1 * l2 + eps)
    return ((cos + 1) ** 2).sum(1)


def arap_loss(dx, x, faces):
    """(dx, x) [N,V,3] -> [N]: mean over directed edges (a,b) of | |x_b-x_a|^2 - |dx_b-dx_a|^2 |
    (nnutils/loss_utils.py:46-64), through the same dense [V,V] intermediates."""
    A = dense_adjacency(x.shape[1], faces)
    mask = A.astype(bool)
    out = []
    for n in range(x.shape[0]):
        diffx = np.zeros_like(A)
        diffdx = np.zeros_like(A)
        for c in range(3):
            dsub = A * dx[n, None, :, c].astype(np.float64)          # A @ diag(d): entry [a,b] = A[a,b] d_b
            xsub = A * x[n, None, :, c].astype(np.float64)
            diffdx += (dsub - dx[n, :, c:c + 1]) ** 2
            diffx += (xsub - x[n, :, c:c + 1]) ** 2
        out.append(np.abs(diffx - diffdx)[mask].mean())
    return np.asarray(out)
