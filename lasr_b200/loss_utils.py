"""Mesh regularisers with the reference's module surface:

* ``LaplacianLoss(vertex, faces, average=False)``  third_party/ext_nnutils/loss_utils.py:34-65
* ``FlattenLoss(faces, average=False)``            third_party/ext_nnutils/loss_utils.py:67-152
* ``ARAPLoss(vertex, faces, average=False)``       nnutils/loss_utils.py:29-64

The reference builds dense V x V matrices (the Laplacian; for ARAP six [N,V,V] tensors per call -- 2.2 GB
each at 64 frames of a 2930-vertex mesh).  Here the topology becomes a CSR adjacency / edge list once, at
construction, and CUDA tensors are served by O(N*E) kernels through the C ABI (csrc/meshreg.cu).  CPU
tensors take the same sparse formulation in plain PyTorch (what the host-side tests pin against the golden
vectors minted from the reference's Python); on a GPU box a missing library raises -- no silent switch.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function


# ----------------------------------------------------------------------------------------------
# topology tables (host, once per mesh)
# ----------------------------------------------------------------------------------------------
def vertex_adjacency(num_vertices: int, faces) -> tuple[np.ndarray, np.ndarray]:
    """Unique neighbours of every vertex, ascending, as CSR (offsets [V+1], neighbours [E2]) -- the non-zeros
    of the adjacency matrix the reference fills (loss_utils.py:36-43: assignment, so duplicates collapse)."""
    f = np.asarray(faces, dtype=np.int64)
    a = np.concatenate([f[:, [0, 1]], f[:, [1, 0]], f[:, [1, 2]], f[:, [2, 1]], f[:, [2, 0]], f[:, [0, 2]]], 0)
    key = np.unique(a[:, 0] * num_vertices + a[:, 1])
    src, dst = key // num_vertices, key % num_vertices
    offsets = np.zeros(num_vertices + 1, np.int64)
    np.add.at(offsets, src + 1, 1)
    return np.cumsum(offsets).astype(np.int32), dst.astype(np.int32)


def dihedral_edges(faces) -> np.ndarray:
    """[E,4] (v0, v1, v2, v3): the edges taken from face columns (0,1) and (1,2), each with the vertex
    opposite to it in its first and in its second incident face (ascending face id), which is what the
    reference's constructor collects (ext loss_utils.py:73-103).  Rows sorted by (v0, v1).  Needs every such
    edge to have exactly two incident faces (closed manifold), as the reference's indexing does."""
    f = np.asarray(faces, dtype=np.int64)
    F = len(f)
    pairs = np.sort(np.concatenate([f[:, 0:2], f[:, 1:3]], 0), axis=1)
    nv = int(f.max()) + 1
    wanted = np.unique(pairs[:, 0] * nv + pairs[:, 1])
    # every (sorted edge, face, opposite vertex) incidence of the mesh
    e_all = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    opp = np.concatenate([f[:, 2], f[:, 0], f[:, 1]], 0)
    face_id = np.tile(np.arange(F), 3)
    e_sorted = np.sort(e_all, axis=1)
    k_all = e_sorted[:, 0] * nv + e_sorted[:, 1]
    order = np.lexsort((face_id, k_all))
    k_all, opp = k_all[order], opp[order]
    start = np.searchsorted(k_all, wanted, side="left")
    stop = np.searchsorted(k_all, wanted, side="right")
    if not np.all(stop - start == 2):
        raise ValueError("FlattenLoss needs a closed manifold mesh: every edge must have exactly two faces")
    return np.stack([wanted // nv, wanted % nv, opp[start], opp[start + 1]], 1).astype(np.int32)


def corner_incidence(num_vertices: int, edges: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """vertex -> (edge, slot) incidences as CSR; entry = edge*4 + slot, ascending per vertex."""
    flat = edges.reshape(-1).astype(np.int64)
    order = np.argsort(flat, kind="stable")
    offsets = np.zeros(num_vertices + 1, np.int64)
    np.add.at(offsets, flat + 1, 1)
    return np.cumsum(offsets).astype(np.int32), order.astype(np.int32)


def _native(t):
    return t.is_cuda


def _f32c(t):
    return t.detach().float().contiguous()


# ----------------------------------------------------------------------------------------------
# autograd Functions over the C ABI
# ----------------------------------------------------------------------------------------------
class _Laplacian(Function):
    @staticmethod
    def forward(ctx, x, off, nbr):
        from . import _cabi
        N, V = x.shape[:2]
        xc = _f32c(x)
        lap = torch.empty_like(xc)
        loss = torch.empty(N, dtype=torch.float32, device=x.device)
        ws = _cabi.meshreg_workspace(N, V, 0, x.device)
        _cabi.meshreg_call("laplacian_forward", xc, off, nbr, lap, loss, ws, ws.numel(), N, V)
        ctx.save_for_backward(lap, off, nbr)
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import _cabi
        lap, off, nbr = ctx.saved_tensors
        N, V = lap.shape[:2]
        gx = torch.empty_like(lap)
        _cabi.meshreg_call("laplacian_backward", lap, _f32c(g), off, nbr, gx, N, V)
        return gx, None, None


class _Arap(Function):
    @staticmethod
    def forward(ctx, dx, x, off, nbr, n_edges):
        from . import _cabi
        N, V = x.shape[:2]
        dc, xc = _f32c(dx), _f32c(x)
        loss = torch.empty(N, dtype=torch.float32, device=x.device)
        ws = _cabi.meshreg_workspace(N, V, 0, x.device)
        _cabi.meshreg_call("arap_forward", dc, xc, off, nbr, loss, ws, ws.numel(), N, V, n_edges)
        ctx.save_for_backward(dc, xc, off, nbr)
        ctx.n_edges = n_edges
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import _cabi
        dc, xc, off, nbr = ctx.saved_tensors
        N, V = xc.shape[:2]
        gd, gx = torch.empty_like(dc), torch.empty_like(xc)
        _cabi.meshreg_call("arap_backward", dc, xc, _f32c(g), off, nbr, gd, gx, N, V, ctx.n_edges)
        return gd, gx, None, None, None


class _Flatten(Function):
    @staticmethod
    def forward(ctx, x, edges, inc_off, inc, eps):
        from . import _cabi
        N, V = x.shape[:2]
        E = edges.shape[0]
        xc = _f32c(x)
        loss = torch.empty(N, dtype=torch.float32, device=x.device)
        ws = _cabi.meshreg_workspace(N, V, E, x.device)
        _cabi.meshreg_call("flatten_forward", xc, edges, loss, ws, ws.numel(), N, V, E, float(eps))
        ctx.save_for_backward(xc, edges, inc_off, inc)
        ctx.eps = float(eps)
        return loss

    @staticmethod
    def backward(ctx, g):
        from . import _cabi
        xc, edges, inc_off, inc = ctx.saved_tensors
        N, V = xc.shape[:2]
        E = edges.shape[0]
        gx = torch.empty_like(xc)
        ws = _cabi.meshreg_workspace(N, V, E, xc.device)
        _cabi.meshreg_call("flatten_backward", xc, edges, _f32c(g), inc_off, inc, gx, ws, ws.numel(), N, V, E, ctx.eps)
        return gx, None, None, None, None


# ----------------------------------------------------------------------------------------------
# modules (the reference's constructors and forward signatures)
# ----------------------------------------------------------------------------------------------
def _np_faces(faces):
    return faces.detach().cpu().numpy() if torch.is_tensor(faces) else np.asarray(faces)


class LaplacianLoss(nn.Module):
    """sum over vertices of |x_v - mean of its neighbours|^2 per batch item ([N]); ``average`` -> scalar / N."""

    def __init__(self, vertex, faces, average=False):
        super().__init__()
        self.nv = int(vertex.shape[0])
        self.nf = int(faces.shape[0])
        self.average = average
        off, nbr = vertex_adjacency(self.nv, _np_faces(faces))
        self.register_buffer("nbr_offsets", torch.from_numpy(off))
        self.register_buffer("nbr", torch.from_numpy(nbr))

    def forward(self, x):
        batch_size = x.size(0)
        if _native(x):
            loss = _Laplacian.apply(x, self.nbr_offsets, self.nbr)
        else:
            deg = (self.nbr_offsets[1:] - self.nbr_offsets[:-1]).long()
            src = torch.repeat_interleave(torch.arange(self.nv), deg)
            s = torch.zeros_like(x).index_add_(1, src, x[:, self.nbr.long()])
            has = (deg > 0).to(x.dtype)[None, :, None]
            lap = (x - s / deg.clamp(min=1).to(x.dtype)[None, :, None]) * has
            loss = lap.pow(2).sum((1, 2))
        return loss.sum() / batch_size if self.average else loss


class ARAPLoss(nn.Module):
    """mean over edges of | |x_b-x_a|^2 - |dx_b-dx_a|^2 | per batch item ([N]); forward(dx, x) as the reference."""

    def __init__(self, vertex, faces, average=False):
        super().__init__()
        self.nv = int(vertex.shape[0])
        self.nf = int(faces.shape[0])
        self.average = average
        off, nbr = vertex_adjacency(self.nv, _np_faces(faces))
        self.n_edges = int(len(nbr))
        self.register_buffer("nbr_offsets", torch.from_numpy(off))
        self.register_buffer("nbr", torch.from_numpy(nbr))

    def forward(self, dx, x):
        if _native(x):
            return _Arap.apply(dx, x, self.nbr_offsets, self.nbr, self.n_edges)
        deg = (self.nbr_offsets[1:] - self.nbr_offsets[:-1]).long()
        a = torch.repeat_interleave(torch.arange(self.nv), deg)
        b = self.nbr.long()
        q = (x[:, b] - x[:, a]).pow(2).sum(-1) - (dx[:, b] - dx[:, a]).pow(2).sum(-1)
        return q.abs().mean(1)


class FlattenLoss(nn.Module):
    """sum over edges of (cos(dihedral angle) + 1)^2 per batch item ([N]); ``average`` -> scalar / N."""

    def __init__(self, faces, average=False):
        super().__init__()
        self.nf = int(faces.shape[0])
        self.average = average
        f = _np_faces(faces)
        edges = dihedral_edges(f)
        nv = int(f.max()) + 1
        inc_off, inc = corner_incidence(nv, edges)
        self.register_buffer("edges", torch.from_numpy(edges))
        self.register_buffer("inc_offsets", torch.from_numpy(inc_off))
        self.register_buffer("inc", torch.from_numpy(inc))
        for k, name in enumerate(("v0s", "v1s", "v2s", "v3s")):      # the reference's buffer names
            self.register_buffer(name, torch.from_numpy(edges[:, k].astype(np.int64)))

    def forward(self, vertices, eps=1e-6):
        batch_size = vertices.size(0)
        if _native(vertices):
            if vertices.shape[1] < self.inc_offsets.numel() - 1:
                raise ValueError("FlattenLoss: fewer vertices than the topology it was built for")
            inc_off = self.inc_offsets
            if vertices.shape[1] > inc_off.numel() - 1:   # trailing vertices no face uses: no incidences
                pad = inc_off[-1:].expand(vertices.shape[1] - (inc_off.numel() - 1))
                inc_off = torch.cat([inc_off, pad]).contiguous()
            loss = _Flatten.apply(vertices, self.edges, inc_off, self.inc, eps)
        else:
            p0, p1, p2, p3 = (vertices[:, i] for i in (self.v0s, self.v1s, self.v2s, self.v3s))

            def side(b):
                a = p1 - p0
                al2, bl2 = a.pow(2).sum(-1), b.pow(2).sum(-1)
                al1, bl1 = (al2 + eps).sqrt(), (bl2 + eps).sqrt()
                ab = (a * b).sum(-1)
                cos = ab / (al1 * bl1 + eps)
                sin = (1 - cos.pow(2) + eps).sqrt()
                return b - a * (ab / (al2 + eps))[:, :, None], bl1 * sin

            cb1, l1 = side(p2 - p0)
            cb2, l2 = side(p3 - p0)
            cos = (cb1 * cb2).sum(-1) / (l1 * l2 + eps)
            loss = (cos + 1).pow(2).sum(1)
        return loss.sum() / batch_size if self.average else loss
