"""chamfer_3DFunction / chamfer_3DDist with the reference's names and signatures
(third_party/chamfer3D/dist_chamfer_3D.py:27-75) on the B200 kernels of ``csrc/chamfer.cu``.

GPU tensors only, as in the reference ("GPU tensors only", :26): CPU tensors raise; there is no eager fallback.
Returns ``(dist1, dist2, idx1, idx2)``: SQUARED nearest-neighbour distances and int32 indices (lowest index among
equal distances).  LASR uses ``idx1`` only (nnutils/mesh_net.py:477); scripts/eval_mesh.py:168 uses ``dist1``.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function

from .. import _cabi


class chamfer_3DFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
            raise RuntimeError("chamfer_3D: expected xyz1 [B,n,3] and xyz2 [B,m,3]")
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        B, n, _ = xyz1.size()
        m = xyz2.size(1)
        dev = xyz1.device
        dist1 = torch.empty(B, n, device=dev)
        dist2 = torch.empty(B, m, device=dev)
        idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
        idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
        _cabi.native_call("chamfer_forward", xyz1, xyz2, dist1, dist2, idx1, idx2, B, n, m)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1 = graddist1.contiguous()
        graddist2 = graddist2.contiguous()
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
        B, n, _ = xyz1.size()
        _cabi.native_call("chamfer_backward", xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2,
                          B, n, xyz2.size(1))
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)
