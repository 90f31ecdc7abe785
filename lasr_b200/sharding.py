"""Frame sharding of the render hot path across the GPUs of one box.

Frames are independent units: each rank renders a contiguous block of frames with the mesh topology
and shared parameters replicated; there is NO collective on the data path.  Two small exchanges
exist around it:

* the batch-global depth window (nnutils/mesh_net.py:304-311) -- two scalars, MIN / MAX;
* the gradients of parameters shared by every frame (mean shape, vertex colours, bone controls) --
  one flat fp32 bucket, one all-reduce(SUM) per step (the reference gets this from DDP,
  nnutils/train_utils.py:100-109,277).

Works on any ``torch.distributed`` backend (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def frame_shard(n_frames: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of frames for `rank`; blocks differ by at most one frame."""
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_depth_window(local_zmin: torch.Tensor, local_zmax: torch.Tensor, group=None):
    """near/far over the frames of ALL ranks: all-reduce the two extrema, then the reference's
    formula.  Makes every frame's render independent of how frames were sharded."""
    zmin, zmax = local_zmin.detach().clone().reshape(1), local_zmax.detach().clone().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(zmin, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(zmax, op=dist.ReduceOp.MAX, group=group)
    half = (zmax - zmin) / 2
    return (zmin - half)[0], (zmax + half)[0]


def allreduce_shared_grads(grads, group=None, average=False):
    """Sum (or average) the gradients of the shared parameters over ranks with ONE collective.
    `grads` is a list of tensors; they are packed into a flat fp32 bucket, all-reduced and unpacked
    in place.  Returns the list."""
    grads = [g for g in grads if g is not None]
    if not grads:
        return grads
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads
