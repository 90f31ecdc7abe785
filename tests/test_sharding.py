"""N>1 path on CPU: world_size-2 gloo processes exercise frame sharding, the depth-window
exchange and the single-bucket gradient all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lasr_b200 import sharding


def test_frame_shard_partitions():
    for n in (1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            blocks = [sharding.frame_shard(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lasr_b200 import geom_utils as gu
        g = torch.Generator().manual_seed(0)
        N, V = 6, 40
        verts = torch.randn(V, 3, generator=g) * 0.3                       # shared mesh (replicated)
        R = gu.quaternion_to_rotation_matrix(torch.randn(N, 4, generator=g))
        T = torch.zeros(N, 1, 3)
        T[:, 0, 2] = 10 + torch.arange(N) * 0.1                              # per-frame depth differs
        lo, hi = sharding.frame_shard(N, rank, world)
        v = verts.clone().requires_grad_(True)
        cam = gu.obj_to_cam(v[None].repeat(hi - lo, 1, 1), R[lo:hi], T[lo:hi], 1, 1, None)
        near, far = sharding.global_depth_window(cam[:, :, 2].min(), cam[:, :, 2].max())
        loss = (cam ** 2).sum()
        loss.backward()
        sharding.allreduce_shared_grads([v.grad])
        # single-process reference over all frames
        v2 = verts.clone().requires_grad_(True)
        cam2 = gu.obj_to_cam(v2[None].repeat(N, 1, 1), R, T, 1, 1, None)
        (cam2 ** 2).sum().backward()
        n2, f2 = gu.depth_window(cam2[:, :, 2])
        ok = torch.allclose(v.grad, v2.grad, rtol=1e-5, atol=1e-5) and abs(float(near - n2)) < 1e-6 \
            and abs(float(far - f2)) < 1e-6
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
