"""Golden vectors for the image losses L1-L3 from the REFERENCE's own code.

The reference computes them inline in ``LASR.forward`` (nnutils/mesh_net.py:361-363 and :372-447), in a module that
cannot be imported without its CUDA extensions.  This script (build container only: needs /root/reference, no GPU)
reads exactly those source lines from the reference file, dedents them and executes them -- unmodified -- on seeded CPU
tensors, with ``self`` / ``opts`` stand-ins carrying the attributes the lines read, ``Tensor.cuda`` made the identity
and the LPIPS network (``self.ptex_loss``, out of scope: DESIGN.md) replaced by a stub returning zeros.  Stored in
tests/golden/losses_ref.npz: the inputs and ``mask_loss_sub / flow_rd_loss_sub / texture_loss_sub`` + the three means.

    python tests/golden/make_losses_golden.py
"""
import os
import textwrap
import types

import numpy as np
import torch

SRC = "/root/reference/nnutils/mesh_net.py"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "losses_ref.npz")


def reference_lines(first, last):
    with open(SRC) as fh:
        lines = fh.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def main():
    prep = reference_lines(361, 363)       # img_obs / img_rnd / img_white
    body = reference_lines(372, 447)       # mask, flow and texture losses
    assert "img_obs = self.imgs" in prep and "img_white" in prep and "# 1) mask loss" in body and "self.total_loss += self.texture_loss" in body
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}
    for case, (bs, H, S, seed) in {"h1": (3, 1, 32, 0), "h2": (2, 2, 24, 1)}.items():
        g = torch.Generator().manual_seed(seed)
        n2 = 2 * bs
        R = lambda *sh: torch.rand(*sh, generator=g)
        masks = (R(n2, S, S) > 0.45).float()
        occ = torch.randn(n2, S, S, generator=g); occ[R(n2, S, S) < 0.2] = 0
        imgs = R(n2, 3, S, S)
        flow = torch.randn(n2, 3, S, S, generator=g) * 0.05
        render = R(n2 * H, 4, S, S)                         # texture_render [N,4,S,S]: rgb + silhouette
        flow_rd = torch.randn(n2 * H, S, S, 2, generator=g) * 0.05
        bgmask = R(n2 * H, S, S) < 0.4
        bgmask[1] = True                                    # one (frame, hypothesis) with an empty flow mask
        self = types.SimpleNamespace(imgs=imgs, masks=masks, occ=occ, flow=flow, flow_rd=flow_rd, bgmask=bgmask,
                                     texture_render=render[:, :3], fgmask_tex=render[:, -1], mask_pred=render[:, -1],
                                     iters=1,
                                     ptex_loss=types.SimpleNamespace(forward_pair=lambda a, b: torch.zeros(a.shape[0])))
        opts = types.SimpleNamespace(n_hypo=H, img_size=S, opt_tex="yes", l1tex_wt=0.7, n_bones=1)
        ns = dict(self=self, opts=opts, torch=torch, local_batch_size=bs, np=np)
        exec(prep, ns)
        exec(body, ns)
        for k, v in dict(masks=masks, occ=occ, imgs=imgs, flow=flow, render=render, flow_rd=flow_rd, bgmask=bgmask).items():
            out[f"{case}/in/{k}"] = v.numpy()
        out[f"{case}/meta"] = np.asarray([bs, H, S], np.int64); out[f"{case}/l1tex_wt"] = np.asarray(0.7)
        for k in ("mask_loss_sub", "flow_rd_loss_sub", "texture_loss_sub", "mask_loss", "flow_rd_loss", "texture_loss", "total_loss"):
            out[f"{case}/out/{k}"] = np.asarray(getattr(self, k).detach().numpy())
        print(case, {k: float(getattr(self, k)) for k in ("mask_loss", "flow_rd_loss", "texture_loss", "total_loss")})
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
