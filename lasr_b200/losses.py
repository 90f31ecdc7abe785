"""Image losses of the LASR render hot path.

In the reference these are inline regions of ``LASR.forward`` (nnutils/mesh_net.py:372-447)
built from Python double loops over (frame, hypothesis) with boolean-mask indexing and a host
synchronisation per pair (``if mask[i,j].sum()==0``, :412).  Here each loss is one masked
reduction over the whole batch with no host round trip; values follow the reference's
definitions exactly (masked means per (frame, hypothesis), then a plain mean).
The 0.005*LPIPS term of the texture loss (:442-443) is a convnet and outside this path.
"""
from __future__ import annotations

import torch


def _masked_mean(x, m, dims):
    """sum(x*m)/sum(m) over dims; 0/0 -> nan, exactly like ``x[m].mean()`` on an empty selection."""
    m = m.to(x.dtype)
    return (x * m).sum(dims) / m.sum(dims)


def mask_loss(mask_pred, masks, occ):
    """Silhouette loss (mesh_net.py:374-390).
    mask_pred [N,H,S,S] rendered alpha; masks [N,S,S]; occ [N,S,S] (0 = pixel outside the frame).
    Returns (scalar, per-(frame,hypothesis) [N,H])."""
    valid = (occ != 0)[:, None]
    sub = 0.5 * _masked_mean((mask_pred - masks[:, None]).pow(2), valid.expand_as(mask_pred), (2, 3))
    return sub.mean(), sub


def flow_loss(flow_rd, bgmask, flow_obs, occ, masks):
    """Flow re-projection loss (mesh_net.py:393-416).
    flow_rd [N,H,S,S,2]; bgmask [N,H,S,S] bool; flow_obs [N,3,S,S]; occ, masks [N,S,S].
    Returns (scalar, per-(frame,hypothesis) [N,H], visibility mask [N,H,S,S])."""
    N, H = flow_rd.shape[:2]
    mask = (~bgmask) & ((occ != 0) & (masks > 0))[:, None]
    diff = flow_rd - flow_obs[:, None, :2].permute(0, 1, 3, 4, 2)
    rd_map = torch.norm(diff, 2, -1)
    w = (-occ).sigmoid()[:, None].expand(N, H, -1, -1)
    w = w / _masked_mean(w, mask, (1, 2, 3))[:, None, None, None]   # per frame, over all hypotheses (:406-407)
    rd_map = rd_map * w
    cnt = mask.sum((2, 3))
    sub = (rd_map * mask).sum((2, 3)) / cnt.clamp(min=1)
    sub = 0.5 * torch.where(cnt == 0, torch.zeros_like(sub), sub)     # empty mask -> 0 (:412)
    return sub.mean(), sub, mask


def texture_l1_loss(texture_render, fgmask, imgs, masks, occ, l1tex_wt=1.0):
    """Texture L1 terms (mesh_net.py:361-363, 420-441, 444-445) without the LPIPS term.
    texture_render [N,H,3,S,S]; fgmask [N,H,S,S]; imgs [N,3,S,S]; masks, occ [N,S,S]."""
    m = (masks > 0).float()[:, None]
    img_obs = imgs * m
    img_white = 1 - m + img_obs
    img_rnd = texture_render * fgmask[:, :, None]
    valid = (occ != 0)[:, None]
    a = (img_obs[:, None] - img_rnd).abs().mean(2)
    b = (img_white[:, None] - texture_render).abs().mean(2)
    sub = _masked_mean(a, valid.expand_as(a), (2, 3)) + _masked_mean(b, valid.expand_as(b), (2, 3))
    sub = 0.25 * (sub * 2 * l1tex_wt)
    return sub.mean(), sub
