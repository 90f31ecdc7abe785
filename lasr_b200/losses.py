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
from torch.autograd import Function


def _as_u8(mask):
    """A boolean mask as the kernels read it (one byte per pixel, non-zero = set): a bool tensor is reinterpreted in
    place (same storage, no conversion kernel); anything else is converted."""
    m = mask.detach()
    if m.dtype == torch.bool:
        return m.contiguous().view(torch.uint8)
    return m.to(torch.uint8).contiguous()


class _LossSums(Function):
    """[N,H,7] masked sums of lasr_b200/csrc/losses.cu (see the layout there); differentiable w.r.t.
    mask_pred, flow_rd, texture_render and fgmask."""

    @staticmethod
    def forward(ctx, mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs):
        from . import _cabi
        N, H, S = mask_pred.shape[0], mask_pred.shape[1], mask_pred.shape[-1]
        c = lambda t: None if t is None else t.detach().float().contiguous()
        mp, fr, tx, fgc, im, mk, oc, fo = (c(t) for t in (mask_pred, flow_rd, tex, fg, imgs, masks, occ, flow_obs))
        bg = None if bgmask is None else _as_u8(bgmask)
        sums = torch.empty(N, H, 7, dtype=torch.float32, device=mask_pred.device)
        _cabi.losses_forward(mp, fr, bg, tx, fgc, im, mk, oc, fo, sums, N, H, S)
        ctx.save_for_backward(*[t for t in (mp, fr, bg, tx, fgc, im, mk, oc, fo) if t is not None])
        ctx.present = [t is not None for t in (mp, fr, bg, tx, fgc, im, mk, oc, fo)]
        ctx.dims = (N, H, S)
        return sums

    @staticmethod
    def backward(ctx, gsums):
        from . import _cabi
        it = iter(ctx.saved_tensors)
        mp, fr, bg, tx, fgc, im, mk, oc, fo = [next(it) if p else None for p in ctx.present]
        N, H, S = ctx.dims
        need = ctx.needs_input_grad
        g_mp = torch.empty_like(mp) if need[0] else None
        g_fr = torch.empty_like(fr) if (fr is not None and need[1]) else None
        g_tx = torch.empty_like(tx) if (tx is not None and (need[3] or need[4])) else None
        g_fg = torch.empty_like(fgc) if (tx is not None and need[4]) else None
        _cabi.losses_backward(mp, fr, bg, tx, fgc, im, mk, oc, fo, gsums.float().contiguous(), g_mp, g_fr, g_tx, g_fg,
                              N, H, S)
        return g_mp, g_fr, None, (g_tx if need[3] else None), g_fg, None, None, None, None


class _FusedImageLosses(Function):
    """All three losses in two launches forward (sums; reduce + finalize) and three backward
    (coefficients, per-pixel gradients): no [N,H]-sized PyTorch arithmetic, no autograd graph of tiny ops."""

    @staticmethod
    def forward(ctx, mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs, l1tex_wt):
        from . import _cabi
        N, H, S = mask_pred.shape[0], mask_pred.shape[1], mask_pred.shape[-1]
        c = lambda t: t.detach().float().contiguous()
        mp, fr, tx, fgc, im, mk, oc, fo = (c(t) for t in (mask_pred, flow_rd, tex, fg, imgs, masks, occ, flow_obs))
        bg = _as_u8(bgmask)
        dev = mask_pred.device
        sums = torch.empty(N, H, 7, dtype=torch.float32, device=dev)
        sub = torch.empty(3, N, H, dtype=torch.float32, device=dev)
        out3 = torch.empty(4, dtype=torch.float32, device=dev)     # three means and their sum
        _cabi.losses_forward(mp, fr, bg, tx, fgc, im, mk, oc, fo, sums, N, H, S, sub, out3, l1tex_wt)
        ctx.save_for_backward(mp, fr, bg, tx, fgc, im, mk, oc, fo, sums)
        ctx.dims = (N, H, S, float(l1tex_wt))
        ctx.set_materialize_grads(False)        # an unused output (the per-pair losses) costs no zero-fill in the backward
        return out3, sub

    @staticmethod
    def backward(ctx, g_out3, g_sub):
        from . import _cabi
        mp, fr, bg, tx, fgc, im, mk, oc, fo, sums = ctx.saved_tensors
        N, H, S, wt = ctx.dims
        dev = mp.device
        g3 = torch.zeros(4, device=dev) if g_out3 is None else g_out3.float().contiguous()
        gsub = None if g_sub is None else g_sub.float().contiguous()
        gsums = torch.empty_like(sums)
        _cabi.losses_coefficients(sums, g3, gsub, gsums, N, H, wt)
        need = ctx.needs_input_grad
        g_mp = torch.empty_like(mp) if need[0] else None
        g_fr = torch.empty_like(fr) if need[1] else None
        g_tx = torch.empty_like(tx) if (need[3] or need[4]) else None
        g_fg = torch.empty_like(fgc) if need[4] else None
        _cabi.losses_backward(mp, fr, bg, tx, fgc, im, mk, oc, fo, gsums, g_mp, g_fr, g_tx, g_fg, N, H, S)
        return g_mp, g_fr, None, (g_tx if need[3] else None), g_fg, None, None, None, None, None


class _FusedHead(Function):
    """Flow re-projection (both directions) + all three losses straight from the three-set render
    [N*H,10,S,S] (lasr_b200_head_forward/backward): the flow, background-mask, texture and silhouette
    intermediates of mesh_net.py:88-104 / 372-447 never exist as separate tensors, and the backward
    writes the image gradient in one pass."""

    @staticmethod
    def forward(ctx, image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, l1tex_wt, N, H, want_flow,
                grad_link=None, pair_shift=0):
        from . import _cabi
        S = image.shape[-1]
        dev = image.device
        c = lambda t, sh: t.detach().float().reshape(sh).contiguous()
        img = image.detach().float().contiguous()
        po, pq = c(pp_own, (N * H, 2)), c(pp_pair, (N * H, 2))
        fo, fq = c(f_own, (N * H,)), c(f_pair, (N * H,))
        im, mk, oc, fl = (t.detach().float().contiguous() for t in (imgs, masks, occ, flow_obs))
        sums = torch.empty(N, H, 7, dtype=torch.float32, device=dev)
        flow_rd = bg = None
        if want_flow:
            flow_rd = torch.empty(N * H, S, S, 2, dtype=torch.float32, device=dev)
            bg = torch.empty(N * H, S, S, dtype=torch.uint8, device=dev)
        sub = torch.empty(3, N, H, dtype=torch.float32, device=dev)
        out3 = torch.empty(4, dtype=torch.float32, device=dev)     # three means and their sum
        _cabi.head_forward(img, po, pq, fo, fq, im, mk, oc, fl, sums, flow_rd, bg, N, H, S, sub, out3, l1tex_wt,
                           pair_shift)
        ctx.pair_shift = int(pair_shift)
        ctx.save_for_backward(img, po, pq, fo, fq, im, mk, oc, fl, sums)
        ctx.dims = (N, H, S, float(l1tex_wt))
        ctx.shapes = (pp_pair.shape, f_pair.shape)
        ctx.grad_link = grad_link
        ctx.set_materialize_grads(False)
        if want_flow:
            bgb = bg.view(torch.bool)          # the kernel wrote 0 / 1
            ctx.mark_non_differentiable(flow_rd, bgb)
            return out3, sub, flow_rd, bgb
        return out3, sub, None, None

    @staticmethod
    def backward(ctx, g_out3, g_sub, _gflow, _gbg):
        from . import _cabi
        img, po, pq, fo, fq, im, mk, oc, fl, sums = ctx.saved_tensors
        N, H, S, wt = ctx.dims
        dev = img.device
        g3 = torch.zeros(4, device=dev) if g_out3 is None else g_out3.float().contiguous()
        gsub = None if g_sub is None else g_sub.float().contiguous()
        gsums = torch.empty_like(sums)
        _cabi.losses_coefficients(sums, g3, gsub, gsums, N, H, wt)
        gpp = torch.empty(N * H, 2, dtype=torch.float32, device=dev)
        gf = torch.empty(N * H, dtype=torch.float32, device=dev)
        link = ctx.grad_link
        if link is not None:
            if link.aggrs is None:
                raise RuntimeError("fused_head(grad_link=...): the link was not attached to the render that produced this "
                                   "image (pass the same PackedGradLink to render_indexed)")
            # packed per-pixel gradient for the rasterizer's backward; autograd carries a zero-stride placeholder
            link.packed = torch.empty(N * H, 3, S * S, 4, dtype=torch.float32, device=dev)
            _cabi.head_backward_packed(img, po, pq, fo, fq, im, mk, oc, fl, gsums, link.aggrs, link.packed, gpp, gf,
                                       N, H, S, ctx.pair_shift)
            gimg = link.make_placeholder(img)
        else:
            if ctx.pair_shift:
                raise RuntimeError("fused_head(pair_shift=...): the planar backward has no shifted form")
            gimg = torch.empty_like(img)
            _cabi.head_backward(img, po, pq, fo, fq, im, mk, oc, fl, gsums, gimg, gpp, gf, N, H, S)
        return (gimg, None, gpp.reshape(ctx.shapes[0]), None, gf.reshape(ctx.shapes[1]), None, None, None, None,
                None, None, None, None, None, None)


def fused_head(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs, n_frames, n_hypo, l1tex_wt=1.0,
               want_flow=True, grad_link=None, pair_shift=0):
    """image [N*H,10,S,S] from the three-set render -> (mask_loss, flow_loss, texture_loss, per-pair dict (+ "total":
    (mask + flow) + texture, formed in the kernel), flow_rd [N*H,S,S,2] | None, bgmask [N*H,S,S] bool | None).  The flow outputs are for reporting only
    (not differentiable: their gradient path is inside the losses).
    pair_shift: frame ij takes entry (ij + pair_shift) mod N*H of pp_pair / f_pair (pass the own tensors again and
    bs * n_hypo instead of rolled copies; needs grad_link when differentiated)."""
    if (pair_shift and grad_link is None and torch.is_grad_enabled()
            and any(t.requires_grad for t in (image, pp_pair, f_pair))):
        raise ValueError("fused_head(pair_shift=...) differentiates through the packed gradient only: pass grad_link")
    out3, sub, flow_rd, bg = _FusedHead.apply(image, pp_own, pp_pair, f_own, f_pair, imgs, masks, occ, flow_obs,
                                              l1tex_wt, n_frames, n_hypo, want_flow, grad_link, pair_shift)
    return out3[0], out3[1], out3[2], dict(mask=sub[0], flow=sub[1], texture=sub[2], total=out3[3]), flow_rd, bg


def _native(t):
    return t.is_cuda


def _sums(mask_pred, masks, occ, flow_rd=None, bgmask=None, flow_obs=None, tex=None, fg=None, imgs=None):
    return _LossSums.apply(mask_pred, flow_rd, bgmask, tex, fg, imgs, masks, occ, flow_obs)


def _masked_mean(x, m, dims):
    """sum(x*m)/sum(m) over dims; 0/0 -> nan, exactly like ``x[m].mean()`` on an empty selection."""
    m = m.to(x.dtype)
    return (x * m).sum(dims) / m.sum(dims)


def mask_loss(mask_pred, masks, occ):
    """Silhouette loss (mesh_net.py:374-390).
    mask_pred [N,H,S,S] rendered alpha; masks [N,S,S]; occ [N,S,S] (0 = pixel outside the frame).
    Returns (scalar, per-(frame,hypothesis) [N,H])."""
    if _native(mask_pred):
        q = _sums(mask_pred, masks, occ)
        sub = 0.5 * q[..., 0] / q[..., 6]
        return sub.mean(), sub
    valid = (occ != 0)[:, None]
    sub = 0.5 * _masked_mean((mask_pred - masks[:, None]).pow(2), valid.expand_as(mask_pred), (2, 3))
    return sub.mean(), sub


def flow_loss(flow_rd, bgmask, flow_obs, occ, masks):
    """Flow re-projection loss (mesh_net.py:393-416).
    flow_rd [N,H,S,S,2]; bgmask [N,H,S,S] bool; flow_obs [N,3,S,S]; occ, masks [N,S,S].
    Returns (scalar, per-(frame,hypothesis) [N,H], visibility mask [N,H,S,S])."""
    N, H = flow_rd.shape[:2]
    mask = (~bgmask) & ((occ != 0) & (masks > 0))[:, None]
    if _native(flow_rd):
        q = _sums(torch.zeros_like(bgmask, dtype=torch.float32), masks, occ, flow_rd, bgmask, flow_obs)
        A, B, C = q[..., 1], q[..., 2], q[..., 3]
        wmean = B.sum(1, keepdim=True) / C.sum(1, keepdim=True)      # per frame, over all hypotheses (:406-407)
        sub = 0.5 * torch.where(C == 0, torch.zeros_like(A), A / wmean / C.clamp(min=1))
        return sub.mean(), sub, mask
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
    if _native(texture_render):
        q = _sums(fgmask, masks, occ, tex=texture_render, fg=fgmask, imgs=imgs)
        sub = 0.25 * ((q[..., 4] + q[..., 5]) / q[..., 6] * 2 * l1tex_wt)
        return sub.mean(), sub
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


def image_losses(mask_pred, masks, occ, flow_rd, bgmask, flow_obs, texture_render, fgmask, imgs, l1tex_wt=1.0):
    """All three losses from ONE fused pass (CUDA tensors): returns (mask_loss, flow_loss, texture_loss,
    per-(frame,hypothesis) dict).  Values equal mask_loss / flow_loss / texture_l1_loss.  If the
    silhouette and the texture foreground mask are the same tensor (LASR: both are the rendered alpha,
    mesh_net.py:358-359) their gradients are summed by autograd as usual."""
    out3, sub = _FusedImageLosses.apply(mask_pred, flow_rd, bgmask, texture_render, fgmask, imgs, masks, occ,
                                        flow_obs, l1tex_wt)
    return out3[0], out3[1], out3[2], dict(mask=sub[0], flow=sub[1], texture=sub[2], total=out3[3])
