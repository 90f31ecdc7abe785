"""oracle/dataloader.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement (no cv2) of the per-item preprocessing of the reference's video dataloader,
``BaseDataset.__getitem__`` (dataloader/vidbase.py:52-193), from decoded frames to the arrays LASR's
``set_input`` (nnutils/train_utils.py:140-172) moves to the GPU:

  complement-colour background (:78-82), mask bounding box -> square crop of half-size int(1.2 * max half-extent)
  (:96-106), integer-offset cv2.remap with constant border (:108-123; at integer coordinates bilinear remap is a
  copy), cv2.resize to img_size (:128-139; INTER_LINEAR for image / flow / occlusion, INTER_NEAREST for the mask),
  flow re-expressed between the two crops and normalised to [-1,1] units (:141-160), validity channel (:151,160),
  focal scale 1/alp (:175-183), principal-point offsets (:172-173).

cv2's resize arithmetic is restated from its documented algorithm: ``scale = 1 / (dst / src)`` (two roundings),
``fx = float((dx + 0.5) * scale - 0.5)``, ``sx = floor(fx)``, ``fx -= sx``, clamped at both ends with weight 0,
float32 weights, horizontal pass then vertical; nearest: ``sx = min(floor(dx * scale), src - 1)``.

Parity status: pinned against the reference's own code run in the build container on three synthetic videos -- the
object inside the frame; at the frame corner, so that the crop's constant border is exercised; masks stored at half
resolution with per-frame camera files (tests/golden/dataloader_ref.npz, tests/golden/make_dataloader_golden.py).
Bit for bit when img_size is a power of two (LASR's 256 / 512; the goldens' 64): the crop is 2L x 2L with integer L, so
the tap weights are dyadic and every summation order gives the same float.  For other sizes (checked on 14 random
videos at img_size 100 while writing this) masks, boxes, cameras and principal points stay exact and image / flow /
occlusion agree to float32 rounding (<= 4e-6 relative): OpenCV's SIMD resize orders and fuses the two-tap sums its own way.
"""
from __future__ import annotations

import numpy as np


def crop_box(mask):
    """(cx, cy, L) of vidbase.py:96-106 for one mask [H,W] (uint8)."""
    ys, xs = np.where(mask > 0)
    cx, cy = (int(xs.max()) + int(xs.min())) // 2, (int(ys.max()) + int(ys.min())) // 2
    lx, ly = (int(xs.max()) - int(xs.min())) // 2, (int(ys.max()) - int(ys.min())) // 2
    return cx, cy, int(1.2 * max(lx, ly))


def _linear_taps(dst, src):
    scale = 1.0 / (float(dst) / float(src))
    fx = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
    sx = np.floor(fx).astype(np.int64)
    fx = (fx - sx.astype(np.float32)).astype(np.float32)
    lo = sx < 0
    sx[lo] = 0; fx[lo] = 0
    hi = sx >= src - 1
    sx[hi] = src - 1; fx[hi] = 0
    return sx, np.minimum(sx + 1, src - 1), fx


def _nearest_taps(dst, src):
    scale = 1.0 / (float(dst) / float(src))
    return np.minimum(np.floor(np.arange(dst) * scale).astype(np.int64), src - 1)


def _crop(src, cx, cy, L, border):
    """The 2L x 2L window at integer offset (cx-L, cy-L), constant border (cv2.remap at integer coordinates)."""
    H, W = src.shape[:2]
    ys = np.arange(2 * L) + (cy - L); xs = np.arange(2 * L) + (cx - L)
    out = np.empty((2 * L, 2 * L) + src.shape[2:], src.dtype)
    out[...] = border
    vy = (ys >= 0) & (ys < H); vx = (xs >= 0) & (xs < W)
    out[np.ix_(vy, vx)] = src[np.ix_(ys[vy], xs[vx])]
    return out


def _resize_linear(a, S):
    sx, sx1, fx = _linear_taps(S, a.shape[1]); sy, sy1, fy = _linear_taps(S, a.shape[0])
    shp = (1, S) + (1,) * (a.ndim - 2)
    fxb = fx.reshape(shp).astype(a.dtype); fyb = fy.reshape((S, 1) + (1,) * (a.ndim - 2)).astype(a.dtype)
    rows = a[:, sx] * (1 - fxb) + a[:, sx1] * fxb                  # horizontal pass
    return rows[sy] * (1 - fyb) + rows[sy1] * fyb                   # vertical pass


def _resize_nearest(a, S):
    return a[_nearest_taps(S, a.shape[0])][:, _nearest_taps(S, a.shape[1])]


def preprocess_pair(bgr0, mask0, flow0, occ0, bgr1, mask1, flow1, occ1, S):
    """One item of BaseDataset.__getitem__: frame 0 and its paired frame n, already decoded
    (bgr uint8 [H,W,3], mask uint8 [H,W], flow float32 [H,W,3], occ float32 [H,W]).  Returns the elem arrays."""
    res = {}
    boxes, alps = [], []
    for tag, bgr, mask, flow, occ in (("", bgr0, mask0, flow0, occ0), ("n", bgr1, mask1, flow1, occ1)):
        img = bgr[:, :, ::-1] / 255.0                                                      # :56
        m = (mask > 0)
        color = 1 - img[m].mean(0)                                                         # :78
        img = img * m[..., None].astype(float) + color[None, None] * (1 - m[..., None].astype(float))
        cx, cy, L = crop_box(mask)
        boxes.append((cx, cy, L)); alps.append(2 * L / S)
        res["img" + tag] = np.transpose(_resize_linear(_crop(img, cx, cy, L, color), S), (2, 0, 1))
        res["mask" + tag] = (_resize_nearest(_crop(mask.astype(np.int64), cx, cy, L, 0), S) > 0).astype(float)
        res["flow" + tag] = _resize_linear(_crop(flow.astype(np.float32), cx, cy, L, 0), S)
        res["occ" + tag] = _resize_linear(_crop(occ.astype(np.float32), cx, cy, L, 0), S)
    (cx, cy, L), (cxn, cyn, Ln) = boxes
    alp, alpn = alps
    bx, by = np.meshgrid(range(S), range(S))
    for tag, (ax, ay, aL, a_own, a_other), (ox, oy, oL) in (("", (cx, cy, L, alp, alpn), (cxn, cyn, Ln)),
                                                           ("n", (cxn, cyn, Ln, alpn, alp), (cx, cy, L))):
        f = res["flow" + tag]
        f[:, :, 0] += (ax - aL) - (ox - oL) + bx * (a_own - a_other)                      # :141-142 / :153-154
        f[:, :, 1] += (ay - aL) - (oy - oL) + by * (a_own - a_other)
        f /= np.float32(a_other)                                                           # :143 / :155
        f[:, :, 0] = 2 * (f[:, :, 0] / S)
        f[:, :, 1] = 2 * (f[:, :, 1] / S)
        f[:, :, 2] = np.logical_and(f[:, :, 2] != 0, res["occ" + tag] < 10)                # :151 / :160
        res["flow" + tag] = np.transpose(f, (2, 0, 1))
    res["mask"] = np.stack([res.pop("mask"), res.pop("maskn")])
    res["cam"] = np.asarray([1. / alp, 0., 0., 1., 0., 0., 0.])
    res["camn"] = np.asarray([1. / alpn, 0., 0., 1., 0., 0., 0.])
    res["pps"] = np.stack([np.asarray([float(cx - L), float(cy - L)]), np.asarray([float(cxn - Ln), float(cyn - Ln)])])
    return res


def pair_inputs(g, item, pre=""):
    """Decoded inputs of dataset item `item` of golden video `pre` ("" or "b_") from the golden file's frame
    stacks (vidbase.py:53-94)."""
    i0 = int(g[pre + "baselist"][item]); fwd = int(g[pre + "directlist"][item]) == 1
    i1 = i0 + 1 if fwd else i0 - 1
    G = lambda k: g[pre + k]
    f0, o0 = (G("flow_fw")[i0], G("occ_fw")[i0]) if fwd else (G("flow_bw")[i0], G("occ_bw")[i0])
    f1, o1 = (G("flow_bw")[i1], G("occ_bw")[i1]) if fwd else (G("flow_fw")[i1], G("occ_fw")[i1])
    return (G("frames_bgr")[i0], G("masks")[i0], f0, o0, G("frames_bgr")[i1], G("masks")[i1], f1, o1)
