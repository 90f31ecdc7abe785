"""oracle/texture_io.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatements of the reference's two texture-atlas kernels, float32 instantiation:

* ``load_textures``        third_party/softras/soft_renderer/cuda/load_textures_cuda_kernel.cu:8-66
* ``create_texture_image`` third_party/softras/soft_renderer/cuda/create_texture_image_cuda_kernel.cu:10-73

The float kernels mix precisions (double literals) and nvcc contracts their float sums of products; both are
reproduced: a float32 fma is emulated through float64 (exact product, one extra rounding), with the association
read off ``cuobjdump -sass oracle/_ref/{load_textures,create_texture_image}_kernel.o``
(a*b + c*d + e*f -> fma(e,f, fma(a,b, c*d));  a*b - c*d -> fma(a,b, -(c*d));  a*b + c*d + e -> fma(a,b, c*d) + e;
c += p*q -> fma(p, q, c)).

Parity status: pinned against the reference's own kernels run on a B200 (tests/golden/texture_io_ref.npz, produced
by tests/golden/make_aux_golden.py through oracle/_ref/*.so).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def _mul(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32)


def load_textures(image, faces, textures, is_update):
    """image [H,W,3], faces [F,3,2], textures [F,R*R,3] (copied), is_update [F] -> textures."""
    image = np.asarray(image, f32); faces = np.asarray(faces, f32)
    out = np.array(textures, f32, copy=True)
    F, T, _ = out.shape
    R = int(np.sqrt(T))                                     # load_textures_cuda_kernel.cu:78
    H, W = image.shape[:2]
    i = np.arange(R * R)
    w_y, w_x = (i // R), (i % R)
    lower = (w_x + w_y) < R
    w0 = np.where(lower, (w_x + 1. / 3.) / R, ((R - 1. - w_x) + 2. / 3.) / R).astype(f32)     # :28-35 (double -> float)
    w1 = np.where(lower, (w_y + 1. / 3.) / R, ((R - 1. - w_y) + 2. / 3.) / R).astype(f32)
    w2 = (1. - w0.astype(np.float64) - w1.astype(np.float64)).astype(f32)
    for fn in range(F):
        if is_update[fn] == 0:
            continue
        f = faces[fn].reshape(6)
        px = _mul(_fma(f[4], w2, _fma(f[0], w0, _mul(f[2], w1))), f32(W - 1))                  # :40-43
        py = _mul(_fma(f[5], w2, _fma(f[1], w0, _mul(f[3], w1))), f32(H - 1))
        x0 = px.astype(np.int32); y0 = py.astype(np.int32)
        y1 = (py + f32(1)).astype(f32).astype(np.int32); x1 = x0 + 1
        wx1 = (px - x0.astype(f32)).astype(f32); wx0 = (f32(1) - wx1).astype(f32)
        wy1 = (py - y0.astype(f32)).astype(f32); wy0 = (f32(1) - wy1).astype(f32)
        cx0, cx1 = np.clip(x0, 0, W - 1), np.clip(x1, 0, W - 1)   # the reference reads past the end with weight 0
        cy0, cy1 = np.clip(y0, 0, H - 1), np.clip(y1, 0, H - 1)
        for k in range(3):
            c = _fma(image[cy0, cx0, k], _mul(wx0, wy0), f32(0))                                # :48-54
            c = _fma(image[cy1, cx0, k], _mul(wx0, wy1), c)
            c = _fma(image[cy0, cx1, k], _mul(wx1, wy0), c)
            c = _fma(image[cy1, cx1, k], _mul(wx1, wy1), c)
            out[fn, :, k] = c
    return out


def create_texture_image(faces, textures, image, eps=1e-5):
    """faces [F,3,2] (atlas pixels), textures [F,R*R,3], image [H,W,3] (copied) -> image."""
    faces = np.asarray(faces, f32); textures = np.asarray(textures, f32)
    out = np.array(image, f32, copy=True)
    H, W = out.shape[:2]
    F = textures.shape[0]
    R = int(np.sqrt(textures.shape[1]))
    tile_width = int(np.sqrt(F - 1)) + 1                    # create_texture_image_cuda_kernel.cu:84
    res_out = W // tile_width                               # :85
    eps = f32(eps)
    i = np.arange(H * W)
    x = (i % (tile_width * res_out)).astype(np.int64); y = (i // (tile_width * res_out)).astype(np.int64)
    fn = x // res_out + (y // res_out) * tile_width         # :23-25
    ok = fn < F
    i, x, y, fn = i[ok], x[ok], y[ok], fn[ok]
    p = faces[fn]                                           # [P,3,2]
    p0, p1, p2 = p[:, 0], p[:, 1], p[:, 2]
    sub = lambda a, b: (a - b).astype(f32)
    cross = lambda a, b, c, d: _fma(a, b, -_mul(c, d))      # a*b - c*d
    inv = [sub(p1[:, 1], p2[:, 1]), sub(p2[:, 0], p1[:, 0]), cross(p1[:, 0], p2[:, 1], p2[:, 0], p1[:, 1]),
           sub(p2[:, 1], p0[:, 1]), sub(p0[:, 0], p2[:, 0]), cross(p2[:, 0], p0[:, 1], p0[:, 0], p2[:, 1]),
           sub(p0[:, 1], p1[:, 1]), sub(p1[:, 0], p0[:, 0]), cross(p0[:, 0], p1[:, 1], p1[:, 0], p0[:, 1])]
    # p2x*(p0y-p1y) + p0x*(p1y-p2y) + p1x*(p2y-p0y)  ->  fma(e,f, fma(a,b, c*d))
    den = _fma(p1[:, 0], sub(p2[:, 1], p0[:, 1]), _fma(p2[:, 0], sub(p0[:, 1], p1[:, 1]), _mul(p0[:, 0], sub(p1[:, 1], p2[:, 1]))))
    de = (den + eps).astype(f32)
    inv = [(v / de).astype(f32) for v in inv]
    xf, yf = x.astype(f32), y.astype(f32)
    w = []
    for k in range(3):
        v = (_fma(inv[3 * k], xf, _mul(inv[3 * k + 1], yf)) + inv[3 * k + 2]).astype(f32)      # :53
        w.append(np.maximum(np.minimum(v.astype(np.float64), 1.), 0.).astype(f32))             # :54 (double clamp)
    w_sum = ((w[0] + w[1]).astype(f32) + w[2]).astype(f32)
    ws = (w_sum + eps).astype(f32)
    w = [(v / ws).astype(f32) for v in w]
    w_x = _mul(w[0], f32(R)).astype(np.int32)
    w_y = _mul(w[1], f32(R)).astype(np.int32)
    # (w0 + w1) * R - w_x - w_y <= 1 : the product is contracted with the first subtraction
    frac = (_fma((w[0] + w[1]).astype(f32), f32(R), -w_x.astype(f32)) - w_y.astype(f32)).astype(f32)
    texel = np.where(frac <= 1, w_y * R + w_x, (R - 1 - w_y) * R + (R - 1 - w_x))
    texel = np.clip(texel, 0, R * R - 1)
    out.reshape(-1, 3)[i] = textures[fn, texel]
    return out
