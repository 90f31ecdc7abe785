"""Seeded inputs for the chamfer / texture-atlas kernels, shared by the golden generator (reference kernels on a GPU
box, tests/golden/make_aux_golden.py), the CPU oracle tests and the GPU parity tests."""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def chamfer_cases():
    """name -> (xyz1 [B,n,3], xyz2 [B,m,3], graddist1 [B,n], graddist2 [B,m]) float32."""
    rng = np.random.RandomState(11)
    out = {}

    def add(name, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        g1 = rng.normal(size=a.shape[:2]).astype(np.float32); g2 = rng.normal(size=b.shape[:2]).astype(np.float32)
        out[name] = (a, b, g1, g2)

    add("small", rng.normal(size=(2, 50, 3)), rng.normal(size=(2, 70, 3)))
    spot = np.load(os.path.join(_HERE, "golden", "spot_mesh.npz"))["vertices"].astype(np.float32)
    add("mirror", spot[None], (spot * np.array([-1, 1, 1], np.float32))[None])       # nnutils/mesh_net.py:474-477
    lat = rng.randint(0, 4, size=(1, 300, 3)).astype(np.float32)                      # integer lattice: exact ties
    add("ties", lat, rng.randint(0, 4, size=(1, 400, 3)).astype(np.float32))
    add("tiles", rng.uniform(-1, 1, size=(3, 1500, 3)), rng.uniform(-1, 1, size=(3, 1100, 3)))   # crosses 512 / 1024
    add("single", rng.normal(size=(2, 33, 3)), rng.normal(size=(2, 1, 3)))
    return out


def load_textures_cases():
    """name -> (image [H,W,3], faces [F,3,2] uv, textures [F,R*R,3], is_update [F] int32)."""
    rng = np.random.RandomState(12)
    out = {}
    for name, (H, W, F, R) in {"r4": (37, 53, 40, 4), "r1": (16, 16, 25, 1), "r6": (64, 48, 13, 6)}.items():
        image = rng.uniform(0, 1, size=(H, W, 3)).astype(np.float32)
        faces = rng.uniform(0, 0.999, size=(F, 3, 2)).astype(np.float32)
        textures = rng.uniform(0, 1, size=(F, R * R, 3)).astype(np.float32)
        is_update = (rng.uniform(size=F) < 0.7).astype(np.int32)
        out[name] = (image, faces, textures, is_update)
    return out


def atlas_vertices(num_faces, res):
    """The per-face triangle the exporter paints, in atlas pixels (functional/save_obj.py:14-23)."""
    tile_width = int((num_faces - 1.) ** 0.5) + 1
    n = np.arange(num_faces)
    column = (n % tile_width).astype(np.float32)
    row = (n // tile_width).astype(np.float32)
    v = np.zeros((num_faces, 3, 2), np.float32)
    v[:, 0, 0] = column * res + res / 2
    v[:, 0, 1] = row * res + 1
    v[:, 1, 0] = column * res + 1
    v[:, 1, 1] = (row + 1) * res - 1 - 1
    v[:, 2, 0] = (column + 1) * res - 1 - 1
    v[:, 2, 1] = (row + 1) * res - 1 - 1
    return v


def create_texture_image_cases():
    """name -> (faces [F,3,2] atlas pixels, textures [F,R*R,3], image [H,W,3] pre-filled with ones)."""
    rng = np.random.RandomState(13)
    out = {}
    for name, (F, R, res) in {"f10": (10, 4, 8), "f100": (100, 2, 16), "f7_random": (7, 3, 12)}.items():
        tile_width = int((F - 1.) ** 0.5) + 1
        tile_height = int((F - 1.) / tile_width) + 1
        faces = atlas_vertices(F, res)
        if "random" in name:   # arbitrary (also degenerate-ish) triangles, not the exporter's layout
            faces = (faces + rng.normal(0, 2.0, size=faces.shape)).astype(np.float32)
        textures = rng.uniform(0, 1, size=(F, R * R, 3)).astype(np.float32)
        image = np.ones((tile_height * res, tile_width * res, 3), np.float32)
        out[name] = (faces, textures, image)
    return out
