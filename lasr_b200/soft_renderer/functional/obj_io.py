"""Wavefront OBJ import / export with texture atlases -- the host side of the two texture kernels.

Mirrors ``load_obj`` / ``load_textures`` / ``load_mtl`` (third_party/softras/soft_renderer/functional/load_obj.py:9-167)
and ``save_obj`` / ``create_texture_image`` (functional/save_obj.py:9-90): same arguments, same returned tensors
(CUDA), same file layout on export.  The resampling itself runs in ``csrc/texture_io.cu`` through the C ABI
(``lasr_b200_load_textures`` / ``lasr_b200_create_texture_image``); images are read and written with Pillow
(the reference uses scikit-image, which decodes to the same uint8 arrays).

Out of scope and still refused: ``save_voxel`` / ``voxelization``.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ... import _cabi


def _tokens(path):
    """Yield the whitespace-split fields of every non-empty line."""
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if t:
                yield t


def load_mtl(filename_mtl):
    """Diffuse colours (Kd) and texture file names (map_Kd) per material (load_obj.py:9-25)."""
    colors, texture_filenames = {}, {}
    current = ""
    for t in _tokens(filename_mtl):
        if t[0] == "newmtl":
            current = t[1]
        elif t[0] == "map_Kd":
            texture_filenames[current] = t[1]
        elif t[0] == "Kd":
            colors[current] = np.array([float(v) for v in t[1:4]])
    return colors, texture_filenames


def _read_image(path):
    from PIL import Image
    im = Image.open(path)
    if im.mode not in ("L", "RGB", "RGBA"):
        im = im.convert("RGB")
    return np.asarray(im)


def _uv_index(field):
    """`v/vt/vn` -> vt (1-based), 0 when the corner carries no texture coordinate (load_obj.py:50-62)."""
    return int(field.split("/")[1]) if ("/" in field and "//" not in field) else 0


def parse_obj_uv(filename_obj):
    """Host-side parse of the texture side of an OBJ (load_obj.py:29-67): per-corner texture coordinates
    [F,3,2] float32 (polygons fan-triangulated; corners without a `vt` index take the LAST `vt`, the reference's
    `vertices[0 - 1]`) and the material name of every triangle."""
    uv, corner_uv, face_material = [], [], []
    material = ""
    for t in _tokens(filename_obj):
        if t[0] == "vt":
            uv.append([float(v) for v in t[1:3]])
        elif t[0] == "f":
            idx = [_uv_index(v) for v in t[1:]]
            for i in range(len(idx) - 2):            # fan triangulation, as the geometry loader does
                corner_uv.append((idx[0], idx[i + 1], idx[i + 2]))
                face_material.append(material)
        elif t[0] == "usemtl":
            material = t[1]
    uv = np.asarray(uv, np.float32).reshape(-1, 2)
    faces = uv[np.asarray(corner_uv, np.int32).reshape(-1, 3) - 1]           # [F,3,2]
    return np.ascontiguousarray(faces), face_material


def parse_obj(filename_obj):
    """Host-side parse of the geometry of an OBJ (load_obj.py:113-137, 147-153): vertices [V,3] float32, per-vertex
    colours [V,3] (the optional `v x y z r g b` form; empty rows otherwise), faces [F,3] int32 0-based with polygons
    fan-triangulated, and the `mtllib` file name (None if absent)."""
    verts, colours, tris, mtllib = [], [], [], None
    for t in _tokens(filename_obj):
        if t[0] == "v":
            verts.append([float(v) for v in t[1:4]])
            colours.append([float(v) for v in t[4:7]])
        elif t[0] == "f":
            idx = [int(v.split("/")[0]) for v in t[1:]]
            tris.extend((idx[0], idx[i + 1], idx[i + 2]) for i in range(len(idx) - 2))
        elif t[0] == "mtllib":
            mtllib = t[1]
    return (np.asarray(verts, np.float32).reshape(-1, 3), colours, np.asarray(tris, np.int32).reshape(-1, 3) - 1, mtllib)


def load_textures(filename_obj, filename_mtl, texture_res):
    """'surface' textures [F, texture_res**2, 3] sampled from the materials' images (load_obj.py:28-101)."""
    faces, face_material = parse_obj_uv(filename_obj)
    faces = torch.from_numpy(faces).cuda()
    faces[1 < faces] = faces[1 < faces] % 1                                    # wrap (load_obj.py:70)

    colors, texture_filenames = load_mtl(filename_mtl)
    F = faces.shape[0]
    textures = torch.ones(F, texture_res ** 2, 3, dtype=torch.float32, device=faces.device)
    names = np.asarray(face_material)
    for name, color in colors.items():
        sel = torch.from_numpy(names == name).to(faces.device)
        textures[sel] = torch.from_numpy(color).to(textures)[None, None, :]
    for name, filename_texture in texture_filenames.items():
        image = _read_image(os.path.join(os.path.dirname(filename_obj), filename_texture)).astype(np.float32) / 255.
        if image.ndim == 2:                       # grey
            image = np.stack((image,) * 3, -1)
        if image.shape[2] == 4:                   # alpha ignored
            image = image[:, :, :3]
        image = torch.from_numpy(np.ascontiguousarray(image[::-1])).cuda()
        is_update = torch.from_numpy((names == name).astype(np.int32)).cuda()
        _cabi.native_call("load_textures", image, faces.contiguous(), is_update, textures, F, int(texture_res),
                          int(image.shape[0]), int(image.shape[1]))
    return textures


def load_obj(filename_obj, normalization=False, load_texture=False, texture_res=4, texture_type="surface"):
    """Vertices [V,3] float32, faces [F,3] int32 (0-based, polygons fan-triangulated) and optionally textures, all
    CUDA tensors (load_obj.py:104-167)."""
    assert texture_type in ["surface", "vertex"]
    verts, colours, tris, mtllib = parse_obj(filename_obj)
    vertices = torch.from_numpy(verts).cuda()
    faces = torch.from_numpy(tris).cuda()

    textures = None
    if load_texture and texture_type == "surface":
        if mtllib is None:
            raise Exception("Failed to load textures.")
        textures = load_textures(filename_obj, os.path.join(os.path.dirname(filename_obj), mtllib), texture_res)
    elif load_texture and texture_type == "vertex":
        textures = torch.from_numpy(np.asarray(colours, np.float32).reshape(-1, 3)).cuda()

    if normalization:                              # into a unit cube centred at zero (load_obj.py:156-160)
        vertices -= vertices.min(0)[0][None, :]
        vertices /= torch.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[0][None, :] / 2
    return (vertices, faces, textures) if load_texture else (vertices, faces)


def create_texture_image(textures, texture_res=16):
    """Atlas image [tile_h*res, tile_w*res, 3] (numpy, flipped for export) and the per-corner texture coordinates
    [F,3,2] in [0,1] (save_obj.py:9-36)."""
    num_faces = textures.shape[0]
    tile_width = int((num_faces - 1.) ** 0.5) + 1
    tile_height = int((num_faces - 1.) / tile_width) + 1
    dev = textures.device if textures.is_cuda else torch.device("cuda")
    image = torch.ones(tile_height * texture_res, tile_width * texture_res, 3, dtype=torch.float32, device=dev)
    n = torch.arange(num_faces)
    column = (n % tile_width).float()
    # Deliberate deviation: the reference writes `face_nums / tile_width` (save_obj.py:17), which was floor division
    # on the torch it was written for but is TRUE division on LASR's pinned torch 1.7.1 -- the triangles then
    # straddle the cells of other faces and the exported atlas does not reload.  The cell row is an integer.
    row = torch.div(n, tile_width, rounding_mode="floor").float()
    vertices = torch.zeros((num_faces, 3, 2), dtype=torch.float32)
    vertices[:, 0, 0] = column * texture_res + texture_res / 2
    vertices[:, 0, 1] = row * texture_res + 1
    vertices[:, 1, 0] = column * texture_res + 1
    vertices[:, 1, 1] = (row + 1) * texture_res - 1 - 1
    vertices[:, 2, 0] = (column + 1) * texture_res - 1 - 1
    vertices[:, 2, 1] = (row + 1) * texture_res - 1 - 1
    vertices = vertices.to(dev)
    textures = textures.to(dev).contiguous().float()
    R = int(round(textures.shape[1] ** 0.5))
    _cabi.native_call("create_texture_image", vertices, textures, image, num_faces, R, int(image.shape[0]),
                      int(image.shape[1]), 1e-5)
    vertices[:, :, 0] /= (image.shape[1] - 1)
    vertices[:, :, 1] /= (image.shape[0] - 1)
    return image.detach().cpu().numpy()[::-1, ::1], vertices.detach().cpu().numpy()


def save_obj(filename, vertices, faces, textures=None, texture_res=16, texture_type="surface"):
    """Write <name>.obj (+ .mtl and .png atlas for 'surface' textures): save_obj.py:39-90."""
    assert vertices.ndimension() == 2
    assert faces.ndimension() == 2
    assert texture_type in ["surface", "vertex"]
    assert texture_res >= 2
    surface = textures is not None and texture_type == "surface"
    stem = filename[:-4]
    if surface:
        from PIL import Image
        atlas, uv = create_texture_image(textures, texture_res)
        Image.fromarray((atlas.clip(0, 1) * 255).astype("uint8")).save(stem + ".png")
    faces = faces.detach().cpu().numpy()
    vertices = vertices.detach().cpu().numpy()
    out = ["# %s" % os.path.basename(filename), "#", ""]
    if surface:
        out += ["mtllib %s" % os.path.basename(stem + ".mtl"), ""]
    if textures is not None and texture_type == "vertex":
        colours = textures.detach().cpu().numpy()
        out += ["v %.8f %.8f %.8f %.8f %.8f %.8f" % (*v[:3], *c[:3]) for v, c in zip(vertices, colours)]
    else:
        out += ["v %.8f %.8f %.8f" % tuple(v[:3]) for v in vertices]
    out.append("")
    if surface:
        out += ["vt %.8f %.8f" % tuple(t) for t in uv.reshape(-1, 2)]
        out += ["", "usemtl material_1"]
        out += ["f %d/%d %d/%d %d/%d" % (f[0] + 1, 3 * i + 1, f[1] + 1, 3 * i + 2, f[2] + 1, 3 * i + 3)
                for i, f in enumerate(faces)]
        out.append("")
    else:
        out += ["f %d %d %d" % (f[0] + 1, f[1] + 1, f[2] + 1) for f in faces]
    with open(filename, "w") as fh:
        fh.write("\n".join(out) + "\n")
    if surface:
        with open(stem + ".mtl", "w") as fh:
            fh.write("newmtl material_1\nmap_Kd %s\n" % os.path.basename(stem + ".png"))
