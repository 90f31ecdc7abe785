"""Golden vectors for the dense geometry glue, minted from the REFERENCE's own Python on CPU.

Imports (read-only, this container only) nnutils/geom_utils.py and SoftRas' pure-PyTorch
functionals from /root/reference and records their outputs on seeded inputs:
obj_to_cam (LBS), pinhole_cam, orthographic_cam, look_at, orthogonal, face_vertices,
vertex_normals, ambient/directional lighting.

    python tests/golden/make_geom_golden.py        ->  tests/golden/geom_ref.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geom_ref.npz")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    sys.modules.setdefault("soft_renderer", types.ModuleType("soft_renderer"))  # geom_utils.py:24 imports it
    gu = load(os.path.join(REF, "nnutils/geom_utils.py"), "ref_geom_utils")
    fdir = os.path.join(REF, "third_party/softras/soft_renderer/functional")
    look_at = load(os.path.join(fdir, "look_at.py"), "ref_look_at").look_at
    orthogonal = load(os.path.join(fdir, "orthogonal.py"), "ref_orthogonal").orthogonal
    face_vertices = load(os.path.join(fdir, "face_vertices.py"), "ref_face_vertices").face_vertices
    vertex_normals = load(os.path.join(fdir, "vertex_normals.py"), "ref_vertex_normals").vertex_normals
    ambient = load(os.path.join(fdir, "ambient_lighting.py"), "ref_ambient").ambient_lighting
    directional = load(os.path.join(fdir, "directional_lighting.py"), "ref_directional").directional_lighting
    # torch.cross without dim= is deprecated but still the reference's call; results are identical

    g = torch.Generator().manual_seed(0)
    out = {}
    # ---- LBS: N frames, H hypotheses folded into N, B bones
    N, V, B = 4, 50, 5
    verts = torch.randn(N, V, 3, generator=g) * 0.3

    def rand_rot(n):
        q = torch.randn(n, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        x, y, z, w = q.unbind(1)
        return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                            2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                            2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).view(n, 3, 3)

    Rmat = rand_rot(N * B)
    Tmat = torch.randn(N * B, 1, 3, generator=g) * 0.1
    Tmat[::B, 0, 2] += 10.0
    skin = torch.softmax(torch.randn(N, B - 1, V, 1, generator=g) * 2, dim=1)
    out.update(lbs_verts=verts, lbs_R=Rmat, lbs_T=Tmat, lbs_skin=skin,
               lbs_out=gu.obj_to_cam(verts, Rmat, Tmat, B, 1, skin),
               lbs_out_notocam=gu.obj_to_cam(verts, Rmat, Tmat, B, 1, skin, tocam=False),
               rigid_out=gu.obj_to_cam(verts, Rmat[::B].contiguous(), Tmat[::B].contiguous(), 1, 1, None))
    # ---- pinhole / orthographic: 2 frames x 2 hypotheses
    v4 = torch.cat([out["lbs_out"], torch.ones(N, V, 1)], -1)
    pp = torch.randn(2, 2, generator=g) * 0.1
    fl = 8 + torch.rand(N, 1, generator=g) * 4
    out.update(cam_v4=v4, cam_pp=pp, cam_fl=fl, pinhole_out=gu.pinhole_cam(v4, pp, fl),
               ortho_out=gu.orthographic_cam(v4, pp, fl))
    # ---- SoftRas functionals
    eye = [0, 0, -(1.0 / np.tan(np.radians(30.0)) + 1)]
    # batch of 4, not 3: the reference calls torch.cross without dim= (look_at.py:50-51), which picks
    # the FIRST dimension of size 3 -- at batch size 3 that is the batch axis and the camera axes
    # come out zero.  That legacy-default artefact is not reproduced by lasr_b200.
    vv = torch.randn(4, 40, 3, generator=g)
    faces = torch.randint(0, 40, (4, 60, 3), generator=g, dtype=torch.int32)
    out.update(sr_verts=vv, sr_faces=faces, look_at_out=look_at(vv, eye),
               look_at_eye2=look_at(vv, [0.3, -0.2, -2.0]),
               orthogonal_out=orthogonal(vv, 1.0), orthogonal_s2=orthogonal(vv, 2.0),
               face_vertices_out=face_vertices(vv, faces), vertex_normals_out=vertex_normals(vv, faces))
    light = torch.zeros(4, 40, 3)
    light = ambient(light, 0.3, (1, 0.5, 0.25))
    light = directional(light, out["vertex_normals_out"], 0.7, (0.2, 1, 1), (0, 1, 0))
    out.update(light_out=light)
    np.savez_compressed(DST, **{k: v.detach().numpy() for k, v in out.items()})
    print("wrote", DST, {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
