"""Golden vectors from the REFERENCE's own synthetic-data generator, ``scripts/render_syn.py``, run UNMODIFIED and whole
on top of this repository's drop-in ``soft_renderer`` package.

render_syn.py is where BASELINE.json's benchmark shape comes from (Spot, the camera orbit of lines 143-156) and it is a
caller of the render path with settings LASR's trainer never uses: ``Mesh.from_obj(load_texture=True, texture_res=5,
texture_type='surface')`` (the load_textures kernel), a 512^2 ``SoftRenderer(sigma_val=1e-12, aggr_func_rgb='hard')``
render of 'surface' textures, and ``render_flow`` (two 'vertex' renders of 3-D positions + re-projection).

The script is executed with ``runpy`` from the reference tree (cwd = /root/reference, read only; nothing is written
there) with

  * ``soft_renderer`` = ``lasr_b200.soft_renderer`` (``lasr_b200.install_as_soft_renderer()``); because there is no GPU
    in the build container the two native calls are routed to the CPU oracles (``oracle/oracle.py`` rasterizer,
    ``oracle/texture_io.py`` texture sampler, both pinned on the reference's kernels) and ``Tensor.cuda`` is the identity;
  * ``kornia`` (not installed) = a stand-in with ``rotation_matrix_to_quaternion`` / ``quaternion_to_rotation_matrix`` in
    kornia 0.5.3's XYZW convention (the script only uses their composition, R -> q -> conj(q) -> R^T);
  * ``trimesh`` (not installed), ``ext_utils.io.mkdir_p``, ``cv2.imwrite``, ``np.savetxt``, ``util_flow.write_pfm``:
    recorders that capture what the script would have written to ``database/DAVIS/...`` instead of writing it;
  * ``np.random.seed(0)`` before the run (the script draws ``rotx`` from the global generator).

Stored: the camera files, the posed vertices, the silhouettes and colour frames as the 8-bit values cv2.imwrite stores and the flow
files subsampled by 4 with float64 checksums of the full arrays.  Build container only; no GPU.

    python tests/golden/make_render_syn_golden.py      ->  tests/golden/render_syn_ref.npz
"""
import os
import runpy
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "render_syn_ref.npz")
NFRAMES = 3


def rotation_matrix_to_quaternion(R):
    """[3,3] -> [4] XYZW (Shepperd's method; kornia 0.5.3's output order)."""
    R = R.double()
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0:
        s = torch.sqrt(t + 1.0) * 2
        q = [(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = torch.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s]
    elif R[1, 1] > R[2, 2]:
        s = torch.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s]
    else:
        s = torch.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s]
    return torch.stack(q).float()


def main():
    import cv2
    import lasr_b200
    import lasr_b200.soft_renderer.rasterizer as rast
    from lasr_b200 import _cabi
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    from oracle import texture_io as OT
    from test_host_glue import _OracleRasterize

    def oracle_rasterize(face_vertices, textures, image_size=256, background_color=[0, 0, 0], near=1, far=100,
                         fill_back=True, eps=1e-3, sigma_val=1e-5, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4,
                         aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="surface"):
        return _OracleRasterize.apply(face_vertices, textures, dict(
            image_size=image_size, background_color=list(background_color), near=near, far=far, fill_back=fill_back,
            eps=eps, sigma_val=sigma_val, dist_func=dist_func, dist_eps=dist_eps, gamma_val=gamma_val,
            aggr_func_rgb=aggr_func_rgb, aggr_func_alpha=aggr_func_alpha, texture_type=texture_type))
    rast.srf.soft_rasterize = oracle_rasterize

    def native_call(name, *a):                      # the one native call of Mesh.from_obj(load_texture=True)
        assert name == "load_textures", name
        image, faces, is_update, textures = a[:4]
        textures.copy_(torch.from_numpy(OT.load_textures(image.numpy(), faces.numpy(), textures.numpy().copy(),
                                                         is_update.numpy())))
    _cabi.native_call = native_call
    torch.Tensor.cuda = lambda self, *a, **k: self

    written = {}

    class Trimesh:
        def __init__(self, vertices=None, faces=None):
            self.vertices, self.faces = np.array(vertices), np.array(faces)      # copies: on the GPU `.cpu()` detaches them from the
                                                                                 # tensor the script goes on to project in place

        def export(self, path):
            written[path] = self.vertices.copy()

    flowlib = types.ModuleType("ext_utils.flowlib"); flowlib.flow_to_image = lambda f: np.zeros(f.shape[:2] + (3,), np.uint8)
    util_flow = types.ModuleType("ext_utils.util_flow")
    util_flow.write_pfm = lambda path, a: written.__setitem__(path, np.asarray(a).copy())
    ext_io = types.ModuleType("ext_utils.io"); ext_io.mkdir_p = lambda p: None
    ext_utils = types.ModuleType("ext_utils"); ext_utils.flowlib, ext_utils.util_flow, ext_utils.io = flowlib, util_flow, ext_io
    kornia = types.ModuleType("kornia")
    kornia.rotation_matrix_to_quaternion = rotation_matrix_to_quaternion
    kornia.quaternion_to_rotation_matrix = quaternion_to_rotation_matrix
    tm = types.ModuleType("trimesh"); tm.Trimesh = Trimesh
    sys.modules.update({"ext_utils": ext_utils, "ext_utils.flowlib": flowlib, "ext_utils.util_flow": util_flow,
                        "ext_utils.io": ext_io, "kornia": kornia, "trimesh": tm})
    lasr_b200.install_as_soft_renderer()
    cv2.imwrite = lambda path, a: written.__setitem__(path, np.asarray(a).copy())
    np.savetxt = lambda path, a: written.__setitem__(path, np.asarray(a).copy())

    os.chdir(REF)
    sys.argv = ["render_syn.py", "--outdir", "syn", "--model", "spot", "--nframes", str(NFRAMES)]
    np.random.seed(0)
    ns = runpy.run_path(os.path.join(REF, "scripts/render_syn.py"), run_name="__main__")

    base = "database/DAVIS/%s/Full-Resolution/syn/"
    out = {"nframes": np.int64(NFRAMES), "img_size": np.int64(ns["img_size"]), "bgcolor": np.asarray(ns["bgcolor"])}
    for i in range(NFRAMES):
        out[f"cam{i}"] = written[base % "Camera" + "%05d.txt" % i]
        out[f"mesh{i}"] = written[base % "Meshes" + "%05d.obj" % i].astype(np.float32)
        out[f"verts_px{i}"] = ns["verts_list"][i][0].numpy()
        mask = written[base % "Annotations" + "%05d.png" % i]
        to_u8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)      # cv2.imwrite's saturating round of floats
        out[f"mask{i}"] = to_u8(mask)                      # 128 * alpha: a handful of edge pixels are in between
        out[f"mask_sum{i}"] = np.float64(mask.astype(np.float64).sum())
        out[f"img{i}"] = to_u8(written[base % "JPEGImages" + "%05d.jpg" % i])
    for i in range(1, NFRAMES):
        for tag, j in (("FlowFW", i - 1), ("FlowBW", i)):
            fl = written[base % tag + "flo-%05d.pfm" % j]
            out[f"{tag}{j}"] = fl[::4, ::4].astype(np.float32)
            out[f"{tag}{j}_sum"] = fl.astype(np.float64).sum((0, 1))
            occ = written[base % tag + "occ-%05d.pfm" % j]
            assert (occ == -1).all()
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes;", len(written), "files captured")


if __name__ == "__main__":
    main()
