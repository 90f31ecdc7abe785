"""Golden vectors for the module layer of the render chain (R2) from the REFERENCE's own classes.

Loads, by path and unmodified, third_party/softras/soft_renderer/{mesh,lighting,transform,rasterizer,renderer}.py and
the pure-PyTorch functionals they call (look_at, perspective, orthogonal, face_vertices, vertex_normals,
ambient_lighting, directional_lighting, get_points_from_angles) into a stand-in ``soft_renderer`` package whose
``functional.soft_rasterize`` is a recorder: it stores everything the reference's ``SoftRenderer`` hands to the
rasterizer -- face vertices after lighting and camera, face textures, every scalar -- and returns a seeded "render"
so the anti-aliasing pool after it is pinned too.  (``import soft_renderer`` itself cannot run here: its
functional/__init__ imports the CUDA extensions.)  ``Tensor.cuda`` is the identity.

Cases cover what the classes offer beyond LASR's one configuration: perspective and orthographic look_at cameras,
eyes given as lists, set through ``set_eyes`` / ``set_eyes_from_angles`` (floats and per-frame tensors), 'surface'
and 'vertex' lighting with a directional light, both texture types, default textures, anti-aliasing, fill_back on and
off, ``set_sigma`` / ``set_gamma``, ``SoftRenderer.forward``.  (camera_mode='look' cannot be pinned: the reference's
``Transform`` passes Look's arguments in the wrong order and ``srf.look`` dereferences ``up=None``; 'projection' is out
of scope.)  Batch sizes and face counts avoid 3: the reference calls ``torch.cross`` without ``dim``.

    python tests/golden/make_modules_golden.py      ->  tests/golden/modules_ref.npz
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SR = "/root/reference/third_party/softras/soft_renderer"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "modules_ref.npz")
ARG_NAMES = ("image_size", "background_color", "near", "far", "fill_back", "eps", "sigma_val", "dist_func", "dist_eps",
             "gamma_val", "aggr_func_rgb", "aggr_func_alpha", "texture_type")


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def cases():
    """name -> (SoftRenderer kwargs, texture_type, textures ('rand' | None), texture_res, post-construction calls,
    call through forward()?)"""
    lasr = dict(image_size=8, sigma_val=1e-4, gamma_val=1e-2, camera_mode="look_at", perspective=False,
                aggr_func_rgb="softmax", light_mode="vertex", light_intensity_ambient=1.0, light_intensity_directionals=0.0)
    return {
        "lasr": (lasr, "vertex", "rand", 1, [], False),
        "persp_surface_aa": (dict(image_size=8, camera_mode="look_at", anti_aliasing=True, background_color=[0.2, 0.3, 0.4],
                                  near=0.5, far=50, fill_back=False, eps=1e-2, dist_func="barycentric", dist_eps=1e-3,
                                  aggr_func_alpha="sum"), "surface", "rand", 2, [], False),
        "ortho_vertex_dirlight": (dict(image_size=6, camera_mode="look_at", perspective=False, viewing_scale=1.7,
                                       eye=[0.3, -0.2, -3.0], light_mode="vertex", light_intensity_ambient=0.4,
                                       light_color_ambient=[1.0, 0.9, 0.8], light_intensity_directionals=0.7,
                                       light_color_directionals=[1.0, 0.5, 0.2], light_directions=[0.3, 0.8, -0.5],
                                       aggr_func_rgb="hard", dist_func="hard", aggr_func_alpha="hard"), "vertex", "rand", 1, [], False),
        "eyes_from_angles_float": (dict(image_size=8, camera_mode="look_at", viewing_angle=45), "surface", "rand", 1,
                                   [("set_eyes_from_angles", (2.7, 20.0, 40.0))], False),
        "eyes_from_angles_tensor": (dict(image_size=8, camera_mode="look_at", perspective=False), "vertex", "rand", 1,
                                    [("set_eyes_from_angles", "tensor")], False),
        "set_eyes_sigma_gamma": (dict(image_size=8, camera_mode="look_at"), "surface", None, 2,
                                 [("set_eyes", ([0.1, 0.4, -2.5],)), ("set_sigma", (3e-5,)), ("set_gamma", (2e-3,))], False),
        "forward_default_textures": (dict(image_size=8, camera_mode="look_at", light_mode="vertex"), "vertex", None, 1, [], True),
        "forward_surface": (dict(image_size=8, camera_mode="look_at", perspective=False), "surface", "rand", 1, [], True),
    }


def inputs(seed, texture_type, textures, texture_res):
    from lasr_b200 import synth
    v, f = synth.icosphere(1, 0.6)
    g = torch.Generator().manual_seed(seed)
    N = 2
    verts = torch.from_numpy(v)[None].repeat(N, 1, 1) + torch.randn(N, len(v), 3, generator=g) * 0.03
    faces = torch.from_numpy(f)[None].repeat(N, 1, 1)
    tex = None
    if textures == "rand":
        tex = torch.rand(N, len(f), texture_res ** 2, 3, generator=g) if texture_type == "surface" \
            else torch.rand(N, len(v), 3, generator=g)
    return verts, faces, tex, g


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    fdir = os.path.join(SR, "functional")
    srf = types.ModuleType("soft_renderer.functional")
    for fn in ("look_at", "look", "perspective", "orthogonal", "face_vertices", "vertex_normals", "ambient_lighting",
               "directional_lighting", "get_points_from_angles"):
        setattr(srf, fn, getattr(load(os.path.join(fdir, fn + ".py"), "ref_fn_" + fn), fn))
    sr = types.ModuleType("soft_renderer")
    sr.functional = srf
    sys.modules["soft_renderer"] = sr
    sys.modules["soft_renderer.functional"] = srf
    for mod, names in (("mesh", ("Mesh",)), ("lighting", ("AmbientLighting", "DirectionalLighting", "Lighting")),
                       ("transform", ("Projection", "LookAt", "Look", "Transform")), ("rasterizer", ("SoftRasterizer",)),
                       ("renderer", ("SoftRenderer",))):
        m = load(os.path.join(SR, mod + ".py"), "soft_renderer." + mod)
        for n in names:
            setattr(sr, n, getattr(m, n))

    out = {}
    for idx, (name, (kw, ttype, textures, tres, calls, via_forward)) in enumerate(cases().items()):
        verts, faces, tex, g = inputs(idx, ttype, textures, tres)
        rec = {}

        def soft_rasterize(face_vertices, textures_, *args):
            rec["face_vertices"], rec["face_textures"] = face_vertices.clone(), textures_.clone()
            rec["args"] = dict(zip(ARG_NAMES, args))
            S = rec["args"]["image_size"]
            rec["render"] = torch.rand(face_vertices.shape[0], 4, S, S, generator=g)
            return rec["render"].clone()
        srf.soft_rasterize = soft_rasterize
        r = sr.SoftRenderer(**kw)
        call_log = []
        for fn, a in calls:
            if a == "tensor":
                a = (torch.tensor([2.5, 3.0]), torch.tensor([10.0, -25.0]), torch.tensor([30.0, 200.0]))
                call_log.append([fn, [t.tolist() for t in a]])
            else:
                call_log.append([fn, list(a)])
            getattr(r, fn, None)(*a) if hasattr(r, fn) else getattr(r.transform, fn)(*a)
        if via_forward:
            img = r(verts.clone(), faces.clone(), None if tex is None else tex.clone(), texture_type=ttype)
        else:
            mesh = sr.Mesh(verts.clone(), faces.clone(), textures=None if tex is None else tex.clone(),
                           texture_res=tres, texture_type=ttype)
            img = r.render_mesh(mesh)
        out[f"{name}/in/verts"], out[f"{name}/in/faces"] = verts.numpy(), faces.numpy()
        if tex is not None:
            out[f"{name}/in/textures"] = tex.numpy()
        out[f"{name}/rec/face_vertices"] = rec["face_vertices"].numpy()
        out[f"{name}/rec/face_textures"] = rec["face_textures"].numpy()
        out[f"{name}/rec/render"] = rec["render"].numpy()
        out[f"{name}/out/image"] = img.numpy()
        out[f"{name}/meta"] = np.asarray(json.dumps(dict(kwargs=kw, texture_type=ttype, texture_res=tres, calls=call_log,
                                                         via_forward=via_forward, args=rec["args"])))
        print(name, tuple(rec["face_vertices"].shape), tuple(rec["face_textures"].shape), tuple(img.shape), rec["args"])
    # Mesh bookkeeping outside the render call: fill_back_ / reset_ and the cached derived tensors (mesh.py:67-146)
    for ttype in ("surface", "vertex"):
        verts, faces, tex, g = inputs(100, ttype, "rand", 2)
        mesh = sr.Mesh(verts.clone(), faces.clone(), textures=tex.clone(), texture_res=2, texture_type=ttype)
        n0 = mesh.surface_normals.clone()
        mesh.fill_back_()
        mesh.fill_back_()                                    # idempotent
        pre = f"mesh_ops_{ttype}"
        out[f"{pre}/in/verts"], out[f"{pre}/in/faces"], out[f"{pre}/in/textures"] = verts.numpy(), faces.numpy(), tex.numpy()
        out[f"{pre}/filled/faces"], out[f"{pre}/filled/textures"] = mesh.faces.numpy(), mesh.textures.numpy()
        out[f"{pre}/filled/face_vertices"] = mesh.face_vertices.numpy()
        out[f"{pre}/filled/surface_normals"] = mesh.surface_normals.numpy()
        out[f"{pre}/filled/vertex_normals"] = mesh.vertex_normals.numpy()
        if ttype == "surface":
            out[f"{pre}/filled/face_textures"] = mesh.face_textures.numpy()
        mesh.vertices = mesh.vertices * 2.0                  # the setter invalidates the caches
        out[f"{pre}/scaled/face_vertices"] = mesh.face_vertices.numpy()
        mesh.reset_()
        out[f"{pre}/reset/faces"], out[f"{pre}/reset/textures"] = mesh.faces.numpy(), mesh.textures.numpy()
        out[f"{pre}/reset/face_vertices"] = mesh.face_vertices.numpy()
        out[f"{pre}/meta"] = np.asarray(json.dumps(dict(num_faces=[int(n0.shape[1]), int(out[f"{pre}/filled/faces"].shape[1])],
                                                         texture_res=int(mesh.texture_res), batch_size=int(mesh.batch_size))))
        print(pre, out[f"{pre}/filled/faces"].shape, out[f"{pre}/filled/textures"].shape)
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
