"""Golden vectors for the buffer set-up the REFERENCE's ``SoftRasterizeFunction`` performs around its pybind calls.

``oracle/ref_gpu.py`` restates that set-up (the reference's Python cannot travel to the GPU box) to drive the reference's
compiled kernels when golden vectors are minted and in ``bench.py --impl reference``.  This script pins the restatement:
it loads third_party/softras/soft_renderer/functional/soft_rasterize.py unmodified with a recording stand-in for the
pybind module ``soft_renderer.cuda.soft_rasterize`` and stores every argument of ``forward_soft_rasterize`` /
``backward_soft_rasterize`` (tensors as handed over -- zero / background pre-fills included -- and scalars) for a few
settings.  tests/test_oracle_golden.py replays the same inputs through ``ref_gpu.RefSoftRasterize`` with the same
stand-in and demands identical arguments.  Build container only; no GPU (``.to(device)`` of CPU tensors is a no-op).

    python tests/golden/make_function_glue_golden.py      ->  tests/golden/function_glue_ref.npz
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

SRC = "/root/reference/third_party/softras/soft_renderer/functional/soft_rasterize.py"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "function_glue_ref.npz")
FWD_T = ("face_vertices", "textures", "faces_info", "aggrs_info", "soft_colors")
BWD_T = ("face_vertices", "textures", "soft_colors", "faces_info", "aggrs_info", "grad_faces", "grad_textures", "grad_soft_colors")
SCALARS = ("image_size", "near", "far", "eps", "sigma_val", "func_dist", "dist_eps", "gamma_val", "func_rgb", "func_alpha",
           "texture_sample", "fill_back")
CASES = {
    "lasr": dict(image_size=6, background_color=[0, 0, 0], near=8.5, far=11.5, fill_back=True, eps=1e-3, sigma_val=1e-4,
                 dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2, aggr_func_rgb="softmax", aggr_func_alpha="prod",
                 texture_type="vertex"),
    "white_surface": dict(image_size=5, background_color=[1, 0.5, 0.25], near=1, far=100, fill_back=False, eps=1e-2,
                          sigma_val=1e-5, dist_func="barycentric", dist_eps=1e-3, gamma_val=1e-4, aggr_func_rgb="hard",
                          aggr_func_alpha="sum", texture_type="surface"),
    "hard": dict(image_size=4, background_color=[0.1, 0.2, 0.3], near=0.5, far=20, fill_back=True, eps=1e-3, sigma_val=1e-6,
                 dist_func="hard", dist_eps=1e-4, gamma_val=1e-3, aggr_func_rgb="softmax", aggr_func_alpha="hard",
                 texture_type="vertex"),
}
ORDER = ("image_size", "background_color", "near", "far", "fill_back", "eps", "sigma_val", "dist_func", "dist_eps",
         "gamma_val", "aggr_func_rgb", "aggr_func_alpha", "texture_type")


class Recorder:
    """Stand-in for the pybind module: records the call, fills the outputs with seeded values."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.calls = {}

    def forward_soft_rasterize(self, *a):
        self.calls["fwd"] = ([t.clone() for t in a[:5]], list(a[5:]))
        faces_info, aggrs_info, soft_colors = a[2], a[3], a[4]
        faces_info.copy_(torch.rand(faces_info.shape, generator=self.g))
        aggrs_info.copy_(torch.rand(aggrs_info.shape, generator=self.g))
        soft_colors.copy_(torch.rand(soft_colors.shape, generator=self.g))
        return [faces_info, aggrs_info, soft_colors]

    def backward_soft_rasterize(self, *a):
        self.calls["bwd"] = ([t.clone() for t in a[:8]], list(a[8:]))
        a[5].copy_(torch.rand(a[5].shape, generator=self.g))
        a[6].copy_(torch.rand(a[6].shape, generator=self.g))
        return [a[5], a[6]]


def case_inputs(name, s):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    nb, nf, T = 2, 7, (3 if s["texture_type"] == "vertex" else 4)
    fv = torch.randn(nb, nf, 3, 3, generator=g).requires_grad_(True)
    tex = torch.rand(nb, nf, T, 3, generator=g).requires_grad_(True)
    grad = torch.randn(nb, 4, s["image_size"], s["image_size"], generator=g)
    return fv, tex, grad


def main():
    out = {}
    for name, s in CASES.items():
        rec = Recorder(1)
        cuda_pkg = types.ModuleType("soft_renderer.cuda")
        cuda_pkg.soft_rasterize = rec
        sys.modules.update({"soft_renderer": types.ModuleType("soft_renderer"), "soft_renderer.cuda": cuda_pkg,
                            "soft_renderer.cuda.soft_rasterize": rec})
        spec = importlib.util.spec_from_file_location("ref_soft_rasterize_py", SRC)
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
        fv, tex, grad = case_inputs(name, s)
        img = m.soft_rasterize(fv, tex, *[s[k] for k in ORDER])
        img.backward(grad)
        out[f"{name}/in/face_vertices"], out[f"{name}/in/textures"], out[f"{name}/in/grad"] = \
            fv.detach().numpy(), tex.detach().numpy(), grad.numpy()
        for tag, names in (("fwd", FWD_T), ("bwd", BWD_T)):
            tensors, scalars = rec.calls[tag]
            for n, t in zip(names, tensors):
                out[f"{name}/{tag}/{n}"] = t.detach().numpy()
            out[f"{name}/{tag}/scalars"] = np.asarray(json.dumps([float(x) if not isinstance(x, bool) else x for x in scalars]))
        out[f"{name}/out/image"] = img.detach().numpy()
        out[f"{name}/out/grad_faces"], out[f"{name}/out/grad_textures"] = fv.grad.numpy(), tex.grad.numpy()
        out[f"{name}/settings"] = np.asarray(json.dumps(s))
        print(name, rec.calls["fwd"][1])
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
