"""Mint golden vectors for the SoftRas / chamfer sibling kernels from the REFERENCE extensions.

Run on a B200 box (needs a GPU and oracle/_ref/*.so, built by `make -C oracle ref_aux` where /root/reference
exists):   python tests/golden/make_aux_golden.py [outdir]

Writes chamfer_ref.npz and texture_io_ref.npz: seeded inputs + the outputs of the unmodified reference kernels
(chamfer_3D.forward/backward, load_textures.load_textures, create_texture_image.create_texture_image), and prints
how this repo's kernels and the numpy oracles compare.  Copy the .npz files into tests/golden/ and commit them.
"""
import glob
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import aux_cases  # noqa: E402


def ref_ext(name):
    path = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + "*.so"))[0]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = "cuda"
    cham = ref_ext("chamfer_3D")
    out = {}
    for name, (a, b, g1, g2) in aux_cases.chamfer_cases().items():
        A, Bt = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        Bn, n, m = a.shape[0], a.shape[1], b.shape[1]
        d1 = torch.zeros(Bn, n, device=dev); d2 = torch.zeros(Bn, m, device=dev)
        i1 = torch.zeros(Bn, n, dtype=torch.int32, device=dev); i2 = torch.zeros(Bn, m, dtype=torch.int32, device=dev)
        cham.forward(A, Bt, d1, d2, i1, i2)
        ga = torch.zeros_like(A); gb = torch.zeros_like(Bt)
        cham.backward(A, Bt, ga, gb, torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev), i1, i2)
        torch.cuda.synchronize()
        for k, v in (("dist1", d1), ("dist2", d2), ("idx1", i1), ("idx2", i2), ("grad1", ga), ("grad2", gb)):
            out[f"{name}/{k}"] = v.cpu().numpy()
    np.savez_compressed(os.path.join(outdir, "chamfer_ref.npz"), **out)

    lt = ref_ext("load_textures"); cti = ref_ext("create_texture_image")
    out = {}
    for name, (image, faces, textures, is_update) in aux_cases.load_textures_cases().items():
        t = lt.load_textures(torch.from_numpy(image).to(dev), torch.from_numpy(faces).to(dev),
                             torch.from_numpy(textures).to(dev).clone(), torch.from_numpy(is_update).to(dev))
        torch.cuda.synchronize()
        out[f"load/{name}"] = t.cpu().numpy()
    for name, (faces, textures, image) in aux_cases.create_texture_image_cases().items():
        im = cti.create_texture_image(torch.from_numpy(faces).to(dev), torch.from_numpy(textures).to(dev),
                                      torch.from_numpy(image).to(dev).clone(), 1e-5)
        torch.cuda.synchronize()
        out[f"create/{name}"] = im.cpu().numpy()
    np.savez_compressed(os.path.join(outdir, "texture_io_ref.npz"), **out)
    print("wrote", outdir, sorted(os.listdir(outdir)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "aux_golden"))
