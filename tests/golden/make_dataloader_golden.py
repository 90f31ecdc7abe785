"""Golden vectors for the dataloader -> device staging path from the REFERENCE's own code.

Run in the build container (needs /root/reference and cv2; no GPU):
    python tests/golden/make_dataloader_golden.py

Writes a tiny synthetic video (PNG frames + masks, PFM flow / occlusion maps) to a temp directory, imports the
reference's ``dataloader/vidbase.py`` unmodified, builds a ``BaseDataset`` over those files and calls its
``__getitem__`` (vidbase.py:52-231).  Stored in tests/golden/dataloader_ref.npz: the decoded inputs exactly as the
reference read them from disk (so the test needs neither the files nor cv2) and every array of the returned
``elem`` that LASR's ``set_input`` consumes.  Modules the reference imports but that are absent here and not on
this path (absl flags, pypng, skimage's contour tracer) are stubbed; arithmetic is untouched.
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def write_pfm(path, data):
    """Little-endian PFM, rows bottom-to-top (what ext_utils/util_flow.py:36-80 reads back with flipud)."""
    data = np.asarray(data, np.float32)
    with open(path, "wb") as fh:
        fh.write(b"PF\n" if data.ndim == 3 else b"Pf\n")
        fh.write(f"{data.shape[1]} {data.shape[0]}\n".encode())
        fh.write(b"-1.0\n")
        np.flipud(data).astype("<f4").tofile(fh)


def smooth(rng, shape, sigma):
    import cv2
    return cv2.GaussianBlur(rng.normal(size=shape).astype(np.float32), (0, 0), sigma)


def synth_video(tmp, n_frames=3, H=96, W=136, seed=0, centre=(52, 40), step=(9, 4), axes=(24.0, 17.0), mask_scale=1,
                cam_files=False):
    import cv2
    rng = np.random.RandomState(seed)
    lists = dict(img=[], mask=[], fw=[], bw=[], cam=[])
    yy, xx = np.mgrid[0:H, 0:W]
    for t in range(n_frames):
        img = np.clip(127 + 60 * smooth(rng, (H, W, 3), 3.0) * 4, 0, 255).astype(np.uint8)
        cx, cy = centre[0] + step[0] * t, centre[1] + step[1] * t
        mask = ((((xx - cx) / (axes[0] + 2 * t)) ** 2 + ((yy - cy) / (axes[1] + t)) ** 2) < 1).astype(np.uint8) * 255
        p = dict(img=f"{tmp}/img-{t:05d}.png", mask=f"{tmp}/mask-{t:05d}.png", fw=f"{tmp}/flo-fw-{t:05d}.pfm",
                 bw=f"{tmp}/flo-bw-{t:05d}.pfm", cam=f"{tmp}/cam-{t:05d}.txt")
        if mask_scale != 1:          # masks stored at another resolution: the reference resizes + erodes them per item
            mask = cv2.resize(mask, (W // mask_scale, H // mask_scale), interpolation=cv2.INTER_NEAREST)
        cv2.imwrite(p["img"], img); cv2.imwrite(p["mask"], mask)
        if cam_files:                # per-frame camera file: focal, tx, ty, qw, qx, qy, qz, depth (vidbase.py:174-181)
            np.savetxt(p["cam"], np.concatenate([[1.0], rng.normal(size=2) * 0.1, rng.normal(size=4), [5 + t]]))
        for d in ("fw", "bw"):
            flow = np.zeros((H, W, 3), np.float32)
            flow[..., :2] = np.round((6 * smooth(rng, (H, W, 2), 5.0) * 5 + (3 if d == "fw" else -3)) * 32) / 32   # 1/32 px: compresses
            flow[..., 2] = (rng.uniform(size=(H, W)) > 0.05)
            occ = (np.round(np.abs(12 * smooth(rng, (H, W), 4.0) * 4) * 16) / 16).astype(np.float32)
            write_pfm(p[d], flow); write_pfm(p[d].replace("flo-", "occ-"), occ)
        for k in lists:
            lists[k].append(p[k])
    return lists


def main():
    for name in ("absl", "absl.flags", "absl.app", "png"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["absl"].flags = sys.modules["absl.flags"]; sys.modules["absl"].app = sys.modules["absl.app"]
    sys.path.insert(0, os.path.join(REF, "third_party")); sys.path.insert(0, REF)
    from ext_utils import image as image_utils
    image_utils.sample_contour = lambda m: np.zeros((1000, 2))      # skimage contour tracer: not on this path
    from ext_utils.util_flow import readPFM
    import cv2
    from dataloader import vidbase

    out = {}
    # video "" : object well inside the frame; video "b_": object at the frame corner, so the square crop reaches
    # outside the image and cv2.remap's constant border (complement colour / zeros) is exercised; video "m_": masks
    # stored at half resolution (the reference resizes them to the frame and erodes them, vidbase.py:67-76) and per-frame
    # camera files (:174-181)
    videos = {"": dict(seed=0), "b_": dict(seed=1, H=80, W=120, centre=(22, 14), step=(-3, 2), axes=(26.0, 15.0)),
              "m_": dict(seed=2, H=64, W=96, centre=(44, 30), step=(4, 2), axes=(20.0, 14.0), mask_scale=2, cam_files=True)}
    for pre, kw in videos.items():
        with tempfile.TemporaryDirectory() as tmp:
            L = synth_video(tmp, **kw)
            ds = vidbase.BaseDataset(types.SimpleNamespace(img_size=64))
            ds.imglist, ds.masklist, ds.flowfwlist, ds.flowbwlist, ds.camlist = L["img"], L["mask"], L["fw"], L["bw"], L["cam"]
            ds.dframe, ds.can_frame = 1, 0
            ds.baselist = [0, 1, 1, 2]                  # forward pairs then backward pairs, as dataloader/vid.py builds them
            ds.directlist = [1, 1, 0, 0]
            ds.num_imgs = len(ds.baselist)
            items = (0, 1, 3)
            out[pre + "items"] = np.asarray(items); out[pre + "img_size"] = np.asarray(64)
            out[pre + "baselist"] = np.asarray(ds.baselist); out[pre + "directlist"] = np.asarray(ds.directlist)
            # decoded inputs, exactly as the reference decodes them
            out[pre + "frames_bgr"] = np.stack([cv2.imread(p) for p in L["img"]])
            out[pre + "masks"] = np.stack([cv2.imread(p, 0) for p in L["mask"]])
            if os.path.exists(L["cam"][0]):
                out[pre + "cams"] = np.stack([np.loadtxt(p) for p in L["cam"]])
            for d in ("fw", "bw"):
                out[pre + f"flow_{d}"] = np.stack([np.ascontiguousarray(readPFM(p)[0]) for p in L[d]])
                out[pre + f"occ_{d}"] = np.stack([np.ascontiguousarray(readPFM(p.replace("flo-", "occ-"))[0]) for p in L[d]])
            for i in items:
                elem = ds[i]
                for k in ("img", "imgn", "mask", "flow", "flown", "occ", "occn", "cam", "camn", "pps", "shape", "id0", "id1", "depth", "depthn"):
                    out[pre + f"item{i}/{k}"] = np.asarray(elem[k])
    path = os.path.join(ROOT, "tests", "golden", "dataloader_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("img", "mask", "flow", "occ", "cam", "pps"):
        v = out[f"item0/{k}"]
        print(k, v.shape, v.dtype, float(np.min(v)), float(np.max(v)))


if __name__ == "__main__":
    main()
