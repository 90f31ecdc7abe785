"""Host side of the staged dataloader: a Dataset with the reference's attributes and index logic
(dataloader/vidbase.py:40-94: ``imglist``, ``masklist``, ``flowfwlist``, ``flowbwlist``, ``baselist``,
``directlist``, ``dframe``, ``can_frame``) whose ``__getitem__`` stops after file decode and returns the raw
arrays; everything the reference does next on the worker (:78-183) runs on the GPU in ``staging.stage_pairs``.

Decode needs OpenCV like the reference (``cv2.imread``); the PFM reader is ``pfm.readPFM``.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch.utils.data import Dataset

from .pfm import readPFM


class BaseDataset(Dataset):
    """img, mask, flow data loader -- raw (decoded, unprocessed) pairs."""

    def __init__(self, opts, filter_key=None):
        self.opts = opts
        self.img_size = opts.img_size

    def __len__(self):
        return self.num_imgs

    def pair_indices(self, index):
        im0idx = self.baselist[index]
        forward = self.directlist[index] == 1
        return im0idx, (im0idx + self.dframe if forward else im0idx - self.dframe), forward

    def __getitem__(self, index):
        import cv2
        im0idx, im1idx, forward = self.pair_indices(index)
        frames = [cv2.imread(self.imglist[i]) for i in (im0idx, im1idx)]
        masks = [cv2.imread(self.masklist[i], 0) for i in (im0idx, im1idx)]
        for k in (0, 1):
            if frames[k].ndim == 2:                       # grey frames are replicated (vidbase.py:61-65)
                frames[k] = np.repeat(frames[k][:, :, None], 3, axis=2)
            if masks[k].shape[:2] != frames[k].shape[:2]:
                # masks stored at another resolution: nearest-neighbour resize to the frame, then erode -- two
                # iterations for frame 0, one for its partner, as the reference does (vidbase.py:67-76)
                from scipy.ndimage import binary_erosion
                m = cv2.resize(masks[k], frames[k].shape[:2][::-1], interpolation=cv2.INTER_NEAREST)
                masks[k] = binary_erosion(m, iterations=2 if k == 0 else 1).astype(np.uint8) * 255
        if forward:                                       # vidbase.py:85-90
            flowpath, flowpathn = self.flowfwlist[im0idx], self.flowbwlist[im0idx + self.dframe]
        else:
            flowpath, flowpathn = self.flowbwlist[im0idx], self.flowfwlist[im0idx - self.dframe]
        flows = [np.ascontiguousarray(readPFM(p)[0], np.float32) for p in (flowpath, flowpathn)]
        occs = [np.ascontiguousarray(readPFM(p.replace("flo-", "occ-"))[0], np.float32) for p in (flowpath, flowpathn)]
        # per-frame camera files [focal, tx, ty, qw, qx, qy, qz, depth] when the dataset has them (vidbase.py:167-183);
        # the focal entry is replaced by the crop's scale on the device
        cam_init = np.tile(np.asarray([1., 0., 0., 1., 0., 0., 0.], np.float32), (2, 1))
        depth = np.zeros(2, np.float32)
        camlist = getattr(self, "camlist", None)
        if camlist is not None and os.path.exists(camlist[im0idx]):
            for k, i in enumerate((im0idx, im1idx)):
                c = np.loadtxt(camlist[i]).astype(np.float32)
                cam_init[k] = c[:-1]
                depth[k] = c[-1]
        return {"cam_init": cam_init, "depth": depth,
                "frames_bgr": np.stack(frames), "masks": np.stack(masks), "flows": np.stack(flows),
                "occs": np.stack(occs), "inds": index, "indsn": index, "id0": im0idx, "id1": im1idx,
                "is_canonical": self.can_frame == im0idx, "is_canonicaln": self.can_frame == im1idx,
                "dataid": getattr(self, "dataid", 0), "shape": np.asarray(frames[0].shape[:2][::-1]).copy()}


def collate_raw(items):
    """Stack raw items into the [P,2,...] arrays ``stage_pairs`` / ``DeviceStager.submit`` take."""
    out = {k: torch.from_numpy(np.stack([it[k] for it in items])) for k in ("frames_bgr", "masks", "flows", "occs", "shape", "cam_init", "depth")}
    for k in ("inds", "indsn", "id0", "id1", "is_canonical", "is_canonicaln", "dataid"):
        out[k] = torch.as_tensor([it[k] for it in items])
    return out
