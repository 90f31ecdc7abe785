"""Batch staging on the device: decoded frame pairs in, the tensors of the reference's collated batch out.

``stage_pairs`` is the functional form (CUDA tensors in, dict of CUDA tensors out, keys and shapes of the
reference's ``elem`` after ``default_collate``: dataloader/vidbase.py:193-229).  ``DeviceStager`` adds the host
side of it: pinned staging buffers, the H2D copies on a side stream, and an event the consumer stream waits on,
so batch i+1 is uploaded and preprocessed while step i renders.

No CPU fallback: the kernels live in ``csrc/staging.cu`` behind ``lasr_b200_stage_pairs``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi


def stage_pairs(frames_bgr, masks, flows, occs, img_size, cam_init=None):
    """frames_bgr [P,2,H,W,3] uint8, masks [P,2,H,W] uint8, flows [P,2,H,W,3] float32, occs [P,2,H,W] float32
    (CUDA, contiguous; slot 0 = frame 0 with the flow 0->n, slot 1 = frame n with the flow n->0).
    Returns the collated batch: img / imgn [P,3,S,S], mask [P,2,S,S], flow / flown [P,3,S,S], occ / occn [P,S,S],
    cam / camn [P,7], pps [P,2,2] (float32, CUDA) and 'boxes' [P,2,4] int32 (cx, cy, L, mask pixels).
    cam_init [P,2,7] (optional): the rows of the per-frame camera files (``collate_raw``'s 'cam_init'); the staged
    cameras keep their entries 1..6 and get the focal scale 1/alp in entry 0, as the reference does."""
    if frames_bgr.dim() != 5 or frames_bgr.shape[1] != 2 or frames_bgr.shape[4] != 3:
        raise RuntimeError("stage_pairs: frames_bgr must be [P,2,H,W,3]")
    P, _, H, W, _ = frames_bgr.shape
    if tuple(masks.shape) != (P, 2, H, W) or tuple(flows.shape) != (P, 2, H, W, 3) or tuple(occs.shape) != (P, 2, H, W):
        raise RuntimeError("stage_pairs: masks / flows / occs do not match frames_bgr")
    if frames_bgr.dtype != torch.uint8 or masks.dtype != torch.uint8:
        raise RuntimeError("stage_pairs: frames and masks must be uint8 (as decoded)")
    dev, S = frames_bgr.device, int(img_size)
    f32 = dict(dtype=torch.float32, device=dev)
    img = torch.empty(P, 2, 3, S, S, **f32); mask = torch.empty(P, 2, S, S, **f32)
    flow = torch.empty(P, 2, 3, S, S, **f32); occ = torch.empty(P, 2, S, S, **f32)
    focal = torch.zeros(P, 2, **f32); pps = torch.zeros(P, 2, 2, **f32)
    boxes = torch.zeros(P, 2, 4, dtype=torch.int32, device=dev)
    ws = torch.empty(int(_cabi.lib().lasr_b200_stage_workspace_bytes(P)), dtype=torch.uint8, device=dev)
    _cabi.native_call("stage_pairs", frames_bgr.contiguous(), masks.contiguous(), flows.contiguous().float(),
                      occs.contiguous().float(), img, mask, flow, occ, focal, pps, boxes, ws, ws.numel(), P, H, W, S)
    cam = torch.zeros(P, 2, 7, **f32)              # [1/alp, 0, 0, 1, 0, 0, 0]   (vidbase.py:167-183)
    cam[:, :, 3] = 1.0
    if cam_init is not None:                       # per-frame camera files: everything but the focal scale (:174-181)
        cam[:, :, 1:] = torch.as_tensor(cam_init, dtype=torch.float32).to(dev)[:, :, 1:]
    cam[:, :, 0] = focal
    return {"img": img[:, 0], "imgn": img[:, 1], "mask": mask, "flow": flow[:, 0], "flown": flow[:, 1],
            "occ": occ[:, 0], "occn": occ[:, 1], "cam": cam[:, 0], "camn": cam[:, 1], "pps": pps, "boxes": boxes}


def check_boxes(boxes):
    """The reference raises inside numpy when a mask is empty (``xid.max()`` of an empty array, vidbase.py:98);
    here the condition is detected from the staged pixel counts (one small D2H read)."""
    if bool(((boxes[..., 3] == 0) | (boxes[..., 2] <= 0)).any()):
        raise ValueError("stage_pairs: a frame has an empty or one-pixel-wide mask (the reference's crop box is "
                         "undefined / empty there); its pair was not staged")


class DeviceStager:
    """Pinned host buffers + side-stream upload + device preprocessing for batches of `pairs` items of HxW frames.

    Two pinned slots alternate, so the host can fill slot i+1 while the copy engine is still reading slot i; a slot is
    refilled only after the upload that last used it has completed (its event)."""

    def __init__(self, pairs, height, width, img_size, device="cuda"):
        self.P, self.H, self.W, self.S = int(pairs), int(height), int(width), int(img_size)
        self.device = torch.device(device)
        shp = (self.P, 2, self.H, self.W)
        mk = lambda: {"frames_bgr": torch.empty(shp + (3,), dtype=torch.uint8).pin_memory(),
                      "masks": torch.empty(shp, dtype=torch.uint8).pin_memory(),
                      "flows": torch.empty(shp + (3,), dtype=torch.float32).pin_memory(),
                      "occs": torch.empty(shp, dtype=torch.float32).pin_memory()}
        self._host = [mk(), mk()]
        self._uploaded = [None, None]            # event recorded after the H2D copies of each slot
        self._slot = 0
        self._dev = {k: torch.empty_like(v, device=self.device) for k, v in self._host[0].items()}
        self._stream = torch.cuda.Stream(device=self.device)
        self._done = torch.cuda.Event()
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self._host[0].values())

    def submit(self, raw):
        """raw: dict of numpy / CPU tensors with the keys of ``collate_raw``.  Starts upload + staging on the side
        stream and returns immediately; ``result()`` makes the current stream wait for it."""
        slot = self._slot
        self._slot ^= 1
        if self._uploaded[slot] is not None:       # the copy engine may still be reading this slot
            self._uploaded[slot].synchronize()
        host = self._host[slot]
        for k, buf in host.items():
            buf.copy_(torch.as_tensor(np.ascontiguousarray(raw[k])) if not torch.is_tensor(raw[k]) else raw[k])
        self._stream.wait_stream(torch.cuda.current_stream(self.device))     # consumers of the previous outputs
        with torch.cuda.stream(self._stream):
            for k, buf in host.items():
                self._dev[k].copy_(buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
            self._uploaded[slot] = ev
            self._out = stage_pairs(self._dev["frames_bgr"], self._dev["masks"], self._dev["flows"], self._dev["occs"],
                                    self.S, cam_init=raw.get("cam_init") if hasattr(raw, "get") else None)
            self._done.record(self._stream)
        return self

    def result(self):
        torch.cuda.current_stream(self.device).wait_event(self._done)
        for v in self._out.values():
            v.record_stream(torch.cuda.current_stream(self.device))
        return self._out


def pair_frame_indices(im0, direction, dframe=1):
    """(frame 0, frame n, flow of frame 0 is the forward flow?) for one dataset item, as the reference picks them
    (dataloader/vidbase.py:53-55, 85-90): direction 1 pairs im0 with im0 + dframe and reads flow-forward[im0] /
    flow-backward[im0 + dframe]; direction 0 pairs it with im0 - dframe and reads flow-backward[im0] /
    flow-forward[im0 - dframe]."""
    forward = int(direction) == 1
    return int(im0), int(im0) + (dframe if forward else -dframe), forward


class ResidentVideo:
    """A whole decoded video kept in HBM and staged there.

    LASR optimises one video of a few dozen frames for thousands of steps; decoded, an 80-frame 480p video with its
    flow and occlusion maps is ~0.65 GB -- it is uploaded ONCE.  The per-item preprocessing is deterministic (the
    reference's dataloader has no augmentation), so every (frame, direction) item is staged once, in chunks, on first
    use, and a batch is an ``index_select`` over the staged items: no PCIe traffic and no preprocessing per step.

    frames_bgr [T,H,W,3] uint8, masks [T,H,W] uint8, flow_fw / flow_bw [T,H,W,3] float32, occ_fw / occ_bw [T,H,W]
    float32 (numpy or tensors; flow_fw[t] is the flow t -> t + dframe, flow_bw[t] the flow t -> t - dframe)."""

    def __init__(self, frames_bgr, masks, flow_fw, occ_fw, flow_bw, occ_bw, img_size, dframe=1, device="cuda", chunk=16):
        dev = torch.device(device)
        up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a) if not torch.is_tensor(a) else a).to(dev, dt).contiguous()
        self.frames, self.masks = up(frames_bgr, torch.uint8), up(masks, torch.uint8)
        self.flow = {True: up(flow_fw, torch.float32), False: up(flow_bw, torch.float32)}
        self.occ = {True: up(occ_fw, torch.float32), False: up(occ_bw, torch.float32)}
        self.T, self.S, self.dframe, self.chunk, self.device = int(self.frames.shape[0]), int(img_size), int(dframe), int(chunk), dev
        self._staged = None          # dict of tensors over all 2*T (frame, direction) slots, filled on first use
        self._valid = None

    def _slot(self, im0, direction):
        return 2 * int(im0) + (1 if int(direction) == 1 else 0)

    def _stage_all(self):
        items = [(t, d) for t in range(self.T) for d in (0, 1)]
        ok = [0 <= pair_frame_indices(t, d, self.dframe)[1] < self.T for t, d in items]
        out = None
        todo = [it for it, v in zip(items, ok) if v]
        for s in range(0, len(todo), self.chunk):
            part = todo[s:s + self.chunk]
            sel = [pair_frame_indices(t, d, self.dframe) for t, d in part]
            i0 = torch.tensor([a for a, _, _ in sel], device=self.device)
            i1 = torch.tensor([b for _, b, _ in sel], device=self.device)
            fwd = torch.tensor([f for _, _, f in sel], device=self.device)
            f4, f3 = fwd[:, None, None, None], fwd[:, None, None]
            fr = torch.stack([self.frames[i0], self.frames[i1]], 1)
            mk = torch.stack([self.masks[i0], self.masks[i1]], 1)
            # frame 0 carries the flow towards its partner, the partner the flow back (vidbase.py:85-90)
            fl = torch.stack([torch.where(f4, self.flow[True][i0], self.flow[False][i0]),
                              torch.where(f4, self.flow[False][i1], self.flow[True][i1])], 1)
            oc = torch.stack([torch.where(f3, self.occ[True][i0], self.occ[False][i0]),
                              torch.where(f3, self.occ[False][i1], self.occ[True][i1])], 1)
            res = stage_pairs(fr, mk, fl, oc, self.S)
            check_boxes(res["boxes"])
            if out is None:
                out = {k: torch.zeros((2 * self.T,) + tuple(v.shape[1:]), dtype=v.dtype, device=self.device) for k, v in res.items()}
            slots = torch.tensor([self._slot(t, d) for t, d in part], device=self.device)
            for k, v in res.items():
                out[k][slots] = v
        self._staged = out
        self._valid = torch.tensor(ok, device=self.device)

    def stage(self, base_frames, directions):
        """The collated batch (keys / shapes of ``stage_pairs``) for items (base frame, direction)."""
        if self._staged is None:
            self._stage_all()
        slots = torch.tensor([self._slot(t, d) for t, d in zip(base_frames, directions)], device=self.device)
        if not bool(self._valid[slots].all()):
            raise IndexError("ResidentVideo.stage: an item pairs a frame with one outside the video")
        return {k: v.index_select(0, slots) for k, v in self._staged.items()}
