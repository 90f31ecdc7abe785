"""Dataloader staging: device path (pinned H2D + two kernels, `lasr_b200.dataloader.DeviceStager`) against the
reference's per-item CPU sequence (cv2.remap / cv2.resize / numpy, dataloader/vidbase.py:78-160 restated with the
same cv2 calls) on the box's host cores.  DAVIS-sized frames (480 x 854), img_size 256.
usage: python profiles/staging_timing.py [pairs] [reps]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from lasr_b200.dataloader import DeviceStager
from lasr_b200.dataloader.staging import stage_pairs

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
H, W, S = 480, 854, 256
rng = np.random.RandomState(0)
fr = rng.randint(0, 256, (P, 2, H, W, 3)).astype(np.uint8)
yy, xx = np.mgrid[0:H, 0:W]
mk = np.stack([np.stack([(((xx - 300 - 20 * p - 9 * s) / 140.0) ** 2 + ((yy - 230) / 120.0) ** 2 < 1).astype(np.uint8) * 255
                         for s in range(2)]) for p in range(P)])
fl = rng.normal(size=(P, 2, H, W, 3)).astype(np.float32); fl[..., 2] = rng.uniform(size=(P, 2, H, W)) > 0.1
oc = (rng.uniform(size=(P, 2, H, W)) * 20).astype(np.float32)
raw = {"frames_bgr": torch.from_numpy(fr), "masks": torch.from_numpy(mk), "flows": torch.from_numpy(fl), "occs": torch.from_numpy(oc)}


def cpu_item(p):
    """The reference's sequence for one item, same cv2 calls (vidbase.py:78-160)."""
    import cv2
    out = []
    boxes = []
    for s in range(2):
        img = fr[p, s][:, :, ::-1] / 255.0
        mask = mk[p, s][:, :, None]
        color = 1 - img[mask[:, :, 0].astype(bool)].mean(0)[None, None, :]
        img = img * (mask > 0).astype(float) + color * (1 - (mask > 0).astype(float))
        ind = np.where(mask > 0); xid, yid = ind[1], ind[0]
        c = ((xid.max() + xid.min()) // 2, (yid.max() + yid.min()) // 2)
        L = int(1.2 * max((xid.max() - xid.min()) // 2, (yid.max() - yid.min()) // 2))
        x0, y0 = np.meshgrid(range(2 * L), range(2 * L))
        x0 = (x0 + (c[0] - L)).astype(np.float32); y0 = (y0 + (c[1] - L)).astype(np.float32)
        img = cv2.remap(img, x0, y0, interpolation=cv2.INTER_LINEAR, borderValue=color[0, 0])
        m = cv2.remap(mask.astype(int), x0, y0, interpolation=cv2.INTER_NEAREST)
        f = cv2.remap(fl[p, s], x0, y0, interpolation=cv2.INTER_LINEAR)
        o = cv2.remap(oc[p, s], x0, y0, interpolation=cv2.INTER_LINEAR)
        img = cv2.resize(img, (S, S), interpolation=cv2.INTER_LINEAR)
        m = cv2.resize(m, (S, S), interpolation=cv2.INTER_NEAREST)
        f = cv2.resize(f, (S, S), interpolation=cv2.INTER_LINEAR)
        o = cv2.resize(o, (S, S), interpolation=cv2.INTER_LINEAR)
        out.append((img, m, f, o)); boxes.append((c, L))
    return out


st = DeviceStager(P, H, W, S)
for _ in range(3):
    st.submit(raw).result()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    st.submit(raw).result()
torch.cuda.synchronize()
t_e2e = (time.perf_counter() - t0) / reps * 1e3
dev = {k: v.cuda() for k, v in raw.items()}
for _ in range(3):
    stage_pairs(dev["frames_bgr"], dev["masks"], dev["flows"], dev["occs"], S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    stage_pairs(dev["frames_bgr"], dev["masks"], dev["flows"], dev["occs"], S)
e1.record(); torch.cuda.synchronize()
t_k = e0.elapsed_time(e1) / reps
in_bytes = st.h2d_bytes
out_bytes = P * 2 * S * S * 8 * 4
print(f"staging: {P} pairs of {H}x{W} -> {S}^2: device-resident {t_k:.3f} ms ({(in_bytes + out_bytes) / t_k / 1e6:.0f} GB/s over "
      f"{in_bytes / 1e6:.0f} MB in + {out_bytes / 1e6:.1f} MB out), with pinned H2D {t_e2e:.3f} ms ({2 * P / t_e2e * 1e3:.0f} frames/s)")
try:
    import cv2  # noqa: F401
    n = min(P, 4)
    t0 = time.perf_counter()
    for p in range(n):
        cpu_item(p)
    t_cpu = (time.perf_counter() - t0) / n * 1e3
    print(f"staging: reference-style CPU sequence (cv2 {cv2.__version__}, one worker, no file IO): {t_cpu:.1f} ms per item "
          f"-> {P * t_cpu:.0f} ms per batch of {P} ({2e3 / t_cpu:.0f} frames/s per worker; cores on this box: {os.cpu_count()})")
except ImportError:
    print("staging: cv2 not available for the CPU comparison")
