#!/usr/bin/env python
"""bench.py -- frames/s of the LASR render hot path (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference|cpu] [--config A|B|C|D|E]

Configurations (BASELINE.json `configs`; SURVEY.md section 8 names them A-E).  Weak scaling: every rank works on
its own `frames` frames.

  A (default, the headline: configs[1])  Spot 2930 V / 5856 F, 256^2, 64 frames per GPU -- rasterize fwd+bwd
  B (configs[0]'s shape)                 icosphere-4 2562 V / 5120 F, 256^2, 64 frames per GPU -- rasterize fwd+bwd
  C (configs[2])                         Spot refined to 10 000 faces, 512^2, 80 frames, 36 bones -- the WHOLE render
                                         region of LASR.forward (skin weights -> LBS -> projection -> flow fw/bw +
                                         texture renders -> silhouette + flow + texture-L1 losses), fwd+bwd
  D (configs[3])                         icosphere-5 10242 V / 20480 F, 512^2, 32 frames per GPU (256 on 8 GPUs) --
                                         rasterize fwd+bwd
  E (configs[4])                         Spot, 26 bones (25 parts), 256^2, 64 frames per GPU -- the whole region, fwd+bwd

A "step" is one pass of the hot path over the rank's batch: for A / B / D one forward + one backward of the soft
rasterizer (LASR's training settings: euclidean / prod / softmax / vertex colours, sigma 1e-4, gamma 1e-2); for C / E
one forward + backward of the region (mesh_net.py:255-447).

  value   frames/s with inputs resident in HBM (A/B/D: through the C ABI of include/lasr_b200.h; C/E: through
          lasr_b200.hotpath.RenderHotPath replayed as a CUDA graph), CUDA-event timed, max over ranks.
  e2e     the same through the public Python API with HOST (pinned) buffers: every step copies that step's inputs
          host->device (A/B/D: vertices, vertex colours and the uint8 target image -- LASR's frames are uint8,
          dataloader/vidbase.py:52-193; the static topology is a model buffer and is uploaded once; C/E: the uint8
          frames and silhouettes, float flow / occlusion and the per-frame poses), runs the step, sums the gradients
          of the parameters the frames share (one NCCL all-reduce of one flat bucket when N > 1) and reads the losses
          and those gradients back to the host.  The protocol is the same for both arms; ours overlaps the copy of
          step i+1 with step i (lasr_b200.pipeline.HostPrefetcher) and replays the step as a CUDA graph, the reference
          arm copies at the top of the step as LASR's trainer does.
  roofline  the dominant kernel (raster_backward_kernel): algorithmic bytes (144F + 40S^2 per frame per colour-set
          pass) over its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs -- plus the binding resource
          (FP32 issue) from the committed ncu capture.
  cpu_baseline  the CPU port of the reference algorithm (oracle/) on the box's host cores, bounded sample, rank 0.

--impl reference runs the UNMODIFIED reference SoftRas CUDA kernels (oracle/_ref, compiled from /root/reference's
sources with a force-included host shim) behind a restatement of the reference's Python (oracle/ref_gpu.py) for the
same metric / config.  That arm imports nothing native of this repository: torch, oracle.ref_gpu and the numpy
workload builder only.  The reference has no CPU implementation; if oracle/_ref is
absent the arm falls back to the CPU port.  --impl cpu times the CPU port alone.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_DEFAULT = 256
FRAMES_PER_GPU = 64
METRIC = "rasterize fwd+bwd frames/s (Spot 5856 faces, 256x256, LASR settings)"
REGION_METRIC = "LASR render region fwd+bwd frames/s (skin -> LBS -> projection -> 3 renders -> 3 image losses)"

CONFIGS = {
    "A": dict(kind="raster", mesh="spot", S=256, frames=64, bones=1,
              label="BASELINE.json configs[1]: Spot 2930V/5856F, 256x256, 64 frames per GPU, "
                    "euclidean/prod/softmax/vertex, sigma 1e-4, gamma 1e-2"),
    "B": dict(kind="raster", mesh="ico4", S=256, frames=64, bones=1,
              label="BASELINE.json configs[0] shape: icosphere-4 2562V/5120F, 256x256, 64 frames per GPU, LASR settings"),
    "C": dict(kind="region", mesh="spot10k", S=512, frames=80, bones=36,
              label="BASELINE.json configs[2] shape: Spot refined to 5002V/10000F, 512x512, 80 frames, 36 bones, "
                    "whole render region with silhouette + flow + texture-L1 losses"),
    "D": dict(kind="raster", mesh="ico5", S=512, frames=32, bones=1,
              label="BASELINE.json configs[3]: icosphere-5 10242V/20480F, 512x512, 32 frames per GPU "
                    "(256-frame batch on 8 GPUs), LASR settings"),
    "E": dict(kind="region", mesh="spot", S=256, frames=64, bones=26,
              label="BASELINE.json configs[4] shape: Spot 2930V/5856F, 26 bones (25 parts), 256x256, 64 frames per GPU, "
                    "whole render region (skinning-gradient stress)"),
}
EYE = (0.0, 0.0, -(1.0 / math.tan(math.radians(30.0)) + 1.0))   # SoftRenderer's default look_at eye (renderer.py:60-63)


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic_files():
    import glob
    import re
    files = glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json"))
    # newest round first: traffic_r2 > traffic_r1g > traffic_r1
    return sorted(files, key=lambda p: [int(t) if t.isdigit() else t for t in re.findall(r"\d+|\D+", os.path.basename(p))],
                  reverse=True)


def profiled_issue_stats(kernel_name):
    """Issue-slot utilisation / lane efficiency of the kernel from the committed ncu capture (the binding
    resource on this path; the HBM fraction is low by nature)."""
    for path in _traffic_files():
        try:
            k = json.load(open(path))["kernels"]
        except Exception:
            continue
        for name, v in k.items():
            if name.startswith(kernel_name):
                return {"issue_slots_busy_pct": v.get("issue_active_pct"), "active_lanes_per_instruction": v.get("threads_per_inst"),
                        "warp_instructions": v.get("warp_instructions"), "registers": v.get("registers"),
                        "achieved_occupancy_pct": v.get("achieved_occupancy_pct"), "source": os.path.relpath(path, ROOT)}
    return None


def measured_traffic(kernel_name):
    """dram__bytes_read+write per launch of the dominant kernel, from the committed ncu capture of this
    same command/config (profiles/traffic_r*.json; newest round wins).  None if no capture matches."""
    for path in _traffic_files():
        try:
            k = json.load(open(path))["kernels"]
        except Exception:
            continue
        for name, v in k.items():
            if name.startswith(kernel_name):
                return v["dram_bytes_read"] + v["dram_bytes_write"], os.path.relpath(path, ROOT)
    return None, None


def build_workload(n_frames, S, seed, mesh="spot"):
    from lasr_b200 import synth
    v, f = synth.mesh_by_name(mesh)
    px, cam, near, far = synth.render_inputs(v, f, n_frames, seed)
    colors = np.repeat(synth.vertex_colors(len(v), 0)[None], n_frames, 0)
    settings = dict(image_size=S, background_color=[0, 0, 0], near=near, far=far, fill_back=True, eps=1e-3,
                    sigma_val=1e-4, dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-2,
                    aggr_func_rgb="softmax", aggr_func_alpha="prod", texture_type="vertex")
    rng = np.random.RandomState(1 + seed)
    grad = (rng.randn(n_frames, 4, S, S) * 1e-3).astype(np.float32)
    return dict(verts=v, faces=f, verts_px=px.astype(np.float32), colors=colors.astype(np.float32),
                settings=settings, grad=grad, near=near, far=far)


def cpu_baseline(wl, S, budget_s=12.0):
    """CPU port of the reference algorithm (brute-force all-faces scan, like the reference kernels),
    all host threads, bounded sample of the same workload."""
    from lasr_b200 import synth
    from oracle import oracle as O
    O.build()
    threads = O.max_threads()
    f = wl["faces"]

    def run(nf):
        fv = synth.gather_faces(wl["verts_px"][:nf], f)
        tex = synth.gather_faces(wl["colors"][:nf], f)
        t0 = time.perf_counter()
        r = O.soft_rasterize_forward(fv, tex, wl["settings"], np.float32, accel=False)
        O.soft_rasterize_backward(r, wl["grad"][:nf], accel=False)
        return time.perf_counter() - t0

    t2 = run(2)
    nf = int(max(2, min(len(wl["verts_px"]), round(2 * budget_s / max(t2, 1e-3)))))
    t = run(nf) if nf > 2 else t2
    return {"value": nf / t, "unit": "frames/s", "cores": threads, "kind": "port",
            "sample": f"{nf} frames of the workload, rasterize fwd+bwd, brute-force O(S^2*F) scan as the reference "
                      f"kernels do, OpenMP {threads} threads, {t:.2f} s"}


def timed_loop(torch, fn, iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ----------------------------------------------------------------------------------------------
# rasterize fwd+bwd arms (configs A / B / D)
# ----------------------------------------------------------------------------------------------
class RasterArmBase:
    """Buffers of one rasterize fwd+bwd step, allocated with torch only (no native code of either arm)."""

    def __init__(self, wl, S, device):
        import torch
        from lasr_b200 import synth
        self.torch = torch
        self.S, self.dev, self.wl = S, device, wl
        f = wl["faces"]
        N = len(wl["verts_px"])
        self.N, self.F = N, len(f)
        self.fv = torch.from_numpy(synth.gather_faces(wl["verts_px"], f)).to(device).reshape(N, self.F, 9).contiguous()
        self.tex = torch.from_numpy(synth.gather_faces(wl["colors"], f)).to(device).reshape(N, self.F, 3, 3).contiguous()
        self.grad = torch.from_numpy(wl["grad"]).to(device)
        self.aggrs = torch.empty(N, 2, S, S, device=device)
        self.img = torch.empty(N, 4, S, S, device=device)
        self.gf = torch.empty(N, self.F, 9, device=device)
        self.gt = torch.empty(N, self.F, 3, 3, device=device)
        self.eye = torch.tensor(EYE, dtype=torch.float32, device=device)
        self.pending = []

    def kernel_times(self):
        fwd = [e[0].elapsed_time(e[1]) for e in self.pending]
        bwd = [e[1].elapsed_time(e[2]) for e in self.pending]
        return float(np.mean(fwd)), float(np.mean(bwd))

    def launch_count(self):
        return 0

    # ---- e2e: one step from host buffers --------------------------------------------------------
    def e2e_setup(self, rank):
        torch, wl, N, S = self.torch, self.wl, self.N, self.S
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        rng = np.random.RandomState(7 + rank)
        self.h_in = [pin(wl["verts_px"]), pin(wl["colors"]),
                     pin(rng.randint(0, 256, (N, 4, S, S)).astype(np.uint8))]       # LASR's frames are uint8
        # the topology is static (a model buffer, ext_nnutils/mesh_net.py:171-185): uploaded once, outside the step
        self.d_faces = torch.from_numpy(np.repeat(wl["faces"][None], N, 0).astype(np.int32)).to(self.dev)
        V = wl["verts"].shape[0]
        self.h_out = [torch.empty(1).pin_memory(), torch.empty(V, 3).pin_memory(), torch.empty(V, 3).pin_memory()]
        nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
        return nbytes(self.h_in), nbytes(self.h_out)

    def e2e_compute(self, verts_buf, colors_buf, target_u8):
        verts = verts_buf.detach().requires_grad_(True)
        colors = colors_buf.detach().requires_grad_(True)
        target = target_u8.float().mul_(1.0 / 255.0)
        img = self.render(verts, self.d_faces, colors)
        loss = 0.5 * (img - target).pow(2).mean()
        loss.backward()
        # the mesh is shared by the frames: sum per-frame gradients
        return loss.detach().reshape(1), verts.grad.sum(0), colors.grad.sum(0)

    def describe_e2e(self):
        return "SoftRenderer.render_mesh + autograd, L2 image loss against a uint8 target, pinned host buffers, topology resident"


class OursArm(RasterArmBase):
    name = "ours"

    def __init__(self, wl, S, device):
        super().__init__(wl, S, device)
        from lasr_b200 import _cabi
        import lasr_b200.soft_renderer as sr
        self.cabi, self.sr = _cabi, sr
        self.ws = _cabi.raster_workspace(self.N, self.F, S, device)
        # LASR_B200_KEEP_WORKSPACE=compact: 80 B per face kept between the two calls instead of the workspace (an A/B switch)
        from lasr_b200.soft_renderer.functional import SoftRasterizeFunction
        self.compact = sys.modules[SoftRasterizeFunction.__module__].keep_mode() == "compact"
        self.saved = _cabi.saved_buffer(self.N, self.F, S, device) if self.compact else None
        self.bwd_ws = _cabi.raster_backward_scratch(self.N, self.F, S, device) if self.compact else self.ws
        s = wl["settings"]
        self.scal = (S, s["near"], s["far"], None, s["eps"], s["sigma_val"], 2, float(np.log(1.0 / s["dist_eps"] - 1.0)),
                     s["gamma_val"], 1, 2, 1, True)
        self.renderer = sr.SoftRenderer(image_size=S, sigma_val=s["sigma_val"], gamma_val=s["gamma_val"],
                                        camera_mode="look_at", perspective=False, aggr_func_rgb="softmax",
                                        light_mode="vertex", light_intensity_ambient=1.0,
                                        light_intensity_directionals=0.0)
        self.renderer.rasterizer.near = s["near"]
        self.renderer.rasterizer.far = s["far"]

    def launch_count(self):
        return self.cabi.launch_count()

    def device_step(self, timed):
        torch = self.torch
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        if timed:
            e[0].record()
        # the background colour travels as an argument (lasr_b200_raster_forward_bg): no pre-fill pass over the image --
        # the reference arm performs the fills its Python performs (soft_rasterize.py:50-53)
        self.cabi.raster_forward(self.fv, self.tex, self.aggrs, self.img, self.ws, *self.scal, background=[0.0, 0.0, 0.0],
                                 saved=self.saved)
        if timed:
            e[1].record()
        self.cabi.raster_backward(self.fv, self.tex, self.img, self.aggrs, self.gf, self.gt, self.grad, self.bwd_ws,
                                  *self.scal, setup_done=True, saved=self.saved)
        if timed:
            e[2].record()
            self.pending.append(e)

    def render(self, verts, faces, colors):
        # LASR's call pattern (nnutils/mesh_net.py:81-87): +eye offset, then look_at(eye) undoes it
        mesh = self.sr.Mesh(verts + self.eye[None, None], faces, textures=colors, texture_type="vertex")
        return self.renderer.render_mesh(mesh)

    def describe_e2e(self):
        return (super().describe_e2e() + ", H2D of step i+1 overlapped with step i (HostPrefetcher), step replayed as a "
                "CUDA graph (GraphedStep)")


class ReferenceArm(RasterArmBase):
    """The reference's own CUDA kernels (oracle/_ref) behind a restatement of its Python glue.  Nothing of
    lasr_b200's native side is imported or loaded here."""
    name = "reference"

    def __init__(self, wl, S, device):
        super().__init__(wl, S, device)
        from oracle import ref_gpu
        self.ref = ref_gpu
        self.ref.ext()
        self.finfo = self.torch.zeros(self.N, self.F, 27, device=device)

    def device_step(self, timed):
        torch = self.torch
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timed else None
        s = self.wl["settings"]
        args = (self.S, float(s["near"]), float(s["far"]), float(s["eps"]), float(s["sigma_val"]), 2,
                float(np.log(1.0 / s["dist_eps"] - 1.0)), float(s["gamma_val"]), 1, 2, 1, True)
        # buffer initialisation the reference's Python performs per call (soft_rasterize.py:47-53, 88-89)
        self.finfo.zero_()
        self.aggrs.zero_()
        self.img.fill_(1.0)
        self.img[:, :3] *= 0.0
        if timed:
            e[0].record()
        self.ref.ext().forward_soft_rasterize(self.fv, self.tex, self.finfo, self.aggrs, self.img, *args)
        if timed:
            e[1].record()
        self.gf.zero_()
        self.gt.zero_()
        self.ref.ext().backward_soft_rasterize(self.fv, self.tex, self.img, self.finfo, self.aggrs, self.gf, self.gt,
                                               self.grad, *args)
        if timed:
            e[2].record()
            self.pending.append(e)

    def render(self, verts, faces, colors):
        return self.ref.render_mesh_lasr(verts + self.eye[None, None], faces, colors, self.eye, self.wl["settings"])


# ----------------------------------------------------------------------------------------------
# whole-region arms (configs C / E): nnutils/mesh_net.py:255-447, forward + backward
# ----------------------------------------------------------------------------------------------
class RegionArmBase:
    def __init__(self, rw, device):
        import torch
        self.torch, self.dev, self.rw = torch, device, rw
        self.N, self.F, self.S, self.B, self.V = rw["N"], rw["F"], rw["S"], rw["B"], rw["V"]
        dv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.params = {k: dv(v) for k, v in rw["params"].items()}          # shared by every frame: resident
        self.param_names = list(self.params.keys())
        self.faces = dv(np.repeat(rw["faces"][None], self.N, 0).astype(np.int32))
        self.h_in = None
        order = ["imgs_u8", "masks_u8", "flow", "occ"]
        self.in_names = order + ["Rmat", "Tmat", "ppoint", "scale"]
        # only the two flow channels the losses read cross PCIe (mesh_net.py:400 uses flow[:, :2]; the third is the
        # loader's validity channel, dataloader/vidbase.py:151)
        host = dict(rw["batch"], flow=np.ascontiguousarray(rw["batch"]["flow"][:, :2]))
        self.host_arrays = [host[k] for k in order] + [rw["pose"][k] for k in ("Rmat", "Tmat", "ppoint", "scale")]
        self.resident = [dv(a) for a in self.host_arrays]
        self.pending = []
        self.flow3 = None
        self.eye = torch.tensor(EYE, dtype=torch.float32, device=device)

    def launch_count(self):
        return 0

    def kernel_times(self):
        return None, None

    def e2e_setup(self, rank):
        torch = self.torch
        self.h_in = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in self.host_arrays]
        self.h_out = [torch.empty(4).pin_memory()] + [torch.empty(p.shape).pin_memory() for p in self.params.values()]
        nbytes = lambda ts: sum(t.numel() * t.element_size() for t in ts)
        return nbytes(self.h_in), nbytes(self.h_out)

    def step(self, imgs_u8, masks_u8, flow, occ, Rmat, Tmat, ppoint, scale):
        """One forward + backward of the region from device buffers.  Returns (4 losses, *shared-parameter grads)."""
        torch = self.torch
        if self.flow3 is None:        # [N,3,S,S] as the loader lays it out; the validity channel stays zero (never read)
            self.flow3 = torch.zeros(flow.shape[0], 3, flow.shape[2], flow.shape[3], device=flow.device)
        self.flow3[:, :2] = flow
        batch = dict(imgs=torch.mul(imgs_u8, 1.0 / 255.0), masks=masks_u8.float(), flow=self.flow3, occ=occ)
        P = {k: v.detach().requires_grad_(True) for k, v in self.params.items()}
        pose = [t.detach().requires_grad_(True) for t in (Rmat, Tmat, ppoint, scale)]
        losses = self.region(P, pose, batch)
        losses[0].backward()
        return (torch.stack([l.detach() for l in losses]),) + tuple(P[k].grad for k in self.param_names)

    def e2e_compute(self, *bufs):
        return self.step(*bufs)

    def device_step(self, timed):
        raise NotImplementedError


class OursRegionArm(RegionArmBase):
    name = "ours"

    def __init__(self, rw, device, no_graph=False):
        super().__init__(rw, device)
        from lasr_b200 import _cabi, hotpath
        self.cabi = _cabi
        self.model = hotpath.RenderHotPath(hotpath.HotPathOptions(img_size=self.S, n_hypo=1, n_bones=self.B, fused=True,
                                                                  report_flow=False)).to(device)
        self.no_graph = no_graph
        self.graph = None
        self.replays, self.launches_per_step = 0, 0

    def launch_count(self):
        # kernels replayed inside a CUDA graph do not pass through the library's launch counter: count them per step
        return self.cabi.launch_count() + self.replays * self.launches_per_step

    def region(self, P, pose, batch):
        N = self.N
        pred_v = P["mean_v"][None].repeat(N, 1, 1)            # the mean shape, repeated per frame (ext mesh_net.py:181)
        tex = P["tex"][None].repeat(N, 1, 1)
        bones = {k: P[k] for k in ("ctl_ts", "ctl_rs", "log_ctl", "rest_ts")} if self.B > 1 else None
        o = self.model(pred_v, tex, self.faces, pose[0], pose[1], pose[2], pose[3], batch, bones)
        return [o["total"], o["mask_loss"], o["flow_loss"], o["texture_loss"]]

    def device_step(self, timed):
        if self.no_graph:
            self.step(*self.resident)
            return
        if self.graph is None:
            from lasr_b200.pipeline import GraphedStep
            c0 = self.cabi.launch_count()
            self.step(*self.resident)                          # one eager step: this library's launches per step
            self.launches_per_step = int(self.cabi.launch_count() - c0)
            self.graph = GraphedStep(self.step, self.resident)
        self.graph()
        self.replays += 1

    def describe_e2e(self):
        return ("lasr_b200.hotpath.RenderHotPath (fused: one three-set render + fused head) + autograd, pinned host "
                "buffers (uint8 frames / silhouettes, float flow / occlusion, per-frame poses), shared parameters "
                "resident, H2D of step i+1 overlapped with step i, step replayed as a CUDA graph")


class ReferenceRegionArm(RegionArmBase):
    """The same region restated on the reference's kernels and Python (per-bone matmuls, three renders, Python
    loss loops with their host syncs): oracle/ref_gpu.lasr_region_reference."""
    name = "reference"

    def __init__(self, rw, device):
        super().__init__(rw, device)
        from oracle import ref_gpu
        self.ref = ref_gpu
        self.ref.ext()

    def region(self, P, pose, batch):
        N, B = self.N, self.B
        pred_v = P["mean_v"][None].repeat(N, 1, 1)
        tex = P["tex"][None].repeat(N, 1, 1)
        Rm, Tm, skin = pose[0], pose[1], None
        if B > 1:
            skin, Rm, Tm = self.ref.ref_skin_and_bones(pred_v, Rm, Tm, P, N, 1, B)
        total, m, f, t = self.ref.lasr_region_reference(pred_v, tex, self.faces, Rm, Tm, pose[2], pose[3], batch, 1, B,
                                                        self.S, skin, self.eye)
        return [total, m, f, t]

    def device_step(self, timed):
        self.step(*self.resident)

    def describe_e2e(self):
        return ("restatement of nnutils/mesh_net.py:259-447 on the reference's kernels (oracle/ref_gpu.py) + autograd, "
                "same pinned host buffers, copies at the top of the step")


# ----------------------------------------------------------------------------------------------
# extras (1 GPU): the helper kernels alone, and the whole region at config E's shape on the default line
# ----------------------------------------------------------------------------------------------
def extras_microbench(device, peak_gbs, iters=20):
    """The HBM-bound helper kernels of the path, timed alone with CUDA events (config E shape for the
    LBS: Spot, 64 frames, 26 bones; config A shape for the losses) against their algorithmic bytes.  Everything runs on
    ONE side stream: autograd replays a backward on the stream its forward ran on, so a backward is captured into a
    CUDA graph only if the forward ran on the capturing stream."""
    import torch
    side = torch.cuda.Stream(device=device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        out = _extras_microbench_on(side, device, peak_gbs, iters)
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize()
    return out


def _extras_microbench_on(side, device, peak_gbs, iters):
    import torch
    from lasr_b200 import native_geom, losses, synth
    from lasr_b200.geom_utils import quaternion_to_rotation_matrix
    out = {}
    g = torch.Generator().manual_seed(0)
    v, _ = synth.load_spot()
    N, V, B, S, H = 64, len(v), 26, 256, 1
    verts = torch.from_numpy(v)[None].to(device).requires_grad_(True)
    R = quaternion_to_rotation_matrix(torch.randn(N * B, 4, generator=g)).to(device).requires_grad_(True)
    T = (torch.randn(N * B, 3, generator=g) * 0.02)
    T.view(N, B, 3)[:, 0, 2] += 10
    T = T.to(device).requires_grad_(True)
    skin = torch.softmax(torch.randn(1, B - 1, V, 1, generator=g), 1).to(device).requires_grad_(True)
    pp = torch.zeros(N, 2, device=device, requires_grad=True)
    fl = torch.full((N,), 10.0, device=device, requires_grad=True)

    def timeit(fn):
        """Device time of fn: replayed as a CUDA graph (these calls are tens of microseconds of GPU work behind ~100 us
        of eager-mode Python -- timed launch by launch the loop measures the host); eager if the call cannot be captured."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                fn()
            graph.replay()
            torch.cuda.synchronize()
            t = timed_loop(torch, graph.replay, iters)
            if t < 1e-3:          # an empty graph (nothing was captured) replays in no time: not a measurement
                raise RuntimeError("empty graph")
            return t
        except Exception:
            torch.cuda.synchronize()
            return timed_loop(torch, fn, iters)

    with torch.no_grad():
        t_f = timeit(lambda: native_geom.skin_project(verts, R, T, B, skin, pp, fl, depth_window=True))
    cam, px, _ = native_geom.skin_project(verts, R, T, B, skin, pp, fl, depth_window=True)
    gc, gp = torch.randn_like(cam), torch.randn_like(px)
    for _ in range(3):
        torch.autograd.grad([cam, px], [verts, R, T, skin, pp, fl], [gc, gp], retain_graph=True)
    torch.cuda.synchronize()      # eager: six freshly allocated gradients per call do not replay well from a graph
    t_fb = timed_loop(torch, lambda: torch.autograd.grad([cam, px], [verts, R, T, skin, pp, fl], [gc, gp], retain_graph=True), iters)
    lbs_fwd_bytes = V * 12 + (B - 1) * V * 4 + N * (B * 48 + 12) + N * V * 28       # shared mesh+skin in, cam+px out
    lbs_bwd_bytes = V * 12 + (B - 1) * V * 4 + N * B * 48 + N * V * 28 + N * V * 12 + N * (B - 1) * V * 4
    out["lbs_project"] = {"shape": f"N={N} V={V} B={B} (shared mesh + weights)", "forward_ms": t_f, "backward_ms": t_fb,
                          "forward_GBps": lbs_fwd_bytes / t_f / 1e6, "backward_GBps": lbs_bwd_bytes / t_fb / 1e6,
                          "forward_frac_of_hbm_peak": lbs_fwd_bytes / t_f / 1e6 / peak_gbs,
                          "backward_frac_of_hbm_peak": lbs_bwd_bytes / t_fb / 1e6 / peak_gbs,
                          "note": "backward includes the per-frame -> shared reductions done by torch.sum"}
    mk = lambda *sh: torch.rand(*sh, generator=g).to(device)
    mp, fg, tex = mk(N, H, S, S).requires_grad_(True), mk(N, H, S, S).requires_grad_(True), mk(N, H, 3, S, S).requires_grad_(True)
    fr = (torch.randn(N, H, S, S, 2, generator=g) * 0.05).to(device).requires_grad_(True)
    masks, occ = (mk(N, S, S) > 0.5).float(), torch.randn(N, S, S, generator=g).to(device)
    bg, fo, imgs = mk(N, H, S, S) < 0.4, mk(N, 3, S, S), mk(N, 3, S, S)
    with torch.no_grad():
        t_lf = timeit(lambda: losses.image_losses(mp, masks, occ, fr, bg, fo, tex, fg, imgs))
    lm, lf, lt, _ = losses.image_losses(mp, masks, occ, fr, bg, fo, tex, fg, imgs)
    total = lm + lf + lt
    t_lb = timeit(lambda: torch.autograd.grad(total, [mp, fr, tex, fg], retain_graph=True))
    px_n = N * H * S * S
    loss_fwd_bytes = px_n * (4 + 8 + 1 + 12 + 4) + N * S * S * (12 + 4 + 4 + 8)
    loss_bwd_bytes = loss_fwd_bytes + px_n * (4 + 8 + 12 + 4)
    out["image_losses"] = {"shape": f"N={N} H={H} S={S}", "forward_ms": t_lf, "backward_ms": t_lb,
                           "forward_GBps": loss_fwd_bytes / t_lf / 1e6, "backward_GBps": loss_bwd_bytes / t_lb / 1e6,
                           "forward_frac_of_hbm_peak": loss_fwd_bytes / t_lf / 1e6 / peak_gbs,
                           "backward_frac_of_hbm_peak": loss_bwd_bytes / t_lb / 1e6 / peak_gbs,
                           "note": "two kernels forward (per-pixel sums; reduce + finalize), two backward (coefficients; per-pixel gradients)"}
    # dataloader -> device staging: 16 pairs of 480x854 frames -> 256^2, inputs resident
    from lasr_b200.dataloader.staging import stage_pairs
    P, Hh, Ww = 16, 480, 854
    frs = torch.randint(0, 256, (P, 2, Hh, Ww, 3), generator=g, dtype=torch.uint8).to(device)
    yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    mks = torch.stack([torch.stack([((((xx - 300 - 20 * p - 9 * q) / 140.0) ** 2 + ((yy - 230) / 120.0) ** 2) < 1).to(torch.uint8) * 255
                                    for q in range(2)]) for p in range(P)]).to(device)
    fls, ocs = torch.randn(P, 2, Hh, Ww, 3, generator=g).to(device), (torch.rand(P, 2, Hh, Ww, generator=g) * 20).to(device)
    t_st = timeit(lambda: stage_pairs(frs, mks, fls, ocs, S))
    bx = stage_pairs(frs, mks, fls, ocs, S)["boxes"].cpu().numpy().reshape(-1, 4)
    crop_px = sum(min(2 * int(b[2]), Ww) * min(2 * int(b[2]), Hh) for b in bx)
    st_bytes = 2 * P * Hh * Ww + 3 * int(bx[:, 3].sum()) + crop_px * 20 + 2 * P * S * S * 32
    out["staging"] = {"shape": f"{P} pairs of {Hh}x{Ww} -> {S}^2", "ms": t_st, "GBps": st_bytes / t_st / 1e6,
                      "frac_of_hbm_peak": st_bytes / t_st / 1e6 / peak_gbs,
                      "bytes": "masks + masked frame bytes + crop windows x 20 B + outputs x 32 B (compulsory traffic)"}
    from lasr_b200.chamfer3D import dist_chamfer_3D as C
    cham = C.chamfer_3DDist()
    pts = torch.from_numpy(v)[None].to(device)
    t_ch = timeit(lambda: cham(pts, pts * torch.tensor([-1.0, 1.0, 1.0], device=device)))
    out["chamfer"] = {"shape": f"B=1 n=m={V} (nnutils/mesh_net.py:474-477)", "ms": t_ch,
                      "pairs_per_s": 2.0 * V * V / t_ch * 1e3, "note": "launch / host bound at this size"}
    return out


def hot_path_extra(impl, device, iters=None, frames=64, bones=26, S=256, mesh="spot"):
    """The whole render region of LASR.forward at config E's shape (Spot, 64 frames, 26 bones, 256^2), device
    resident, on the default 1-GPU line of either arm: ours as one CUDA graph (fused three-set render + fused
    head), the reference arm as the reference-style restatement on the reference's kernels (few iterations:
    it takes hundreds of milliseconds per step)."""
    import torch
    from lasr_b200 import synth
    rw = synth.region_workload(mesh, frames, bones, S, seed=0)
    if impl == "ours":
        arm = OursRegionArm(rw, device)
        iters = iters or 15
    else:
        arm = ReferenceRegionArm(rw, device)
        iters = iters or 3
    for _ in range(3 if impl == "ours" else 1):
        arm.device_step(False)
    torch.cuda.synchronize()
    l0 = arm.launch_count()
    w0 = time.perf_counter()
    ms = timed_loop(torch, lambda: arm.device_step(False), iters)
    ms = max(ms, (time.perf_counter() - w0) * 1e3 / iters)
    res = {"shape": f"{mesh}, {frames} frames, {bones} bones, {S}^2 (config E shape), device-resident inputs, fwd+bwd",
           "ms_per_step": ms, "frames_per_s": frames / ms * 1e3, "iters": iters,
           "how": ("RenderHotPath(fused) replayed as a CUDA graph" if impl == "ours" else
                   "oracle/ref_gpu.lasr_region_reference: reference kernels + reference-style Python (loops, syncs)")}
    if impl == "ours":
        res["own_kernel_launches_per_step"] = arm.launches_per_step
    return res


# ----------------------------------------------------------------------------------------------
# the measurement proper
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args, impl):
    import torch
    rank, world, local = dist_env()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist = None
        torch.cuda.set_device(0)
    device = torch.device("cuda", local if world > 1 else 0)
    cfg = CONFIGS[args.config]
    S, kind = args.image_size, cfg["kind"]
    from lasr_b200 import synth                      # numpy-only workload builder
    if kind == "raster":
        wl = build_workload(args.frames, S, seed=rank, mesh=args.mesh)
        arm = (OursArm if impl == "ours" else ReferenceArm)(wl, S, device)
    else:
        wl = None
        rw = synth.region_workload(args.mesh, args.frames, args.bones, S, seed=rank)
        arm = OursRegionArm(rw, device, args.no_graph) if impl == "ours" else ReferenceRegionArm(rw, device)
    N, F = arm.N, arm.F

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    for _ in range(args.warmup):
        arm.device_step(False)
    barrier()
    sampler = ClockSampler(device.index or 0)
    sampler.start()
    launches0 = arm.launch_count()
    w0 = time.perf_counter()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        arm.device_step(True)
    t1.record()
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    launches = arm.launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    if kind == "region" and impl != "ours":
        ms_total = max(ms_total, wall_ms)            # the reference-style region has host syncs inside the step
    fwd_ms, bwd_ms = arm.kernel_times()

    # ---------------- end-to-end through the public API with host buffers ----------------
    h2d, d2h = arm.e2e_setup(rank)
    prefetch = None
    if impl == "ours":
        from lasr_b200.pipeline import HostPrefetcher
        prefetch = HostPrefetcher(device)
    graphs = {}

    def e2e_step():
        if prefetch is not None:
            if prefetch.staged == 0:
                prefetch.stage(arm.h_in)
            bufs = prefetch.get()
            prefetch.stage(arm.h_in)                          # next step's inputs, on the copy stream
            if args.no_graph:
                outs = arm.e2e_compute(*bufs)
            else:                                             # one CUDA graph per staging slot (static addresses)
                key = bufs[0].data_ptr()
                if key not in graphs:
                    from lasr_b200.pipeline import GraphedStep
                    graphs[key] = GraphedStep(arm.e2e_compute, bufs)
                outs = graphs[key]()
        else:
            outs = arm.e2e_compute(*[h.to(device, non_blocking=True) for h in arm.h_in])
        loss, grads = outs[0], list(outs[1:])
        if dist is not None:        # one flat bucket, one all-reduce
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat)
            off, red = 0, []
            for g in grads:
                red.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            grads = red
        arm.h_out[0].copy_(loss, non_blocking=True)
        for h, g in zip(arm.h_out[1:], grads):
            h.copy_(g, non_blocking=True)
        if prefetch is not None:
            prefetch.release()

    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(max(3, min(args.warmup, 5))):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(e2e_steps):
        e2e_step()
    s1.record()
    barrier()
    e2e_ms = max(s0.elapsed_time(s1), (time.perf_counter() - w0) * 1e3) / e2e_steps
    clocks = sampler.stop()                                   # sampled over both timed regions

    # max over ranks
    vals = torch.tensor([ms_total, e2e_ms, fwd_ms or 0.0, bwd_ms or 0.0], device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, fwd_ms, bwd_ms = [float(x) for x in vals.tolist()]
    ms_per_step = ms_total / args.steps
    total_frames = N * world

    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        headline = args.config == "A" and args.mesh == "spot" and S == 256 and N == 64
        label = cfg["label"] if (args.mesh == cfg["mesh"] and S == cfg["S"] and N == cfg["frames"] and
                                 args.bones == cfg["bones"]) else (
            f"non-standard shape of config {args.config}: {args.mesh} {arm.V if kind == 'region' else wl['verts'].shape[0]}V/"
            f"{F}F, {S}x{S}, {N} frames per GPU, {args.bones} bones, LASR settings")
        work_mb = (N * (4 + 2 + 4 + 4) * S * S * 4 + N * F * 72 * 2) / 1e6 if kind == "raster" else N * 30 * S * S * 4 / 1e6
        out = {
            "metric": METRIC if kind == "raster" else REGION_METRIC,
            "value": total_frames / (ms_per_step * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": impl,
            "config": {"workload": label, "name": args.config, "frames_per_gpu": N, "faces": F, "image_size": S,
                       "bones": args.bones, "parallelism": f"frame-shard x{world}",
                       "l2": f"per-step working set ~{work_mb:.0f} MB (inputs+outputs+grads) exceeds the 126 MB L2; no explicit flush"
                             if work_mb > 126 else f"per-step working set ~{work_mb:.0f} MB fits the 126 MB L2 (small non-headline shape)"},
            "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps, "api": arm.describe_e2e()},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if kind == "raster":
            bwd_bytes = (144 * F + 40 * S * S) * N
            fwd_bytes = (72 * F + 24 * S * S) * N
            ach = bwd_bytes / (bwd_ms * 1e-3) / 1e9
            traffic, traffic_src = measured_traffic("raster_backward_kernel") if (impl == "ours" and headline) else (None, None)
            out["kernels"] = {"forward_ms": fwd_ms, "backward_ms": bwd_ms,
                              "forward_GBps_algorithmic": fwd_bytes / (fwd_ms * 1e-3) / 1e9}
            out["roofline"] = {
                "bound": "hbm", "binds_in_practice": "fp32-issue" if impl == "ours" else "fp32/fp64-issue (brute-force face loop)",
                "kernel": "raster_backward_kernel" if impl == "ours" else "backward_soft_rasterize_cuda_kernel",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": bwd_bytes,
                "binding_resource": profiled_issue_stats("raster_backward_kernel") if (impl == "ours" and headline) else None,
                "note": ("BASELINE.json names the HBM roofline, so `bound`/`frac` are the HBM figures; what binds after binning is "
                         "FP32 issue (ncu: issue slots ~60-70% busy, DRAM <5% of peak; DESIGN.md section 4); measured DRAM traffic "
                         "is below the algorithmic bytes because the images the forward just wrote are still L2-resident")
                if impl == "ours" else
                        ("the reference's per-pixel loop over ALL faces (N*S^2*F iterations, with FP64 promotions, .cu:370,538): "
                         "instruction bound by construction, its face reads are warp-uniform L1/L2 broadcasts; HBM figures given "
                         "for the same algorithmic bytes as our arm")}
            try:   # the binding resource as a fraction: warp instructions of the profiled launch over the LIVE kernel time,
                   # against the issue peak (SMs x 4 schedulers x 1 instruction per clock at the sampled SM clock)
                br = out["roofline"]["binding_resource"]
                if br and br.get("warp_instructions"):
                    props = torch.cuda.get_device_properties(device)
                    mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
                    peak_issue = props.multi_processor_count * 4 * mhz * 1e6
                    ach_issue = float(br["warp_instructions"]) / (bwd_ms * 1e-3)
                    br.update(achieved_warp_inst_per_s=ach_issue, peak_warp_inst_per_s=peak_issue,
                              issue_frac=ach_issue / peak_issue)
            except Exception:
                pass
            try:   # SURVEY 8d timing protocol: spread of the per-step kernel times (CUDA events of every timed step)
                def spread(xs):
                    return {"median": float(np.median(xs)), "p10": float(np.percentile(xs, 10)),
                            "p90": float(np.percentile(xs, 90)), "n": len(xs)}
                out["step_stats_ms"] = {
                    "forward": spread([e[0].elapsed_time(e[1]) for e in arm.pending]),
                    "backward": spread([e[1].elapsed_time(e[2]) for e in arm.pending]),
                    "forward_plus_backward": spread([e[0].elapsed_time(e[2]) for e in arm.pending])}
            except Exception:
                pass
        else:
            # three colour sets ride one geometry pass on our arm; the algorithmic bytes of the region's dominant kernel
            # are those of ONE backward pass over the 10-plane image (3 sets + alpha) and the face gradients
            out["roofline"] = {"bound": "hbm", "binds_in_practice": "fp32-issue", "kernel": "raster_backward_kernel<3 sets>"
                               if impl == "ours" else "backward_soft_rasterize_cuda_kernel", "achieved": None, "peak": peak,
                               "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
                               "note": "whole-region line: per-kernel times are in profiles/ (ncu launch list); the headline "
                                       "roofline object is on the config A line"}
        if world == 1 and not args.no_extras and headline:
            try:
                out.setdefault("extras", {})["hot_path"] = hot_path_extra(impl, device)
            except Exception as ex:   # never lose the headline line over an extra
                out.setdefault("extras", {})["hot_path"] = {"error": repr(ex)[:300]}
        if world == 1 and impl == "ours" and args.extras:
            out.setdefault("extras", {}).update(extras_microbench(device, peak))
        if world == 1 and not args.no_cpu_baseline and kind == "raster":
            out["cpu_baseline"] = cpu_baseline(wl, S, args.cpu_budget)
        elif world == 1 and not args.no_cpu_baseline:
            cw = build_workload(2, S, seed=0, mesh=args.mesh)
            cb = cpu_baseline(cw, S, min(args.cpu_budget, 6.0))
            cb["sample"] = "ONE of the region's three rasterize passes only (the CPU port covers the rasterizer): " + cb["sample"]
            out["cpu_baseline"] = cb
        if impl == "reference":
            # the reference arm must not have touched this repository's native code: checked on the process's own map
            try:
                mapped = open("/proc/self/maps").read()
                out["native_code_of_this_repo_loaded"] = "liblasr_b200" in mapped or "lasr_b200._cabi" in sys.modules
            except OSError:
                pass
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_cpu_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    S = args.image_size
    wl = build_workload(args.frames, S, seed=0, mesh=args.mesh)
    vals = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(wl, S, budget_s=max(1.0, args.cpu_budget / 2))
        if i >= args.warmup:
            vals.append(cb["value"])
    v = float(np.mean(vals))
    cb["value"] = v
    out = {"metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * args.frames / v, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference" if args.impl == "reference" else "cpu",
           "config": {"workload": CONFIGS[args.config]["label"] + " (bounded sample on host cores)", "name": args.config,
                      "frames_per_gpu": args.frames, "image_size": S},
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0,
           "note": "CPU port of the reference algorithm (the reference itself has no CPU path)"}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu"])
    ap.add_argument("--config", default="A", choices=sorted(CONFIGS), help="BASELINE.json configuration (default A, the headline)")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU (default: the configuration's)")
    ap.add_argument("--image-size", type=int, default=None)
    ap.add_argument("--mesh", default=None, choices=["spot", "spot10k", "ico3", "ico4", "ico5"],
                    help="override the configuration's mesh (non-headline shapes)")
    ap.add_argument("--bones", type=int, default=None, help="bones of the region configurations (C / E)")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying CUDA graphs")
    ap.add_argument("--no-extras", action="store_true", help="skip extras.hot_path on the default 1-GPU line")
    ap.add_argument("--extras", action="store_true", help="also time the LBS/projection, image-loss, staging and chamfer kernels alone")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    args.frames = args.frames or cfg["frames"]
    args.image_size = args.image_size or cfg["S"]
    args.mesh = args.mesh or cfg["mesh"]
    args.bones = args.bones or cfg["bones"]

    if args.impl == "cpu":
        return run_cpu_arm(args)
    if args.impl == "reference":
        from oracle import ref_gpu
        if not ref_gpu.available():
            return run_cpu_arm(args)
        if args.steps > 20:
            args.steps = 20        # the reference's brute-force kernels take ~80 ms per step at config A; bounded
        return run_gpu_arm(args, "reference")
    return run_gpu_arm(args, "ours")


if __name__ == "__main__":
    main()
