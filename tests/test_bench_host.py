"""bench.py's host side without a GPU: the workload builder, the committed-profile readers and the CPU arm's JSON line
(`--impl cpu`, which is also what `--impl reference` falls back to when oracle/_ref is absent)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_workload_and_profile_readers():
    import bench
    wl = bench.build_workload(3, 32, seed=0)
    assert wl["verts_px"].shape == (3, 2930, 3) and wl["faces"].shape == (5856, 3) and wl["grad"].shape == (3, 4, 32, 32)
    assert wl["settings"]["sigma_val"] == 1e-4 and wl["settings"]["gamma_val"] == 1e-2 and wl["near"] < wl["far"]
    peak, src = bench.measured_peaks()
    assert 3000 < peak < 9000 and src
    traffic, tsrc = bench.measured_traffic("raster_backward_kernel")
    assert traffic and traffic > 1e7 and tsrc.startswith("profiles/")
    st = bench.profiled_issue_stats("raster_backward_kernel")
    assert st and 0 < st["issue_slots_busy_pct"] <= 100 and 0 < st["active_lanes_per_instruction"] <= 32
    json.dumps(st)


def test_cpu_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "cpu", "--steps", "1", "--warmup", "1",
                        "--frames", "2", "--image-size", "32", "--cpu-budget", "1"], capture_output=True, text=True,
                       env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "cpu" and line["unit"] == "frames/s" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
