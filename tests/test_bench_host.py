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


def test_config_table_and_region_workload():
    """bench.py's BASELINE configurations and the numpy region workload (no GPU): shapes, dtypes as LASR's loader
    decodes them (uint8 frames / silhouettes), frame pairs, shared bone parameters independent of the rank."""
    import numpy as np
    import bench
    from lasr_b200 import synth
    assert sorted(bench.CONFIGS) == ["A", "B", "C", "D", "E"]
    assert bench.CONFIGS["A"]["kind"] == "raster" and bench.CONFIGS["A"]["frames"] == 64 and bench.CONFIGS["A"]["S"] == 256
    assert bench.CONFIGS["D"]["mesh"] == "ico5" and bench.CONFIGS["D"]["S"] == 512 and bench.CONFIGS["D"]["frames"] == 32
    assert bench.CONFIGS["E"]["kind"] == "region" and bench.CONFIGS["E"]["bones"] == 26
    w0, w1 = synth.region_workload("ico2", 4, 3, 16, seed=0), synth.region_workload("ico2", 4, 3, 16, seed=1)
    assert w0["batch"]["imgs_u8"].dtype == np.uint8 and w0["batch"]["imgs_u8"].shape == (4, 3, 16, 16)
    assert w0["batch"]["masks_u8"].dtype == np.uint8 and set(np.unique(w0["batch"]["masks_u8"])) <= {0, 1}
    assert w0["pose"]["Rmat"].shape == (12, 3, 3) and w0["pose"]["Tmat"].shape == (12, 3)
    r = w0["pose"]["Rmat"]
    assert np.allclose(np.einsum("nij,nkj->nik", r, r), np.eye(3)[None], atol=1e-5)           # rotations
    for k in ("mean_v", "tex", "ctl_ts", "rest_ts", "ctl_rs", "log_ctl"):
        assert np.array_equal(w0["params"][k], w1["params"][k]), k                           # shared across ranks
    assert not np.array_equal(w0["pose"]["Rmat"], w1["pose"]["Rmat"])                         # frames are the rank's own
    with __import__("pytest").raises(ValueError):
        synth.region_workload("ico2", 3, 1, 16)                                               # pairs: even frame count


def test_refined_spot_is_a_closed_manifold_with_exactly_ten_thousand_faces():
    """BASELINE config C's "10k-face mesh": Spot + 2072 split edges (synth.refine_to_faces)."""
    from collections import Counter
    from lasr_b200 import synth
    v, f = synth.mesh_by_name("spot10k")
    assert f.shape == (10000, 3) and v.shape == (5002, 3) and f.max() == len(v) - 1
    half_edges = Counter()
    for a, b, c in f.tolist():
        for e in ((a, b), (b, c), (c, a)):
            half_edges[e] += 1
    assert all(n == 1 for n in half_edges.values())                    # every directed edge once ...
    assert all((b, a) in half_edges for (a, b) in half_edges)          # ... and its opposite too: closed, consistently oriented
    v0, f0 = synth.load_spot()
    assert (v[:len(v0)] == v0).all()                                   # the original vertices are kept
