"""Turn the scratch ncu outputs of a gpurun trip into the committed per-round summary.
usage: python profiles/summarize.py <launches.csv> <report.ncu-rep> <bench.json> <out.md> <title>"""
import collections
import csv
import json
import subprocess
import sys


def launch_table(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| share | total ms | launches | avg us | kernel |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:14]:
        out.append(f"| {100*sum(v)/tot:.1f}% | {sum(v)/1e6:.3f} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | `{k}` |")
    return "\n".join(out), tot / 1e6


WANT = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def kernel_tables(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(H)}
    out = []
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        out.append(f"\n**`{name}`**\n\n| metric | value |\n|---|---|")
        for m, label in WANT:
            if m in idx:
                out.append(f"| {label} (`{m}`) | {r[idx[m]]} {U[idx[m]]} |")
    return "\n".join(out)


def main():
    launches, rep, benchjson, outmd, title = sys.argv[1:6]
    tab, tot = launch_table(launches)
    b = json.loads(open(benchjson).read().strip().splitlines()[-1])
    md = [f"# {title}", "",
          "Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3` under "
          "`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list) and "
          "`ncu --set full --clock-control none --import-source on -k regex:raster_|lbs_|vertex_gather` "
          "(kernel details).  Times under ncu are cold-cache and serialised: compare SHARES.", "",
          "## bench.py (no profiler), same build", "",
          "```json", json.dumps({k: b[k] for k in ("value", "unit", "ms_per_step", "e2e", "kernels", "roofline",
                                                   "clocks", "gpu_launches", "cpu_baseline") if k in b}, indent=1), "```", "",
          f"## Launch list (top kernels by device time; total {tot:.2f} ms over the profiled run)", "", tab, "",
          "## Kernel details (`ncu --set full`)", kernel_tables(rep), ""]
    open(outmd, "w").write("\n".join(md))
    print("wrote", outmd)


if __name__ == "__main__":
    main()
