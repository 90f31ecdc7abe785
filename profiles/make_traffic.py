"""Extract per-launch DRAM traffic / issue statistics of the raster kernels from an `ncu --set full` report
into profiles/traffic_rN.json (read by bench.py for `roofline.traffic`).
usage: python profiles/make_traffic.py <report.ncu-rep> <out.json> <source note>"""
import csv, json, subprocess, sys

rep, out, note = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H, U = rows[0], rows[1]
ix = {h: i for i, h in enumerate(H)}


def val(r, name, scale_units=True):
    v = float(r[ix[name]].replace(",", ""))
    u = U[ix[name]]
    if scale_units:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v


kern = {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].replace("void ", "").replace("lasr::", "").split("(")[0]
    if name in kern:
        continue
    kern[name] = dict(dram_bytes_read=val(r, "dram__bytes_read.sum"), dram_bytes_write=val(r, "dram__bytes_write.sum"),
                      duration_ms=val(r, "gpu__time_duration.sum"),
                      issue_active_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                      threads_per_inst=val(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False),
                      warp_instructions=val(r, "smsp__inst_executed.sum", False),
                      registers=int(val(r, "launch__registers_per_thread", False)),
                      achieved_occupancy_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False))
json.dump({"source": note, "kernels": kern}, open(out, "w"), indent=1)
print(json.dumps(kern, indent=1))
