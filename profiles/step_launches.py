"""Per-kernel time of ONE hot-path step from an ncu launch list (gpu__time_duration.sum, --csv).
usage: python profiles/step_launches.py <launches.csv> [anchor-kernel-substring] [--all]"""
import collections, csv, re, sys

path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "raster_forward_kernel"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = []
for x in csv.DictReader(lines):
    if x.get("Metric Name") == "gpu__time_duration.sum":
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
        rows.append((x["Kernel Name"], v, x.get("Grid Size")))
idx = [i for i, (k, *_) in enumerate(rows) if anchor in k]
a, b = idx[-2], idx[-1]
step = rows[a:b]
tot = sum(v for _, v, _ in step)
print(f"{len(rows)} launches captured; one step = {len(step)} launches, {tot:.1f} us (serialised, cold cache)")
agg = collections.OrderedDict()
for k, v, _ in step:
    k = re.sub(r"void |at::|native::|lasr::|<unnamed>::", "", k)
    k = re.sub(r"[<(].*", "", k)[:60]
    c = agg.setdefault(k, [0, 0.0]); c[0] += 1; c[1] += v
for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1]):
    print(f"{v:9.1f} us {100 * v / tot:5.1f}%  x{c:3d}  {k}")
if "--all" in sys.argv:
    for i, (k, v, g) in enumerate(step):
        print(f"{i:3d} {v:8.1f} {g:>16} {re.sub(r'void |at::|native::|lasr::', '', k)[:70]}")
