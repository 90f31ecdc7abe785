"""Where a kernel's warp instructions go: per-SASS-instruction executed counts of an `ncu --set full --import-source on`
capture, grouped by how often each instruction ran and by opcode.  This is what profiles/round1g_instruction_budget.md
was written from.

    python profiles/instruction_budget.py gpurun_out/r1g/raster.ncu-rep raster_backward [--per 374784] [--hot 2.0]

`--per N` divides the counts by N (warps of the launch for a warp-per-face kernel, frames for the forward) so the
numbers read "per face" / "per frame"; `--hot X` lists every instruction executed at least X times per unit with its
average active lanes and stall samples.  Needs `ncu` on PATH (reads the report, does not profile).
"""
import argparse
import collections
import csv
import io
import re
import subprocess


def load(report, kernel_regex):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
    h = rows[hdr]
    ia, ie, it, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    ins = [(r[ia].strip(), int(r[ie]), int(r[it]), int(r[isamp])) for r in rows[hdr + 1:] if len(r) > ie and r[ie].isdigit()]
    # several profiled launches of the same kernel are concatenated: keep the first
    first = ins[0][0]
    for k in range(1, len(ins)):
        if ins[k][0] == first and k > 16 and ins[k + 1][0] == ins[1][0]:
            ins = ins[:k]
            break
    return ins


def opcode(sass):
    return re.sub(r"^@!?U?P\w+\s+", "", sass).split()[0].split(".")[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="regex on the kernel name, e.g. raster_backward")
    ap.add_argument("--per", type=float, default=1.0)
    ap.add_argument("--hot", type=float, default=None)
    a = ap.parse_args()
    ins = load(a.report, a.kernel)
    tot = sum(e for _, e, _, _ in ins)
    thr = sum(t for _, _, t, _ in ins)
    print(f"{len(ins)} SASS instructions, {tot / a.per:.1f} executed per unit, {thr / tot:.2f} active lanes per instruction")
    levels = collections.OrderedDict()
    for s, e, t, _ in ins:
        if e == 0:
            continue
        key = round(e / a.per, 2) if e / a.per < 100 else int(e / a.per)
        lv = levels.setdefault(key, [0, 0, 0])
        lv[0] += 1; lv[1] += e; lv[2] += t
    print("\nexecution level (times per unit) -> static instructions, share of executed, lanes   [levels >= 1 % of the total]")
    for key, (n, e, t) in sorted(levels.items(), key=lambda kv: -kv[1][1]):
        if e / tot >= 0.01:
            print(f"  {key:>12} x {n:4d} instr  {100 * e / tot:5.1f} %  {t / e:5.1f} lanes")
    mix = collections.Counter()
    for s, e, _, _ in ins:
        mix[opcode(s)] += e
    print("\nopcode mix: " + "  ".join(f"{op} {100 * e / tot:.1f}" for op, e in mix.most_common(24)))
    if a.hot is not None:
        print(f"\ninstructions executed >= {a.hot} times per unit: index, times, lanes, stall samples, SASS")
        for i, (s, e, t, sm) in enumerate(ins):
            if e / a.per >= a.hot:
                print(f"  {i:5d} {e / a.per:9.2f} {t / e:5.1f} {sm:6d}  {s}")


if __name__ == "__main__":
    main()
