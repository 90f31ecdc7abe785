"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name ...`:
top CUDA source lines by executed warp instructions and by stall samples.
Usage: python ncu_source_top.py file.csv [N]"""
import csv
import sys


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur, hdr, recs = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 10 or r[2] != "-":   # keep only the per-CUDA-line summary rows
            continue
        i_inst = hdr.index("Instructions Executed")
        i_thr = hdr.index("Thread Instructions Executed")
        i_s = hdr.index("# Samples")
        recs.append((cur, r[0], r[1].strip()[:100], num(r[i_inst]), num(r[i_s]), num(r[i_thr])))
    tot_i = sum(x[3] for x in recs) or 1
    tot_s = sum(x[4] for x in recs) or 1
    print(f"total warp-instr {tot_i:.4e}  total samples {tot_s:.0f}")
    print("--- by instructions executed")
    for f, ln, src, inst, samp, thr in sorted(recs, key=lambda x: -x[3])[:top]:
        print(f"{100*inst/tot_i:5.1f}% inst {100*samp/tot_s:5.1f}% smp thr/inst {thr/max(inst,1):4.1f} {f}:{ln} {src}")
    print("--- by stall samples")
    for f, ln, src, inst, samp, thr in sorted(recs, key=lambda x: -x[4])[:max(10, top // 2)]:
        print(f"{100*samp/tot_s:5.1f}% smp {100*inst/tot_i:5.1f}% inst {f}:{ln} {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
