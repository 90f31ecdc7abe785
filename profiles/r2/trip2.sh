#!/bin/bash
# Round-2 trip 2 (1 GPU): tight pixel windows A/B + plane-load bound experiment, the GPU test suite on the new default
# build, and the first lines of the rewritten bench.py (configs A, D, E, C; both arms).
OUT=gpurun_out/r2t2; mkdir -p $OUT
bash profiles/trip_ab.sh r2t2 loose tight planesl1
python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python bench.py --steps 100 --warmup 5 > $OUT/bench_A_ours.json 2> $OUT/bench_A_ours.err; tail -c 600 $OUT/bench_A_ours.err
python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_A_ref.json 2> $OUT/bench_A_ref.err; tail -c 600 $OUT/bench_A_ref.err
for c in D E C; do
  python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_${c}_ours.json 2> $OUT/bench_${c}_ours.err; tail -c 600 $OUT/bench_${c}_ours.err
  python bench.py --config $c --impl reference --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_${c}_ref.json 2> $OUT/bench_${c}_ref.err; tail -c 600 $OUT/bench_${c}_ref.err
done
for f in $OUT/bench_*.json; do echo $f; python -c "
import sys,json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print(' value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'), (d.get('extras') or {}).get('hot_path'))"; done
