#!/bin/bash
# Round-2 trip 22 (1 GPU): compact mode as separate kernel instantiations (trip 21 folded it into the default kernels behind a
# run-time test and the single-set backward lost 13 % in the DEFAULT mode); same runs as trip 21.
OUT=gpurun_out/r2t22; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
for k in 1 compact 1 compact; do
  LASR_B200_KEEP_WORKSPACE=$k python profiles/ab_timing.py keep_$k 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
done
for k in 1 compact; do
  LASR_B200_KEEP_WORKSPACE=$k python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 20 > $OUT/bench_A_keep_$k.json 2> $OUT/bench_A_keep_$k.err
  LASR_B200_KEEP_WORKSPACE=$k python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_keep_$k.json 2> $OUT/bench_E_keep_$k.err
  LASR_B200_KEEP_WORKSPACE=$k python bench.py --config D --steps 30 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 5 > $OUT/bench_D_keep_$k.json 2> $OUT/bench_D_keep_$k.err
done
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),4), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'))"; done
