#!/bin/bash
# Round-2 trip 24 (1 GPU, short): compact mode with the record's loads issued before the vertex fetch (one memory round
# trip per face instead of two) -- GPU tests, A/B timing against the default mode, config A / E in compact mode, and the
# default line of the final build.
OUT=gpurun_out/r2t24; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for k in 1 compact; do
  LASR_B200_KEEP_WORKSPACE=$k python profiles/ab_timing.py keep_$k 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
done
LASR_B200_KEEP_WORKSPACE=compact python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 20 > $OUT/bench_A_keep_compact.json 2> $OUT/bench_A_keep_compact.err
LASR_B200_KEEP_WORKSPACE=compact python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_keep_compact.json 2> $OUT/bench_E_keep_compact.err
python bench.py --no-cpu-baseline > $OUT/bench_A_default.json 2> $OUT/bench_A_default.err
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),4), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'), (d.get('extras',{}).get('hot_path') or {}).get('ms_per_step'))"; done
