#!/bin/bash
# Round-2 trip 27 (1 GPU, the last minute of the budget): the committed final build's default line and config E.
OUT=gpurun_out/r2t27; mkdir -p $OUT
timeout 70 python bench.py --no-cpu-baseline > $OUT/bench_A_default.json 2> $OUT/bench_A_default.err
timeout 40 python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_ours.json 2> $OUT/bench_E_ours.err
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),4), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'), (d.get('extras',{}).get('hot_path') or {}).get('ms_per_step'), d.get('gpu_launches'))"; done
