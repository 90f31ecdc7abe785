#!/bin/bash
# Round-2 trip 16 (1 GPU): bone composition inside the LBS kernels -- GPU tests, region benches (E, C, ico3 small batches).
OUT=gpurun_out/r2t16; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
for c in E C; do
  python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_${c}_ours.json 2> $OUT/bench_${c}_ours.err; tail -c 300 $OUT/bench_${c}_ours.err
done
for fr in 8 16; do
  python bench.py --config E --mesh ico3 --frames $fr --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_E_ico3_$fr.json 2> $OUT/bench_E_ico3_$fr.err
done
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'launches/step', d.get('gpu_launches',0)/max(d.get('steps',1),1))"; done
