#!/bin/bash
# Round-2 trip 14 (1 GPU): final sanity of the round -- GPU tests, smoke, the two default bench lines the driver runs,
# configs E / C / D once more on the final build.
OUT=gpurun_out/r2t14; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $OUT/bench_A_ours.json 2> $OUT/bench_A_ours.err; tail -c 300 $OUT/bench_A_ours.err
python bench.py --impl reference > $OUT/bench_A_reference.json 2> $OUT/bench_A_reference.err; tail -c 300 $OUT/bench_A_reference.err
for c in E C D; do
  python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_${c}_ours.json 2> $OUT/bench_${c}_ours.err; tail -c 300 $OUT/bench_${c}_ours.err
done
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'), (d.get('extras') or {}).get('hot_path',{}).get('ms_per_step'), d.get('native_code_of_this_repo_loaded'))"; done
