#!/bin/bash
# Round-2 trip 18 (1 GPU): loss scalars riding the partial-sum launch, resample pass with hoisted per-frame constants --
# GPU tests, helper timings (bench.py --extras), region benches, launch list of the 8-frame region, ncu of the staging kernels.
OUT=gpurun_out/r2t18; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --extras > $OUT/bench_A_extras.json 2> $OUT/bench_A_extras.err; tail -c 300 $OUT/bench_A_extras.err
python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_ours.json 2> $OUT/bench_E_ours.err; tail -c 300 $OUT/bench_E_ours.err
for fr in 8 16; do
  python bench.py --config E --mesh ico3 --frames $fr --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_E_ico3_$fr.json 2> $OUT/bench_E_ico3_$fr.err
done
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'launches/step', d.get('gpu_launches',0)/max(d.get('steps',1),1), json.dumps(d.get('extras',{}).get('helpers', d.get('extras',{})))[:900])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/hot8_launches.csv python bench.py --config E --mesh ico3 --frames 8 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot8.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"stage_resample|stage_stats|losses_reduce_finalize" -c 4 -o $OUT/helpers python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras --extras > $OUT/ncu_helpers.log 2>&1
ls $OUT
