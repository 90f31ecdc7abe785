#!/bin/bash
# Round-2 trip 20 (1 GPU): LBS / skin-weight backward kernels with one warp per bone, warp-wide partial reductions; same runs as trip 19
OUT=gpurun_out/r2t20; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_ours.json 2> $OUT/bench_E_ours.err; tail -c 300 $OUT/bench_E_ours.err
python bench.py --config C --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_C_ours.json 2> $OUT/bench_C_ours.err; tail -c 300 $OUT/bench_C_ours.err
for fr in 8 16 64; do
  python bench.py --config E --mesh ico3 --frames $fr --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_E_ico3_$fr.json 2> $OUT/bench_E_ico3_$fr.err
done
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --extras > $OUT/bench_A_extras.json 2> $OUT/bench_A_extras.err; tail -c 300 $OUT/bench_A_extras.err
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
x=d.get('extras',{})
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),4), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'launches/step', d.get('gpu_launches',0)/max(d.get('steps',1),1), {k:{a:round(b,4) for a,b in v.items() if isinstance(b,float)} for k,v in x.items() if k in ('image_losses','staging','lbs_project')})"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/hot8_launches.csv python bench.py --config E --mesh ico3 --frames 8 --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot8.log 2>&1
ls $OUT
python bench.py > $OUT/bench_A_default.json 2> $OUT/bench_A_default.err; tail -c 200 $OUT/bench_A_default.err; python -c "
import json; d=json.loads(open(\"$OUT/bench_A_default.json\").read().strip().splitlines()[-1]); print(\"default A\", d[\"value\"], d[\"ms_per_step\"], d[\"e2e\"][\"value\"], d.get(\"extras\",{}).get(\"hot_path\"))"
