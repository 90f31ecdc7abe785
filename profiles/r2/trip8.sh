#!/bin/bash
# Round-2 trip 8 (1 GPU): the round's final build -- GPU tests, smoke, bulk-store setup A/B, every configuration on both
# arms, and the ncu captures profiles/round2_*.md and traffic_r2.json are written from.
OUT=gpurun_out/r2t8; mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 bash profiles/trip_ab.sh r2t8 nostage final
python bench.py > $OUT/bench_A_ours.json 2> $OUT/bench_A_ours.err; tail -c 300 $OUT/bench_A_ours.err
python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_A_reference.json 2> $OUT/bench_A_reference.err; tail -c 300 $OUT/bench_A_reference.err
for c in B C D E; do
  python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_${c}_ours.json 2> $OUT/bench_${c}_ours.err; tail -c 300 $OUT/bench_${c}_ours.err
  python bench.py --config $c --impl reference --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_${c}_reference.json 2> $OUT/bench_${c}_reference.err
done
python bench.py --mesh ico3 --steps 50 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_ico3_ours.json 2> $OUT/bench_ico3_ours.err
python bench.py --mesh ico3 --impl reference --steps 5 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_ico3_reference.json 2> $OUT/bench_ico3_reference.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --extras > $OUT/bench_A_extras.json 2> $OUT/bench_A_extras.err
for f in $OUT/bench_*.json; do python -c "
import json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$f'.split('/')[-1], 'value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), d.get('kernels'))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward_kernel<true, false|raster_backward_kernel<1, 0|raster_setup" -s 9 -c 3 -o $OUT/raster python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/hot_launches.csv python bench.py --config E --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward_kernel|head_" -s 10 -c 5 -o $OUT/hotpath python bench.py --config E --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot_full.log 2>&1
ncu --set full --clock-control none -k regex:"losses_forward|losses_backward|stage_stats|stage_resample" -c 8 -o $OUT/helpers python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras --extras > $OUT/ncu_helpers.log 2>&1
ls $OUT | head -60
