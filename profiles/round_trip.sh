#!/bin/bash
# One GPU-box trip that produces everything profiles/ and the docs quote.  Usage (under gpurun):
#   bash profiles/round_trip.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python tests/golden/make_softras_golden.py $OUT/golden --report-only > $OUT/golden_report.log 2>&1; tail -2 $OUT/golden_report.log
python bench.py --steps 50 --warmup 5 --extras > $OUT/bench_ours.json 2> $OUT/bench_ours.err
python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
for cfg in "ico3 256 64" "ico4 256 64" "ico5 512 32"; do set -- $cfg
  python bench.py --mesh $1 --image-size $2 --frames $3 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_ours_$1_$2.json 2>> $OUT/bench_ours.err
  python bench.py --impl reference --mesh $1 --image-size $2 --frames $3 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_ref_$1_$2.json 2>> $OUT/bench_ref.err
done
# the widened region: whole hot path (fused / three renders / reference-style), colour sets, regularisers
python profiles/hotpath_timing.py 26 32 graph > $OUT/hotpath.log 2>&1
FUSED=0 python profiles/hotpath_timing.py 26 32 graph >> $OUT/hotpath.log 2>&1
python profiles/hotpath_timing.py 26 32 >> $OUT/hotpath.log 2>&1
python profiles/hotpath_timing.py 26 32 reference >> $OUT/hotpath.log 2>&1
python profiles/hotpath_timing.py 1 32 graph >> $OUT/hotpath.log 2>&1
grep hotpath $OUT/hotpath.log
python profiles/multi_set_timing.py > $OUT/multi_set.log 2>&1; cat $OUT/multi_set.log
python profiles/regularizer_timing.py 64 > $OUT/regularizers.log 2>&1; cat $OUT/regularizers.log
python profiles/staging_timing.py 16 20 2>&1 | grep staging | tee $OUT/staging.log
python profiles/aux_timing.py 2>&1 | grep chamfer | tee $OUT/chamfer.log
# profiler passes (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward" -s 6 -c 2 -o $OUT/raster python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/hot_launches.csv python profiles/hotpath_timing.py 26 32 > $OUT/ncu_hot_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward_kernel<true, false|head_|lbs_backward|skin_backward" -s 8 -c 5 -o $OUT/hotpath python profiles/hotpath_timing.py 26 32 > $OUT/ncu_hot_full.log 2>&1
for f in $OUT/bench_*.json; do echo $f; python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(' value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('kernels'))"; done
