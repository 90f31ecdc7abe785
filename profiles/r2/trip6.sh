#!/bin/bash
# Round-2 trip 6 (1 GPU): GPU tests of the current build (packed gradients, vectorised helpers), the bulk-staged forward
# A/B (cp.async.bulk + mbarrier), packed vs planar image gradient on config E, helper-kernel extras and the ncu captures
# the round's summaries are written from.
OUT=gpurun_out/r2t6; mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
timeout 600 bash profiles/trip_ab.sh r2t6 base bulk bulk_mb3
for pk in 1 0; do
  LASR_B200_PACKED_GRAD=$pk python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_packed$pk.json 2> $OUT/bench_E_packed$pk.err
  LASR_B200_PACKED_GRAD=$pk python bench.py --config C --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_C_packed$pk.json 2> $OUT/bench_C_packed$pk.err
done
for f in $OUT/bench_[EC]_packed*.json; do python -c "
import json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])"; done
python bench.py --steps 100 --warmup 5 --extras > $OUT/bench_A_extras.json 2> $OUT/bench_A_extras.err; tail -c 300 $OUT/bench_A_extras.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward_kernel<true, false|raster_backward_kernel<1, 0" -s 6 -c 2 -o $OUT/raster python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/hot_launches.csv python bench.py --config E --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward_kernel|head_" -s 10 -c 5 -o $OUT/hotpath python bench.py --config E --no-graph --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/ncu_hot_full.log 2>&1
LASR_B200_LIB=lasr_b200/csrc/variants/bulk.so ncu --set full --clock-control none --import-source on -k regex:"raster_forward" -s 3 -c 1 -o $OUT/forward_bulk python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_bulk.log 2>&1
ls -la $OUT | head -40
