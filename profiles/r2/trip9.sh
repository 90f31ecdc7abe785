#!/bin/bash
# Round-2 trip 9 (1 GPU, short): ncu capture of the single-set backward of the final build (trip 8's -k regex missed it),
# GPU tests and the default bench line after the identity-camera shortcut.
OUT=gpurun_out/r2t9; mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
python bench.py > $OUT/bench_A_ours.json 2> $OUT/bench_A_ours.err; tail -c 300 $OUT/bench_A_ours.err
python -c "
import json; d=json.loads(open('$OUT/bench_A_ours.json').read().strip().splitlines()[-1]); print('A', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['kernels'])"
ncu --set full --clock-control none --import-source on -k regex:"raster_forward|raster_backward" -s 9 -c 3 -o $OUT/raster python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_full.log 2>&1
ls -la $OUT
