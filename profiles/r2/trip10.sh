#!/bin/bash
# Round-2 trip 10 (1 GPU): GPU tests incl. the new edge cases and the reference-arm self-check; helper kernels timed as CUDA
# graphs (bench.py --extras) and under ncu one kernel at a time.
OUT=gpurun_out/r2t10; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -5 $OUT/pytest_gpu.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras --extras > $OUT/bench_A_extras.json 2> $OUT/bench_A_extras.err; tail -c 300 $OUT/bench_A_extras.err
python -c "
import json; d=json.loads(open('$OUT/bench_A_extras.json').read().strip().splitlines()[-1])
for k,v in d['extras'].items(): print(k, json.dumps(v)[:420])"
for k in losses_backward stage_stats stage_resample head_backward; do
  ncu --set full --clock-control none -k regex:"$k" -s 2 -c 1 -o $OUT/helper_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --extras > $OUT/ncu_$k.log 2>&1
done
ls $OUT
