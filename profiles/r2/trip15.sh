#!/bin/bash
# Round-2 trip 15 (1 GPU, short): cost of not keeping the forward's per-face scratch for the backward
# (LASR_B200_KEEP_WORKSPACE=0: the backward recomputes the records), and the GPU tests.
OUT=gpurun_out/r2t15; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
for k in 1 0 1; do
  LASR_B200_KEEP_WORKSPACE=$k python profiles/ab_timing.py keep$k 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
done
for k in 1 0; do
  LASR_B200_KEEP_WORKSPACE=$k python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_keep$k.json 2> $OUT/bench_E_keep$k.err
  python -c "
import json; d=json.loads(open('$OUT/bench_E_keep$k.json').read().strip().splitlines()[-1]); print('E keep$k', round(d['ms_per_step'],3), round(d['value'],1))"
done
