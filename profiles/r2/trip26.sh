#!/bin/bash
# Round-2 trip 26 (1 GPU, last minutes of the budget): frame sums of the shared-parameter gradients in one launch --
# GPU tests, smoke(), the 8-frame region.
OUT=gpurun_out/r2t26; mkdir -p $OUT
timeout 150 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 60 python bench.py --config E --mesh ico3 --frames 8 --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_E_ico3_8.json 2> $OUT/bench_E_ico3_8.err
python -c "
import json; d=json.loads(open('$OUT/bench_E_ico3_8.json').read().strip().splitlines()[-1]); print('ico3 8 frames', round(d['ms_per_step'],4), round(d['value'],1), d['gpu_launches']/d['steps'])"
