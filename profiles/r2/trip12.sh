#!/bin/bash
# Round-2 trip 12 (1 GPU): per-pixel packing pre-pass for the single-set backward, A/B + tests + bench.
OUT=gpurun_out/r2t12; mkdir -p $OUT
timeout 600 bash profiles/trip_ab.sh r2t12 nopack pack
cat $OUT/ab_compare_pack.log | tail -12
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
for v in nopack pack; do
  LASR_B200_LIB=lasr_b200/csrc/variants/$v.so python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras > $OUT/bench_A_$v.json 2> $OUT/bench_A_$v.err
  LASR_B200_LIB=lasr_b200/csrc/variants/$v.so python bench.py --config D --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_D_$v.json 2> $OUT/bench_D_$v.err
  python -c "
import json
for c in 'AD':
    d=json.loads(open('$OUT/bench_%s_$v.json' % c).read().strip().splitlines()[-1]); print(c, '$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['kernels'])"
done
