#!/bin/bash
# Round-2 trip 7 (1 GPU): half-footprint walk A/B on top of the bulk-staged forward; config D with the staging ring's
# smaller scan chunk; LASR's real batch sizes (ico3, 8 / 16 frames) through the whole region.
OUT=gpurun_out/r2t7; mkdir -p $OUT
timeout 600 bash profiles/trip_ab.sh r2t7 nobulk bulk half
for v in nobulk bulk half; do
  LASR_B200_LIB=lasr_b200/csrc/variants/$v.so python bench.py --config D --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_D_$v.json 2> $OUT/bench_D_$v.err
  LASR_B200_LIB=lasr_b200/csrc/variants/$v.so python bench.py --config E --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 10 > $OUT/bench_E_$v.json 2> $OUT/bench_E_$v.err
  python -c "
import json
d=json.loads(open('$OUT/bench_D_$v.json').read().strip().splitlines()[-1]); print('D $v', d['kernels'], round(d['ms_per_step'],3))
d=json.loads(open('$OUT/bench_E_$v.json').read().strip().splitlines()[-1]); print('E $v', round(d['ms_per_step'],3), round(d['value'],1))"
done
for fr in 8 16 64; do
  python bench.py --config E --mesh ico3 --frames $fr --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_E_ico3_$fr.json 2> $OUT/bench_E_ico3_$fr.err
  python bench.py --config E --mesh ico3 --frames $fr --impl reference --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_Eref_ico3_$fr.json 2> $OUT/bench_Eref_ico3_$fr.err
  python -c "
import json
for t in ('E','Eref'):
    d=json.loads(open('$OUT/bench_%s_ico3_$fr.json' % t).read().strip().splitlines()[-1]); print(t, 'ico3 $fr frames', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches/step', d['gpu_launches']/d['steps'])"
done
