#!/bin/bash
# Round-2 trip 3 (1 GPU): consecutive faces per backward warp (L1 locality of the plane loads), and an ncu capture of the
# forward at config D's shape (tile-scan share at 1024 tiles x 20480 faces).
OUT=gpurun_out/r2t3; mkdir -p $OUT
bash profiles/trip_ab.sh r2t3 fpw1 fpw4 fpw8 fpw16
for v in fpw1 fpw8; do
  LASR_B200_LIB=lasr_b200/csrc/variants/$v.so python bench.py --config D --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 5 > $OUT/bench_D_$v.json 2> $OUT/bench_D_$v.err
  python -c "
import json; d=json.loads(open('$OUT/bench_D_$v.json').read().strip().splitlines()[-1]); print('D $v', d['kernels'])"
done
ncu --set full --clock-control none --import-source on -k regex:"raster_forward_kernel|raster_backward_kernel" -s 6 -c 3 -o $OUT/raster_D \
    python bench.py --config D --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_D.log 2>&1
LASR_B200_LIB=lasr_b200/csrc/variants/fpw8.so ncu --set full --clock-control none --import-source on -k regex:"raster_backward_kernel" -s 4 -c 1 -o $OUT/backward_A_fpw8 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_A.log 2>&1
ls -la $OUT
