#!/bin/bash
# Round-2 trip 23 (1 GPU, short): where the single-set backward loses time in compact mode -- one ncu capture of the
# compact and of the default kernel at config A (same process settings, cold-cache times: compare metrics, not durations).
OUT=gpurun_out/r2t23; mkdir -p $OUT
for k in compact 1; do
  LASR_B200_KEEP_WORKSPACE=$k timeout 300 ncu --set full --clock-control none --import-source on -k regex:"raster_backward_kernel" -s 8 -c 1 -o $OUT/bwd_keep_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-extras > $OUT/ncu_$k.log 2>&1
done
ls -la $OUT
