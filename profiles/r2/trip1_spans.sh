#!/bin/bash
# First A/B of the next round: the span sweep of the backward.
#   on the CPU box:   bash profiles/next_trip_spans.sh build
#   under gpurun:     bash profiles/next_trip_spans.sh run
# `build` compiles the baseline and -DLASR_BWD_SPANS=1 into lasr_b200/csrc/variants/ (they travel with the snapshot);
# `run` bit-compares 87 output arrays of the two builds, times them (config A, ico3, K = 1 / 3), runs the rasterizer and
# hot-path GPU tests on the spans build and takes one ncu capture of its backward.
set -e
V=lasr_b200/csrc/variants
case "$1" in
build)
  mkdir -p $V
  LASR_B200_LIB=$V/base.so python -m lasr_b200._build &
  LASR_B200_LIB=$V/spans.so python -m lasr_b200._build -DLASR_BWD_SPANS=1 &
  wait
  # the backward's 21 denormal-guard instructions per sweep (profiles/round1g_instruction_budget.md): bare .ftz MUFUs were
  # slower on the window sweep (r1f t3); the span sweep changes the loop's register pressure, so measure them together
  LASR_B200_LIB=$V/spans_ftz.so python -m lasr_b200._build -DLASR_BWD_SPANS=1 -DLASR_FTZ_MUFU_BWD=1 &
  LASR_B200_LIB=$V/ftz.so python -m lasr_b200._build -DLASR_FTZ_MUFU_BWD=1 &
  wait
  # the selected edge's constants from per-lane loads of the record instead of three divergent arms of register moves
  # (bit-identical on the host harness; forward visit path -18 static instructions, no stack): both kernels, alone and
  # on top of the span sweep
  LASR_B200_LIB=$V/edgeload.so python -m lasr_b200._build -DLASR_EDGE_LOAD=1 &
  LASR_B200_LIB=$V/spans_edgeload.so python -m lasr_b200._build -DLASR_BWD_SPANS=1 -DLASR_EDGE_LOAD=1 &
  wait
  # three-set forward (the hot path's): 2 resident CTAs at 118 registers and no stack frame instead of 3 at 80 with 56 B spilled
  LASR_B200_LIB=$V/fwdm2.so python -m lasr_b200._build -DLASR_FWD_MIN_BLOCKS_MULTI=2 &
  wait
  for v in base spans spans_ftz ftz edgeload spans_edgeload fwdm2; do rm -rf lasr_b200/csrc/build/$v.so.obj; done     # object caches need not travel
  ls -la $V ;;
run)
  bash profiles/trip_ab.sh spans base spans edgeload spans_edgeload spans_ftz ftz fwdm2     # (the ftz builds differ from base by <= 3e-5 in the gradients)
  LASR_B200_LIB=$V/spans.so python -m pytest tests/test_raster_gpu.py tests/test_hotpath_gpu.py -q -x 2>&1 | tail -3
  LASR_B200_LIB=$V/spans.so ncu --set full --clock-control none --import-source on -k regex:"raster_backward_kernel" -s 4 -c 1 \
      -o gpurun_out/spans/backward python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/spans/ncu.log 2>&1 || true ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
