#!/bin/bash
# A/B trip: bit-compare outputs and time kernels of several library builds (lasr_b200/csrc/variants/*.so, built on
# the CPU box with `LASR_B200_LIB=... python -m lasr_b200._build -D...`), then one ncu source-level capture.
# usage (under gpurun): bash profiles/trip_ab.sh <tag> <baseline-variant> <variant>...
TAG=$1; BASE=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
V=lasr_b200/csrc/variants
LASR_B200_LIB=$V/$BASE.so python profiles/ab_outputs.py $OUT/ab_$BASE.npz > $OUT/ab_$BASE.log 2>&1
LASR_B200_LIB=$V/$BASE.so python profiles/ab_timing.py $BASE 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
for v in "$@"; do
  LASR_B200_LIB=$V/$v.so python profiles/ab_outputs.py $OUT/ab_$v.npz > $OUT/ab_$v.log 2>&1
  python profiles/ab_outputs.py --compare $OUT/ab_$BASE.npz $OUT/ab_$v.npz > $OUT/ab_compare_$v.log 2>&1; tail -1 $OUT/ab_compare_$v.log
  LASR_B200_LIB=$V/$v.so python profiles/ab_timing.py $v 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
done
LASR_B200_LIB=$V/$BASE.so python profiles/ab_timing.py $BASE-again 2>&1 | grep "^AB" | tee -a $OUT/ab_timing.log
rm -f $OUT/ab_*.npz
