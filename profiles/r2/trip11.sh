#!/bin/bash
# Round-2 trip 11 (1 GPU, short): GPU tests after the oracle fix, smoke() with the tiny fused region.
OUT=gpurun_out/r2t11; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
