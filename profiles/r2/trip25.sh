#!/bin/bash
# Round-2 trip 25 (1 GPU, last minutes of the budget): the committed final state -- GPU tests and smoke().
OUT=gpurun_out/r2t25; mkdir -p $OUT
timeout 150 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
