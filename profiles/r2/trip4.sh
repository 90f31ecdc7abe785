#!/bin/bash
# Round-2 trip 4 (2 GPUs): the new GPU tests (full-size region parity, headline launch vs oracle, background argument),
# the parity report, and the rewritten bench.py under torch.distributed.run at N = 2 (both arms, configs A and E).
OUT=gpurun_out/r2t4; mkdir -p $OUT
python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
python profiles/r2/parity_report.py > $OUT/parity_report.jsonl 2> $OUT/parity_report.err; cat $OUT/parity_report.jsonl
run() { # config impl gpus steps
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port 29511 bench.py \
    --config $1 --impl $2 --gpus $3 --steps $4 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-extras > $OUT/bench_$1_$2_$3gpu.json 2> $OUT/bench_$1_$2_$3gpu.err
  tail -c 300 $OUT/bench_$1_$2_$3gpu.err
}
python bench.py --steps 100 --warmup 5 > $OUT/bench_A_ours_1gpu.json 2> $OUT/bench_A_ours_1gpu.err
run A ours 2 100
run A reference 2 5
run E ours 2 30
run E reference 2 3
for f in $OUT/bench_*.json; do echo $f; python -c "
import sys,json
t=open('$f').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print(' value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'launches', d.get('gpu_launches'), d.get('kernels'))"; done
