#!/bin/bash
# Round-2 trip 5 (8 GPUs): frame-sharded scaling of configs A, D and E at 1 / 2 / 4 / 8 GPUs (ours) and 1 / 8 (reference),
# launched as the driver launches bench.py.
OUT=gpurun_out/r2t5; mkdir -p $OUT
run() { # config impl gpus steps
  if [ "$3" = "1" ]; then
    python bench.py --config $1 --impl $2 --gpus 1 --steps $4 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-extras > $OUT/bench_$1_$2_$3gpu.json 2> $OUT/bench_$1_$2_$3gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port 29511 bench.py \
      --config $1 --impl $2 --gpus $3 --steps $4 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-extras > $OUT/bench_$1_$2_$3gpu.json 2> $OUT/bench_$1_$2_$3gpu.err
  fi
  python -c "
import json
t=open('$OUT/bench_$1_$2_$3gpu.json').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$1 $2 $3gpu value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'e2e_ms', round(d.get('e2e',{}).get('ms_per_step',0),3))"
}
for n in 8 4 2 1; do run A ours $n 100; done
for n in 8 1; do run A reference $n 5; done
for n in 8 4 2 1; do run D ours $n 30; done
for n in 8 1; do run D reference $n 2; done
for n in 8 4 2 1; do run E ours $n 30; done
for n in 8 1; do run E reference $n 2; done
nvidia-smi topo -m > $OUT/topo.txt 2>&1
