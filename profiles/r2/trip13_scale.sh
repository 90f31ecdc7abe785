#!/bin/bash
# Round-2 trip 13 (8 GPUs, short): the final build at 1 and 8 GPUs -- config E with two flow channels over PCIe (67 MB per
# step and rank instead of 84), config D with the packing pre-pass, config A.
OUT=gpurun_out/r2t13; mkdir -p $OUT
run() { # config gpus steps
  if [ "$2" = "1" ]; then
    python bench.py --config $1 --gpus 1 --steps $3 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-extras > $OUT/bench_$1_ours_$2gpu.json 2> $OUT/bench_$1_ours_$2gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py \
      --config $1 --gpus $2 --steps $3 --warmup 3 --e2e-steps 20 --no-cpu-baseline --no-extras > $OUT/bench_$1_ours_$2gpu.json 2> $OUT/bench_$1_ours_$2gpu.err
  fi
  python -c "
import json
t=open('$OUT/bench_$1_ours_$2gpu.json').read().strip().splitlines()
d=json.loads(t[-1]) if t else {}
print('$1 ours $2gpu value', round(d.get('value',0),1), 'ms', round(d.get('ms_per_step',0),3), 'e2e', round(d.get('e2e',{}).get('value',0),1), 'e2e_ms', round(d.get('e2e',{}).get('ms_per_step',0),3), 'h2d', d.get('e2e',{}).get('h2d_bytes_per_step'))"
}
for n in 8 1; do run E $n 30; done
for n in 8 1; do run A $n 100; done
for n in 8 1; do run D $n 30; done
