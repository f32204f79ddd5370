#!/bin/bash
# N = 1,2,4,8 back to back (what the driver does for SCALE_rNN.json); writes gpurun_out/scale_*.json
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu --no-extra > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 30 --warmup 3 --no-extra > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
  fi
  python -c "
import json; d=json.load(open('gpurun_out/scale_$N.json')); print($N, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks'])"
done
