#!/bin/bash
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N "$@"; }
run --config gimo --steps 96 --warmup 3 > gpurun_out/r2_gimo_n$N.json 2> gpurun_out/r2_gimo_n$N.err
python -c "
import json
d = json.loads(open('gpurun_out/r2_gimo_n$N.json').read().strip().splitlines()[-1]); print('gimo N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), d['ms_per_step'], d['clocks']['sm_mhz'])"
