#!/bin/bash
# the driver's own invocations, then the other BASELINE configurations
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ego_k20.json 2> gpurun_out/final_ego_k20.err; echo "rc=$?"; tail -c 300 gpurun_out/final_ego_k20.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "rc=$?"
python bench.py --gpus 1 > gpurun_out/final_ego_default.json 2> gpurun_out/final_ego_default.err; echo "rc=$?"
bash tools/multi_gpu_bench.sh 1
python -c "
import json
for f in ('final_ego_k20','final_ego_default','final_ref'):
    b=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(b.get('value',0),1), b.get('e2e',{}).get('value'), b.get('ms_per_step'), b.get('steps'), b.get('clocks'))
"
