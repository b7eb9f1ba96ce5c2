#!/bin/bash
# usage: tools/multi_gpu_bench.sh N  -- runs the BASELINE configs[1..4] benches on N GPUs of this box, one JSON line each into gpurun_out/
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N "$@"; }
run --config gimo --steps 32 --warmup 3 > gpurun_out/r2_gimo_n$N.json 2> gpurun_out/r2_gimo_n$N.err
run --config interactee --replications 10 > gpurun_out/r2_interactee_n$N.json 2> gpurun_out/r2_interactee_n$N.err
run --config smpl-sweep --steps 5 > gpurun_out/r2_smpl_sweep_n$N.json 2> gpurun_out/r2_smpl_sweep_n$N.err
if [ "$2" == "ego" ]; then run --steps 64 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_egobody_n$N.json 2> gpurun_out/r2_egobody_n$N.err; fi
tail -n 2 gpurun_out/r2_*_n$N.err | tail -20
for f in gpurun_out/r2_*_n$N.json; do python - <<PY
import json
try:
    d = json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", d["value"], d["ms_per_step"], d["scaling"], d["n_gpus"])
except Exception as e:
    print("$f", "NO JSON", e)
PY
done
