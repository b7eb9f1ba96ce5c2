#!/bin/bash
# usage: tools/final_n24.sh N : egobody (weak) and gimo (strong) on N GPUs
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N "$@"; }
run --config gimo --steps 96 --warmup 3 > gpurun_out/r2_gimo_n$N.json 2> gpurun_out/r2_gimo_n$N.err
if [ "$N" != "8" ]; then run --steps 64 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_egobody_n$N.json 2> gpurun_out/r2_egobody_n$N.err; fi
if [ "$N" == "8" ]; then run --config interactee --replications 10 > gpurun_out/r2_interactee_n$N.json 2> gpurun_out/r2_interactee_n$N.err; fi
for f in gpurun_out/r2_gimo_n$N.json gpurun_out/r2_egobody_n$N.json gpurun_out/r2_interactee_n$N.json; do python - <<PY
import json
try:
    d = json.loads(open("$f").read().strip().splitlines()[-1]); print("$f", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["ms_per_step"], d["scaling"], d["n_gpus"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "NO JSON", e)
PY
done
