#!/bin/bash
python tools/interactee_probe.py 8 2>&1 | tail -1 | cut -c1-400
b() { python bench.py "$@" 2>gpurun_out/policy_err.txt | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', '->', round(b['value']), 'e2e', round(b['e2e']['value']), 'ms', round(b['ms_per_step'],2))" || tail -5 gpurun_out/policy_err.txt; }
b --config interactee --replications 10
b --steps 20 --warmup 5 --no-extras --no-cpu-baseline
b --steps 64 --warmup 5 --no-extras --no-cpu-baseline
b --config gimo --steps 32 --warmup 3
