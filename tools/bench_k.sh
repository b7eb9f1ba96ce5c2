#!/bin/bash
# headline bench at the driver's K (20) and at a longer K; stderr kept
for k in 20 64 200; do
  timeout 600 python bench.py --steps $k --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_k$k.json 2> gpurun_out/r2_bench_k$k.err || tail -5 gpurun_out/r2_bench_k$k.err
  python -c "
import json,sys
b=json.loads(open('gpurun_out/r2_bench_k$k.json').read().strip().splitlines()[-1])
print('K=$k value', round(b['value']), 'e2e', round(b['e2e']['value']), 'ms', round(b['ms_per_step'],2), 'single', b['kernels'].get('single_batch_latency_ms'))"
done
SEEME_PIPELINE_DEPTH=8 SEEME_SAMPLER_BACKEND=graph timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_k20_d8graph.json 2>/dev/null
python -c "
import json
b=json.loads(open('gpurun_out/r2_bench_k20_d8graph.json').read().strip().splitlines()[-1])
print('K=20 depth 8 graph: value', round(b['value']), 'e2e', round(b['e2e']['value']))"
