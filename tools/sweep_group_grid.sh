# throughput sweep over (sampler_group, pipeline_depth[, scene-encoder grid]); args: "g:d[:grid] ..."
for cfg in "$@"; do
IFS=: read g d pg <<< "$cfg"; pg=${pg:-148}
echo -n "group=$g depth=$d grid=$pg: "
SEEME_SAMPLER_GROUP=$g SEEME_PIPELINE_DEPTH=$d SEEME_PF_GRID=$pg timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 48 2>gpurun_out/sweep.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(round(d['value']), round(d['ms_per_step'],2), round(d['kernels']['single_batch_latency_ms'],1), d['gpu_launches'])
"
done
