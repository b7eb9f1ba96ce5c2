#!/bin/bash
# Which stage bounds the pipelined throughput (depth 32, tile sampler): stage ablations
out=gpurun_out/r2_pipe_ablation.txt
: > $out
run() { echo "skip=$1 cache_scene=$2" >> $out; PROBE_SKIP=$1 PROBE_CACHE_SCENE=$2 PROBE_ASYNC=1 PROBE_STEPS=96 timeout 300 python tools/pipeline_probe.py 32 2>&1 | grep "async depth\|Error" | cut -c1-140 >> $out; }
run none ""
run sampler ""
run vaedec ""
run vaeenc ""
run smpl ""
run sampler,vaedec,vaeenc,smpl ""
run none 1
run sampler 1
run sampler,vaedec,vaeenc,smpl 1
cat $out
