#!/bin/bash
# Pipelined throughput per sampler back-end, pipeline depth, scene-encoder grid and number of hardware work queues
# (CUDA_DEVICE_MAX_CONNECTIONS)
out=gpurun_out/r2_pipe_backend_sweep5.txt
: > $out
run() { echo "backend=$1 depth=$2 pf_grid=$3 connections=$4" >> $out; CUDA_DEVICE_MAX_CONNECTIONS=$4 PROBE_ASYNC=1 PROBE_STEPS=64 PROBE_BACKEND=$1 SEEME_PF_GRID=$3 timeout 300 python tools/pipeline_probe.py $2 2>&1 | grep "async depth" >> $out; }
c=32
runb() { echo "B=$1 backend=$2 depth=$3 enc_handles=$4" >> $out; PROBE_B=$1 PROBE_ENC_HANDLES=$4 PROBE_ASYNC=1 PROBE_STEPS=${5:-64} PROBE_BACKEND=$2 timeout 300 python tools/pipeline_probe.py $3 2>&1 | grep "async depth\|Error" >> $out; }
runb 256 tile 32 4
runb 256 tile 32 0
runb 256 tile 32 2
runb 256 tile 48 4 96
runb 256 auto 32 4
runb 128 tile 32 4 128
runb 128 tile 64 4 128
runb 128 persistent 32 4 128
runb 128 persistent 16 4 128
runb 64 tile 32 4 256
runb 64 tile 64 4 256
runb 64 persistent 32 4 256
runb 64 persistent 16 4 256
cat $out
