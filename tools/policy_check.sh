#!/bin/bash
r() { echo "$@"; env "$@" PROBE_ASYNC=1 PROBE_STEPS=96 PROBE_B=64 python tools/pipeline_probe.py $D 2>&1 | grep "async depth\|Error" | cut -c1-150; }
D=32 r PROBE_BACKEND=tile
D=32 r PROBE_BACKEND=tile PROBE_CONFIG=config_mld_gimo.yaml
D=32 r PROBE_BACKEND=auto PROBE_CONFIG=config_mld_gimo.yaml
D=8 r PROBE_BACKEND=persistent PROBE_CONFIG=config_mld_gimo.yaml
D=8 r PROBE_BACKEND=auto PROBE_CONFIG=config_mld_gimo.yaml
D=16 r PROBE_BACKEND=auto PROBE_CONFIG=config_mld_gimo.yaml
