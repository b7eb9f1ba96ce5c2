#!/bin/bash
python tools/host_profile.py 64 2>&1 | head -40 | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or async or golden" 2>&1 | tail -3
