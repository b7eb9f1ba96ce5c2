python tools/interactee_probe.py 9 2>&1 | tail -1
