#!/bin/bash
# Round 2, call 38: inflate literal run unrolled by three
timeout 600 python -m pytest tests/test_gpu_tum.py -x -q 2>&1 | tail -2
timeout 300 python tools/tum_probe.py 592 2>&1 | tail -1
timeout 300 python tools/tum_probe.py 1184 2>&1 | tail -1
