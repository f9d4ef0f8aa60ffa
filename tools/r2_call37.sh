#!/bin/bash
# Round 2, call 37: inflate with an 8 KB ring (16 streams per SM) + far matches from the flushed output
timeout 600 python -m pytest tests/test_gpu_tum.py -x -q 2>&1 | tail -3
timeout 300 python tools/tum_probe.py 592 2>&1 | tail -2
timeout 300 python tools/tum_probe.py 1184 2>&1 | tail -1
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_tum.py -q -x -k "far or shape0 or rejects" 2>&1 | grep -v "^=========     " | tail -4
