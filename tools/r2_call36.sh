#!/bin/bash
# Round 2, call 36: inflate literal-run fast path — parity (all TUM GPU tests) and ingest time
timeout 600 python -m pytest tests/test_gpu_tum.py -x -q 2>&1 | tail -3
timeout 300 python tools/tum_probe.py 592 2>&1 | tail -2
