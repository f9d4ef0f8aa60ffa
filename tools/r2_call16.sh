#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_tum.py -m gpu -x -q 2>&1 | tail -6
python tools/tum_probe.py 592 2>&1 | tail -3
