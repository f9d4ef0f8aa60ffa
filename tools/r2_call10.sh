#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_hybrid.py -m gpu -x -q 2>&1 | tail -8
