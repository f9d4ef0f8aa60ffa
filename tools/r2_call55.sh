#!/bin/bash
# Round 2, call 55: Node::computeInliersAndError on the device against the oracle
timeout 400 python -m pytest tests/test_gpu_hybrid.py -x -q 2>&1 | tail -15
