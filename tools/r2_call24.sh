#!/bin/bash
# Round 2, call 24: line3d_ransac_kernel with draws one round ahead + pulled hypotheses; pose_kernel Hpp/bp in shared memory
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_pair.py tests/test_gpu_configs.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --batch 592 --unique 148 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 b592', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})"
