#!/bin/bash
# Round 2, call 39: 6 x 6 LU of line_mle_kernel with one row per lane (registers + shuffles instead of shared memory)
timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py tests/test_gpu_lsd.py -x -q 2>&1 | tail -2
LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --no-cpu --no-pipeline --batch 592 --unique 148 --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 b592', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})"
