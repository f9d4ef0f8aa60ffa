#!/bin/bash
# Round 2, call 23: MLE pivot by REDUX + pointer-bump JtJ loop (parity + time), SIFT row pass register-blocked
timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_gpu_sift.py tests/test_gpu_lsd.py -x -q 2>&1 | tail -5
LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --batch 592 --unique 148 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 b592', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 3})"
LSL_SIFT_PROFILE=1 LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu 2> gpurun_out/r2l_cfg3.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg3', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {n: round(t,2) for n, t in k.items() if t > 1})"
grep "sift phases" gpurun_out/r2l_cfg3.err | tail -2
