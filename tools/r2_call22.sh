#!/bin/bash
# Round 2, call 22: SIFT tests after the blur rewrite + phase times at batch 148
timeout 300 python -m pytest tests/test_gpu_sift.py -x -q 2>&1 | tail -5
LSL_SIFT_PROFILE=1 LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu 2> gpurun_out/r2k_cfg3.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg3', round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {n: round(t,2) for n, t in k.items() if t > 1})"
grep "sift phases" gpurun_out/r2k_cfg3.err | tail -3
