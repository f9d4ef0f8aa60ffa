#!/bin/bash
# Round 2, call 42: SIFT extrema with shared-memory DoG tiles, rootsift staged through shared memory
timeout 400 python -m pytest tests/test_gpu_sift.py tests/test_gpu_hybrid.py -x -q 2>&1 | tail -2
LSL_SIFT_PROFILE=1 LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --workload cfg3 --batch 148 --unique 148 --no-pipeline --steps 2 --warmup 3 --no-cpu 2> gpurun_out/r2r_cfg3.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg3 b148', round(d['value'],1), round(d['e2e']['value'],1), {n: round(t,2) for n, t in k.items() if t > 1})"
grep "sift phases" gpurun_out/r2r_cfg3.err | tail -1
