#!/bin/bash
# Round 2, call 29: pair stage on the pair stream (begin / end) — parity test, then the bench with and without the pipeline
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_pair.py tests/test_gpu_threads.py -x -q 2>&1 | tail -4
for flag in "" "--no-pipeline"; do
LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py $flag --steps 4 --warmup 3 --no-cpu 2>gpurun_out/r2o_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 b1184 [$flag]', round(d['value'],1), round(d['e2e']['value'],1), d['host_ms_each_step'], {n: round(t,2) for n, t in k.items() if t > 6})"
tail -2 gpurun_out/r2o_bench.err
done
