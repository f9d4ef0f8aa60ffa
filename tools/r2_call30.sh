#!/bin/bash
# Round 2, call 30: extraction stream at higher priority than the pair stream
LSL_BENCH_NOCLOCKS=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu 2>gpurun_out/r2p_bench.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg2 b1184 prio', round(d['value'],1), round(d['e2e']['value'],1), d['host_ms_each_step'], {n: round(t,2) for n, t in k.items() if t > 6})"
tail -2 gpurun_out/r2p_bench.err
