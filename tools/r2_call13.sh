#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_graph.py tests/test_gpu_threads.py -m gpu -x -q 2>&1 | tail -4
for w in cfg3 cfg5; do
  python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r2h_bench_$w.json 2> gpurun_out/r2h_bench_$w.err; tail -2 gpurun_out/r2h_bench_$w.err
  python -c "
import sys, json
d = json.loads(open('gpurun_out/r2h_bench_$w.json').readline()); k = d['kernel_ms_per_step']
print('$w', round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), {n: round(t,2) for n, t in k.items() if t > 1})
"
done
