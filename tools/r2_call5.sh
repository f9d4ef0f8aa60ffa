#!/bin/bash
# Round 2, call 5: the new bench (148 unique frames, u16 e2e) + the cfg3 / cfg4 / cfg5 lines at N = 1.
python bench.py --steps 5 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err; cut -c1-700 gpurun_out/r2c_bench_cfg2.json; tail -3 gpurun_out/r2c_bench_cfg2.err
for w in cfg3 cfg4 cfg5; do
  python bench.py --workload $w --steps 3 --no-cpu > gpurun_out/r2c_bench_$w.json 2> gpurun_out/r2c_bench_$w.err; cut -c1-900 gpurun_out/r2c_bench_$w.json; tail -3 gpurun_out/r2c_bench_$w.err
done
