#!/bin/bash
# Last pass after the code-size changes: all GPU tests, bench (default flags), ncu launch list.
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r1h_pytest_gpu.log
python bench.py > gpurun_out/r1h_bench_n1.json 2> gpurun_out/r1h_bench_n1.err
cut -c1-330 gpurun_out/r1h_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1h_launches_b592.csv python bench.py --batch 592 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1h_ncu_list.log 2>&1
tail -2 gpurun_out/r1h_ncu_list.log | cut -c1-200
