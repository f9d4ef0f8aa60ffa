#!/bin/bash
# Round 2, call 21: full GPU suite after the SIFT row, cfg3 with device SIFT, cfg2 at the new default batch,
# source-level ncu capture of line_mle_kernel + pose_kernel (batch 148) for the instruction mix.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2j_pytest_gpu.log
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2j_bench_cfg3.json 2> gpurun_out/r2j_bench_cfg3.err; cut -c1-1500 gpurun_out/r2j_bench_cfg3.json; tail -3 gpurun_out/r2j_bench_cfg3.err
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2j_bench_cfg2.json 2> gpurun_out/r2j_bench_cfg2.err; cut -c1-1200 gpurun_out/r2j_bench_cfg2.json; tail -3 gpurun_out/r2j_bench_cfg2.err
LSL_BENCH_NOCLOCKS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"line_mle_kernel|pose_kernel" -c 2 -o gpurun_out/r2j_mle_pose -f python bench.py --batch 148 --unique 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2j_ncu.log 2>&1; tail -3 gpurun_out/r2j_ncu.log
ls -la gpurun_out/r2j_mle_pose.ncu-rep
