#!/bin/bash
# Round 2, call 41: ncu launch list of the SIFT kernels (cfg3, 64 frames, no pipeline)
LSL_BENCH_NOCLOCKS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -k regex:"sift_" --csv --log-file gpurun_out/r2_sift_launches.csv python bench.py --workload cfg3 --batch 64 --unique 64 --no-pipeline --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_sift_ncu.log 2>&1
tail -2 gpurun_out/r2_sift_ncu.log | cut -c1-200; wc -l gpurun_out/r2_sift_launches.csv
