#!/bin/bash
# Round 2, call 1: baseline of the round-1 build on a fresh box, MLE size-class variant, hardware counters for pose_kernel.
python bench.py --no-cpu --steps 5 > gpurun_out/r2a_bench_base.json 2> gpurun_out/r2a_bench_base.err
cut -c1-300 gpurun_out/r2a_bench_base.json
LSL_MLE_SIZE_CLASSES=1 python bench.py --no-cpu --steps 5 > gpurun_out/r2a_bench_sc.json 2> gpurun_out/r2a_bench_sc.err
cut -c1-300 gpurun_out/r2a_bench_sc.json
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section SourceCounters --section InstructionStats \
  --clock-control none --import-source on -k regex:"pose_kernel" -s 3 -c 1 -o gpurun_out/r2a_pose \
  python bench.py --batch 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2a_ncu_pose.log 2>&1
tail -3 gpurun_out/r2a_ncu_pose.log | cut -c1-200
