#!/bin/bash
# Round 2, call 6: hardware counters of the four dominant kernels at batch 592 (sections, no --set full: that fails on pose_kernel).
SEC="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section SourceCounters --section InstructionStats"
timeout 900 ncu $SEC --clock-control none --import-source on -k regex:"pose_kernel|line_mle_kernel|line3d_ransac_kernel|lsd_region_kernel" -s 12 -c 4 -o gpurun_out/r2d_top4 \
  python bench.py --steps 1 --warmup 3 --no-cpu --unique 48 > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log | cut -c1-300
