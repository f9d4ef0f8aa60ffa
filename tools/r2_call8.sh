#!/bin/bash
SEC="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section SourceCounters --section InstructionStats"
timeout 900 ncu $SEC --clock-control none --import-source on -k regex:"pose_kernel" -s 3 -c 1 -o gpurun_out/r2e_pose \
  python bench.py --steps 1 --warmup 3 --no-cpu --unique 48 > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log | cut -c1-200
