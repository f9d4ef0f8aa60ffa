#!/bin/bash
python tools/tc_probe.py 256
LSL_MATCH_TC=0 python tools/tc_probe.py 256
for tc in 1 0; do LSL_MATCH_TC=$tc python bench.py --workload cfg3 --steps 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); k = d['kernel_ms_per_step']
print('cfg3 TC=$tc', round(d['value'],1), round(d['ms_per_step'],2), {n: round(t,3) for n, t in k.items() if 'match' in n or 'pose' in n})
"; done
SEC="--section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section InstructionStats"
timeout 300 ncu $SEC --clock-control none -k regex:"match_points_tc_kernel|match_points_refine_kernel" -s 4 -c 2 -o gpurun_out/r2f_tc python tools/tc_probe.py 256 > gpurun_out/r2f_ncu.log 2>&1; tail -1 gpurun_out/r2f_ncu.log | cut -c1-200
