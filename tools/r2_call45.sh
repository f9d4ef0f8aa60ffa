#!/bin/bash
# Round 2, call 45: ncu full of the mid-size kernels (ll_angle, NFA, MSLD, seed list, ypass) at 592 frames
LSL_BENCH_NOCLOCKS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"ll_angle_kernel|lsd_nfa_kernel|line_msld_kernel|ypass_tma_kernel|xpass_kernel" -c 5 -o gpurun_out/r2_mid5 -f python bench.py --no-pipeline --batch 592 --unique 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_mid5_ncu.log 2>&1
tail -2 gpurun_out/r2_mid5_ncu.log | cut -c1-200
