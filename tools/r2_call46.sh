#!/bin/bash
# Round 2, call 46: occupancy variants of line_msld_kernel / lsd_nfa_kernel, tile height of ll_angle_kernel (592 frames, no pipeline)
LSL_BENCH_BATCH=592 timeout 900 bash tools/variant_probe.sh msld4 msld5 nfa5 nfa6 lla8 lla16 2>&1 | tee gpurun_out/r2t_variants.log
