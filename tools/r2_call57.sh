#!/bin/bash
# Round 2, call 57: hypotheses per RANSAC round (8 / 16 / 24; 12 is the default)
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
LSL_BENCH_BATCH=592 timeout 600 bash tools/variant_probe.sh rch8 rch16 rch24 2>&1 | sed 's/{.*line3d_ransac_kernel/ line3d_ransac_kernel/' | tee gpurun_out/r2aa_variants.log
