#!/bin/bash
# Round 2, call 48: MSLD at 8, NFA at 8, 3D-line RANSAC at 5 / 6 CTAs per SM
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
LSL_BENCH_BATCH=592 timeout 900 bash tools/variant_probe.sh msld8 nfa8 ransac5 ransac6 2>&1 | tee gpurun_out/r2v_variants.log
