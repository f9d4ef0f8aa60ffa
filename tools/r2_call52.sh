#!/bin/bash
# Round 2, call 52: lsd_nfa_kernel with 1 / 2 warps per CTA (a CTA holds its warp slots until its slowest rectangle is done)
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
LSL_BENCH_BATCH=592 timeout 600 bash tools/variant_probe.sh nfaw1 nfaw2 2>&1 | tee gpurun_out/r2y_variants.log
