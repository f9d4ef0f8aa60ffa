#!/bin/bash
# Round 2, call 51: pose_hybrid_kernel at 2 CTAs per SM (cfg3), relmotion_kernel at 4 / 6 CTAs per SM (cfg5)
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
PROBE_FLAGS="--workload cfg3" timeout 500 bash tools/variant_probe.sh hyb2 2>&1 | sed 's/{.*sift_kernels/ sift_kernels/' | tee gpurun_out/r2x_variants.log
PROBE_FLAGS="--workload cfg5" timeout 500 bash tools/variant_probe.sh rm4 rm6 2>&1 | sed 's/{.*line_mle_kernel/ line_mle_kernel/' | tee -a gpurun_out/r2x_variants.log
