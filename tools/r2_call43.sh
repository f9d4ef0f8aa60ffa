#!/bin/bash
# Round 2, call 43: SIFT chunk size (L2 residency of the row-pass output vs launch count)
PROBE_FLAGS="--workload cfg3 --batch 148" timeout 600 bash tools/variant_probe.sh sift16 sift64 sift148 2>&1 | sed 's/{.*sift_kernels/ sift_kernels/' | tee gpurun_out/r2s_variants.log
