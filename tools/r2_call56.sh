#!/bin/bash
# Round 2, call 56: SIFT blurs in frame groups whose row-pass output stays in L2 (40 MB default; 20 / 80 MB; off)
timeout 300 python -m pytest tests/test_gpu_sift.py -x -q 2>&1 | tail -1
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
PROBE_FLAGS="--workload cfg3 --batch 148" timeout 600 bash tools/variant_probe.sh siftl2_20 siftl2_80 siftl2_off 2>&1 | sed 's/{.*sift_kernels/ sift_kernels/' | tee gpurun_out/r2z_variants.log
