#!/bin/bash
# Round 2, call 50: 3D-line RANSAC at 8 CTAs per SM (one wave at 1184 frames instead of two)
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
timeout 600 bash tools/variant_probe.sh ransac8 2>&1 | tee gpurun_out/r2w_variants.log
