#!/bin/bash
# Round 2, call 47: adopted ll_angle 32 x 8 tiles, MSLD at 5 and NFA at 6 CTAs per SM — parity, then probes of 32 x 4 tiles / MSLD at 6
timeout 900 python -m pytest tests/test_gpu_lsd.py tests/test_gpu_extract.py tests/test_gpu_fullsize.py tests/test_gpu_configs.py -x -q 2>&1 | tail -2
sed -i 's/if t > 3/if t > 2/' tools/variant_probe.sh
LSL_BENCH_BATCH=592 timeout 600 bash tools/variant_probe.sh lla4 msld6 2>&1 | tee gpurun_out/r2u_variants.log
