#!/bin/bash
# Round 2, call 25: line_mle_kernel with the covariance factors in an L2-resident scratch (11 KB of shared memory per warp) at 16 / 12 / 14 / 18 warps per SM
timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
LSL_BENCH_BATCH=592 timeout 600 bash tools/variant_probe.sh mle12 mle14 mle18 2>&1 | tee gpurun_out/r2m_variants.log
