#!/bin/bash
# Round 2, call 27: compute-sanitizer on the kernels added / changed this round (SIFT, RANSAC rounds, MLE pivot)
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_sift.py -q -x -k "oracle" 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_memcheck_sift.log; tail -4 gpurun_out/r2_sanitizer_memcheck_sift.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_sift.py -q -x -k "oracle" 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_racecheck_sift.log; tail -4 gpurun_out/r2_sanitizer_racecheck_sift.log
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_extract.py -q -x -k "small" 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_memcheck_extract.log; tail -4 gpurun_out/r2_sanitizer_memcheck_extract.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_extract.py -q -x -k "small" 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_racecheck_extract.log; tail -4 gpurun_out/r2_sanitizer_racecheck_extract.log
