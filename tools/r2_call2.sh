#!/bin/bash
# Round 2, call 2: clock64 phase profile of pose_kernel (-DPOSE_PROFILE variant), batch 592.
cp lineslam_b200/liblsl_b200.so /tmp/lib_base.so
cp gpurun_variants/lib_poseprof.so lineslam_b200/liblsl_b200.so
LSL_BENCH_NOCLOCKS=1 python bench.py --no-cpu --steps 1 --warmup 3 > gpurun_out/r2b_poseprof.json 2> gpurun_out/r2b_poseprof.err
tail -16 gpurun_out/r2b_poseprof.err
cp /tmp/lib_base.so lineslam_b200/liblsl_b200.so
