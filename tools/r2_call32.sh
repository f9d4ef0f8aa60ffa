#!/bin/bash
# Round 2, call 32: pose_kernel with the chain-sum contributions in shared memory (parity + time + phase profile)
cp lineslam_b200/liblsl_b200.so /tmp/lib_keep.so
cp gpurun_variants/lib_posesm.so lineslam_b200/liblsl_b200.so
timeout 300 python -m pytest tests/test_gpu_pair.py tests/test_gpu_configs.py -x -q 2>&1 | tail -2
cp /tmp/lib_keep.so lineslam_b200/liblsl_b200.so
LSL_BENCH_BATCH=592 timeout 400 bash tools/variant_probe.sh posesm 2>&1 | tee gpurun_out/r2q_variants.log
for v in poseprof posesmprof; do
  cp gpurun_variants/lib_$v.so lineslam_b200/liblsl_b200.so
  LSL_BENCH_NOCLOCKS=1 timeout 200 python bench.py --no-cpu --no-pipeline --batch 592 --unique 148 --steps 1 --warmup 3 2>&1 >/dev/null | grep "pose phase" | tail -15 > gpurun_out/r2q_$v.log
  echo "== $v"; cat gpurun_out/r2q_$v.log
done
cp /tmp/lib_keep.so lineslam_b200/liblsl_b200.so
