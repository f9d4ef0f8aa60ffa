#!/bin/bash
# Round 2, profile pass: all GPU tests, bench lines of every config, reference arm, ncu launch list, ncu full of the top kernels, TUM ingest at 1184 frames
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2_final_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2_final_bench_cfg2_n1.json 2> gpurun_out/r2_final_bench_cfg2_n1.err; cut -c1-260 gpurun_out/r2_final_bench_cfg2_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference_arm.json 2>/dev/null; cut -c1-260 gpurun_out/r2_final_bench_reference_arm.json
for wl in cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $wl --no-cpu > gpurun_out/r2_final_bench_${wl}_n1.json 2> gpurun_out/r2_final_bench_${wl}_n1.err; cut -c1-200 gpurun_out/r2_final_bench_${wl}_n1.json
done
LSL_BENCH_NOCLOCKS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches_b1184.csv python bench.py --no-pipeline --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_final_ncu_list.log 2>&1
tail -1 gpurun_out/r2_final_ncu_list.log | cut -c1-200
LSL_BENCH_NOCLOCKS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"line_mle_kernel|pose_kernel|line3d_ransac_kernel|lsd_region_kernel" -c 4 -o gpurun_out/r2_final_top4 -f python bench.py --no-pipeline --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_final_ncu_full.log 2>&1
tail -2 gpurun_out/r2_final_ncu_full.log | cut -c1-200
timeout 300 python tools/tum_probe.py 1184 2>&1 | tail -2 | tee gpurun_out/r2_final_tum_probe.log
