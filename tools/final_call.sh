#!/bin/bash
# End-of-round measurement pass on one B200 (run under gpurun): GPU tests, bench (both arms), ncu launch list, ncu full.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1f_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1f_bench_n1.json 2> gpurun_out/r1f_bench_n1.err
cut -c1-400 gpurun_out/r1f_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1f_bench_ref.json 2>/dev/null
cut -c1-200 gpurun_out/r1f_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1f_launches_b592.csv python bench.py --batch 592 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"line_mle|lsd_region|pose_kernel|line3d_ransac" -s 12 -c 4 -o /tmp/r1f_full python bench.py --batch 592 --steps 1 --warmup 3 --no-cpu > gpurun_out/r1f_ncu_full.log 2>&1
ncu -i /tmp/r1f_full.ncu-rep --page raw --csv > gpurun_out/r1f_full_raw_b592.csv
ncu --set full --clock-control none -k regex:"png_unfilter" -s 2 -c 2 -o /tmp/r1f_png python tools/tum_probe.py 148 > gpurun_out/r1f_ncu_png.log 2>&1
ncu -i /tmp/r1f_png.ncu-rep --page raw --csv > gpurun_out/r1f_png_raw.csv
python tools/tum_probe.py 592 > gpurun_out/r1f_tum_probe.log 2>&1
python bench.py --batch 1184 --steps 5 --warmup 3 --no-cpu 2>/dev/null | cut -c1-300 > gpurun_out/r1f_bench_b1184.json
