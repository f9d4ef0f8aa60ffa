#!/bin/bash
# Last pass of the round on one B200: racecheck of the PNG kernels, all GPU tests, smoke(), bench (default flags).
timeout 200 compute-sanitizer --tool racecheck --print-limit 40 python -m pytest tests/test_gpu_tum.py -q -m gpu -x -k "shape0 or shape1 or shape2 or rejects" 2>&1 | grep -v "^=========     " | tail -30 > gpurun_out/sanitizer_racecheck_tum.log
tail -4 gpurun_out/sanitizer_racecheck_tum.log
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r1g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r1g_smoke.log
python bench.py > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench_n1.err
cut -c1-330 gpurun_out/r1g_bench_n1.json
python tools/tum_probe.py 592 2>&1 | tail -2 | tee gpurun_out/r1g_tum_probe.log
