#!/bin/bash
# Round 2, call 31: cfg3 after moving the SIFT point block to the stream-ordered pool (no device-wide sync under the pair stream)
timeout 300 python -m pytest tests/test_gpu_sift.py tests/test_gpu_hybrid.py -x -q 2>&1 | tail -3
timeout 400 python bench.py --workload cfg3 --no-cpu > gpurun_out/r2_final_bench_cfg3_n1.json 2> gpurun_out/r2_final_bench_cfg3_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_final_bench_cfg3_n1.json').readline())
print('cfg3', round(d['value']), round(d['e2e']['value']), d['host_ms_each_step'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>1})"
