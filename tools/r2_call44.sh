#!/bin/bash
# Round 2, call 44: SIFT in passes of up to 148 frames — tests and the cfg3 line at its default batch
timeout 400 python -m pytest tests/test_gpu_sift.py tests/test_gpu_hybrid.py -q 2>&1 | tail -2
timeout 500 python bench.py --workload cfg3 --no-cpu > gpurun_out/r2_final_bench_cfg3_n1.json 2> gpurun_out/r2_final_bench_cfg3_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_final_bench_cfg3_n1.json').readline())
print('cfg3', d['config'].get('batch_frames_per_gpu'), round(d['value']), round(d['e2e']['value']), d['host_ms_each_step']['value'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>3})"
