#!/bin/bash
# Round 2, call 40: cfg3 at 592 and cfg5 at 148 frames per step (region growing amortised)
for wl in cfg3 cfg5; do
  timeout 500 python bench.py --workload $wl --no-cpu > gpurun_out/r2_final_bench_${wl}_n1.json 2> gpurun_out/r2_final_bench_${wl}_n1.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_final_bench_${wl}_n1.json').readline())
print('$wl', d['config'].get('batch_frames_per_gpu'), round(d['value']), round(d['e2e']['value']), d['host_ms_each_step']['value'], {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if v>3})"
  tail -2 gpurun_out/r2_final_bench_${wl}_n1.err | cut -c1-200
done
