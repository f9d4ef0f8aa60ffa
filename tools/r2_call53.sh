#!/bin/bash
# Round 2, call 53 (8 GPUs): the default bench line under torchrun as the driver launches it
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_final_bench_cfg2_n8.json 2> gpurun_out/r2_final_bench_cfg2_n8.err
tail -1 gpurun_out/r2_final_bench_cfg2_n8.json | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('n8', round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['h2d_gbs_per_rank'], d['host_ms_each_step'])"
tail -3 gpurun_out/r2_final_bench_cfg2_n8.err | cut -c1-300
