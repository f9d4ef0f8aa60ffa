#!/bin/bash
# Round 2, call 33 (4 GPUs): the default bench line under torchrun as the driver launches it
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_final_bench_cfg2_n4.json 2> gpurun_out/r2_final_bench_cfg2_n4.err
tail -1 gpurun_out/r2_final_bench_cfg2_n4.json | cut -c1-1500
tail -3 gpurun_out/r2_final_bench_cfg2_n4.err | cut -c1-300
