#!/bin/bash
# Round 2, call 26 (2 GPUs): cfg2 as ONE stream dealt block-wise to two ranks, block tails shifted rank to rank (lsl_shift_frame), NCCL log kept
export NCCL_DEBUG=INFO
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --split stream --steps 3 --warmup 3 --no-cpu > gpurun_out/r2n_bench_cfg2_split_n2.json 2> gpurun_out/r2n_cfg2_split_n2.err
grep -v "NCCL INFO" gpurun_out/r2n_bench_cfg2_split_n2.json | tail -1 | cut -c1-1400
grep -h "NCCL INFO" gpurun_out/r2n_bench_cfg2_split_n2.json gpurun_out/r2n_cfg2_split_n2.err | grep -iE "Send|Recv|P2P|nranks|Connected|via" | head -12
tail -5 gpurun_out/r2n_cfg2_split_n2.err | cut -c1-300
