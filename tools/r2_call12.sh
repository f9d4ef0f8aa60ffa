#!/bin/bash
# two GPUs: the default workload and the cfg-4 loop-closure batch (query broadcast + sharded keyframes), NCCL log kept
export NCCL_DEBUG=INFO
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_cfg2_n2.json 2> gpurun_out/r2g_cfg2_n2.err
grep -v "NCCL INFO" gpurun_out/r2g_bench_cfg2_n2.json | tail -1 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload cfg4 --steps 5 --warmup 3 > gpurun_out/r2g_bench_cfg4_n2.json 2> gpurun_out/r2g_cfg4_n2.err
grep -v "NCCL INFO" gpurun_out/r2g_bench_cfg4_n2.json | tail -1 | cut -c1-500
grep -h "NCCL INFO" gpurun_out/r2g_bench_cfg4_n2.json gpurun_out/r2g_cfg4_n2.err | grep -iE "Broadcast|AllGather|nranks|Connected|comm 0x" | head -12
tail -3 gpurun_out/r2g_cfg4_n2.err | cut -c1-300
