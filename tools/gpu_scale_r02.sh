#!/bin/bash
# Round-2 multi-GPU session (one 8-GPU box): strong-scaling bench at N=8 and N=4 and C5 on 8 GPUs.
# Usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale_r02.sh > gpurun_out/scale_r02.log 2>&1'
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -3 gpurun_out/bench_n8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29512 tools/run_c5.py > gpurun_out/c5_n8.json 2> gpurun_out/c5_n8.err
tail -3 gpurun_out/c5_n8.err
timeout 240 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -3 gpurun_out/bench_n4.err
cut -c1-300 gpurun_out/bench_n8.json gpurun_out/bench_n4.json gpurun_out/c5_n8.json
