#!/bin/bash
mkdir -p gpurun_out
health() { timeout 60 nvidia-smi --query-gpu=index,memory.used --format=csv,noheader || { echo "GPU UNHEALTHY after $1"; exit 7; }; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29541 bench.py --gpus 2 --steps 30 --warmup 5 --cuda-graphs 1 > gpurun_out/bench2_g1_full.log 2>&1; health g1
grep -v "^$" gpurun_out/bench2_g1_full.log | grep -v "^\[rank1\]" | tail -40 | cut -c1-400
