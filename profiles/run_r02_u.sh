#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/mcts_step.py 2>&1 | tail -1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --mcts-play-games 4096 > gpurun_out/u_bench8.json 2> gpurun_out/u_bench8.err; echo "rc=$?"; tail -c 400 gpurun_out/u_bench8.err
