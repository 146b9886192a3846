#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/g8_bench8.json 2> gpurun_out/g8_bench8.err; echo "rc=$?"; tail -c 300 gpurun_out/g8_bench8.err
