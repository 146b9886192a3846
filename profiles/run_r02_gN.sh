#!/bin/bash
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err; echo "rc=$?"; tail -c 200 gpurun_out/g${N}_bench.err
