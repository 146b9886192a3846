#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/w_bench1.json 2> gpurun_out/w_bench1.err; echo "rc=$?"; tail -c 300 gpurun_out/w_bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/w_ref.json 2> gpurun_out/w_ref.err; echo "rc=$?"; tail -c 600 gpurun_out/w_ref.json
timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1
