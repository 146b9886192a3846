#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_net_golden.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/o_tests.log
timeout 600 python profiles/tower_bench.py --json gpurun_out/o_tower_bench.json > gpurun_out/o_bench.log 2>&1; tail -3 gpurun_out/o_bench.log
timeout 900 python profiles/mcts_ab.py --moves 3 --towers hand --play 4096 --json gpurun_out/o_mcts_ab.json 2>&1 | tail -16
