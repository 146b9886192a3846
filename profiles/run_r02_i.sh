#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tower.py tests/test_gpu_mcts.py tests/test_gpu_selfplay.py -x -q 2>&1 | tail -8 | tee gpurun_out/i_tests.log
timeout 900 python profiles/mcts_ab.py --moves 3 --play 4096 --json gpurun_out/i_mcts_ab.json 2>&1 | tail -40 | tee gpurun_out/i_mcts.log
