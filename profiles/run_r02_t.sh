#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_tower.py tests/test_net_golden.py -q -m gpu 2>&1 | tail -4
bash profiles/run_bounds.sh 2>&1 | tail -25
timeout 300 python profiles/mcts_step.py 2>&1 | tail -1
