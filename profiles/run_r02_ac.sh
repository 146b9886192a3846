#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -x 2>&1 | tail -n 5
for i in 1 2 3; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
