#!/bin/bash
mkdir -p gpurun_out
echo "== 4+1 split (default)"; timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -x 2>&1 | tail -n 3
for i in 1 2 3; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
HZ_NVCC_EXTRA="-DHZ_TOWER_SPLIT41=0" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
echo "== 3+2 split"; timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -x 2>&1 | tail -n 2
for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
HZ_NVCC_EXTRA="-DHZ_TOWER_TRACE=1" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
timeout 300 python profiles/tower_trace.py --json gpurun_out/tower_trace_41.json 2>&1 | tail -n 1
