#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -x 2>&1 | tail -n 3
echo "== turns 1/2 (default)"; for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
for v in "1 1" "1 3" "1 9" "2 2" "3 3"; do
  set -- $v
  HZ_NVCC_EXTRA="-DHZ_TOWER_TURN_P0=$1 -DHZ_TOWER_TURN_P1=$2" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
  echo "== turns $1/$2"; timeout 200 python -m pytest tests/test_gpu_tower.py -q -m gpu -x -k "bit_exact or fused_launch" 2>&1 | tail -n 1
  for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
done
