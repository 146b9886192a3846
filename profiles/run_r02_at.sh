#!/bin/bash
mkdir -p gpurun_out
echo "== 3+2 (default) pair"; for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'])"; done
HZ_NVCC_EXTRA="-DHZ_TOWER_SPLIT41=1" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
echo "== 4+1 pair"; timeout 300 python -m pytest tests/test_gpu_tower.py -q -m gpu -x -k "bit_exact or fused_launch or head" 2>&1 | tail -n 1
for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'])"; done
