#!/bin/bash
mkdir -p gpurun_out
echo "== mma FC"; timeout 900 python -m pytest tests/test_gpu_tower.py tests/test_net_golden.py tests/test_gpu_selfplay.py -q -m gpu -x 2>&1 | tail -n 4
for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'], d['heads_us'], d['step_sum_us'])"; done
echo "== fma FC"; for i in 1 2; do HZ_FC_FMA=1 timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'], d['heads_us'], d['step_sum_us'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"k_heads" -c 2 --csv --log-file gpurun_out/ay_fc.csv python profiles/mcts_step.py --steps 2 > /dev/null 2>&1; grep "gpu__time" gpurun_out/ay_fc.csv | awk -F'","' '{print $5, $(NF)}'
