#!/bin/bash
mkdir -p gpurun_out
echo "== cluster 4"; HZ_TOWER_CLUSTER=4 timeout 600 python -m pytest tests/test_gpu_tower.py tests/test_net_golden.py -q -m gpu -x 2>&1 | tail -n 2
for c in 4 2 4 2; do echo "== cluster $c"; HZ_TOWER_CLUSTER=$c timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'])"
HZ_TOWER_CLUSTER=$c timeout 600 python profiles/play_probe.py --games 8192 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stats']['sims_per_s'], d['search_ms_per_step'])"; done
