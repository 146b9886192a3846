#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 2
for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['tower_us'], d['tower_and_heads_us'], d['step_sum_us'])"; done
timeout 600 python bench.py --no-mcts --no-cpu-baseline --no-python-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
