#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 2
echo "== pair"; timeout 600 python profiles/play_probe.py --games 8192 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stats']['sims_per_s'], d['search_ms_per_step'])"
echo "== no pair"; HZ_TOWER_NO_PAIR=1 timeout 600 python profiles/play_probe.py --games 8192 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stats']['sims_per_s'], d['search_ms_per_step'])"
echo "== pair"; timeout 600 python profiles/play_probe.py --games 8192 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stats']['sims_per_s'], d['search_ms_per_step'])"
echo "== no pair"; HZ_TOWER_NO_PAIR=1 timeout 600 python profiles/play_probe.py --games 8192 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['stats']['sims_per_s'], d['search_ms_per_step'])"
