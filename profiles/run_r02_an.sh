#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -n 2
P='import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v["ms"]*1e3,1), round(v.get("frac_of_hbm_6455.6",0),3)) for k,v in d.items() if k=="score"})'
for i in 1 2; do timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1 | python -c "$P"; done
for i in 1 2; do timeout 300 python bench.py --no-mcts --no-cpu-baseline --no-python-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"; done
