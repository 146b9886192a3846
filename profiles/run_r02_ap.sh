#!/bin/bash
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v["ms"]*1e3,1), round(v.get("frac_of_hbm_6455.6",0),3)) for k,v in d.items() if k in ("legal_mask","canon_hash_ref","score")})'
for h in 0 1 2; do
HZ_NVCC_EXTRA="-DHZ_LOAD_HINT=$h" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
echo "== hint $h"; for i in 1 2; do timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1 | python -c "$P"; done
timeout 300 python bench.py --no-mcts --no-cpu-baseline --no-python-reference 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('playout', d['value'], 'unfused', d['unfused']['value'])"
done
