#!/bin/bash
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v["ms"]*1e3,1), round(v.get("frac_of_hbm_6455.6",0),3)) for k,v in d.items() if k=="score"})'
for t in 64 256 512; do
HZ_NVCC_EXTRA="-DHZ_SCORE_TPB=$t" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
echo "== TPB $t"; for i in 1 2; do timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1 | python -c "$P"; done
done
