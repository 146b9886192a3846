#!/bin/bash
# A/B of the streaming kernels (k_legal, k_hash) at 1 M positions: default library against profiles/_ab/lib_prev.so
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -n 2
for i in 1 2; do
  timeout 300 python profiles/positions_1m.py --iters 9 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new ', {k:(round(v['ms']*1e3,1), round(v.get('algorithmic_GBps',v.get('GBps',0)))) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
  HZ_LIB_PATH=profiles/_ab/lib_prev.so timeout 300 python profiles/positions_1m.py --iters 9 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prev', {k:(round(v['ms']*1e3,1), round(v.get('algorithmic_GBps',v.get('GBps',0)))) for k,v in d.items() if isinstance(v,dict) and 'ms' in v})"
done
