#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -n 3
timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['ms']*1e3,1), round(v.get('frac_of_hbm_6455.6',0),3)) for k,v in d.items() if 'ms' in v})"
timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['ms']*1e3,1), round(v.get('frac_of_hbm_6455.6',0),3)) for k,v in d.items() if 'ms' in v})"
