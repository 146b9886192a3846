#!/bin/bash
# A/B of k_playout builds: the default library against the ones under profiles/_ab/ (HZ_LIB_PATH)
mkdir -p gpurun_out
echo "== parity (engine tests)"; timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -n 3
for i in 1 2; do
  timeout 300 python profiles/playout_ab.py
  for l in profiles/_ab/*.so; do HZ_LIB_PATH=$l timeout 300 python profiles/playout_ab.py; done
done
