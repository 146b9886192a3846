#!/bin/bash
mkdir -p gpurun_out
echo "== parity (engine tests)"; timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu -x 2>&1 | tail -n 3
for i in 1 2; do
  timeout 300 python profiles/playout_ab.py
  HZ_LIB_PATH=profiles/_ab/lib_checked.so timeout 300 python profiles/playout_ab.py
done
