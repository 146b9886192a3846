#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tower.py -x -q 2>&1 | tail -5 | tee gpurun_out/e_tests.log
timeout 600 python profiles/tower_bench.py --json gpurun_out/e_tower_bench.json > gpurun_out/e_bench.log 2>&1; tail -5 gpurun_out/e_bench.log
