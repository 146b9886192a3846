#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu 2>&1 | tail -3
timeout 600 python profiles/tower_bench.py --json gpurun_out/r_bench_late.json > gpurun_out/r_bench1.log 2>&1; tail -2 gpurun_out/r_bench1.log
HZ_NVCC_EXTRA="-DHZ_TOWER_EARLY_HANDOFF=1" python -m harmonies_alphazero_b200.build --force > gpurun_out/r_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu 2>&1 | tail -3
timeout 600 python profiles/tower_bench.py --json gpurun_out/r_bench_early.json > gpurun_out/r_bench2.log 2>&1; tail -2 gpurun_out/r_bench2.log
