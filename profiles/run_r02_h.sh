#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tower.py -x -q 2>&1 | tail -3 | tee gpurun_out/h_tests.log
timeout 600 python profiles/tower_bench.py --json gpurun_out/h_bench_collector.json > gpurun_out/h_bench1.log 2>&1; tail -3 gpurun_out/h_bench1.log
HZ_NVCC_EXTRA="-DHZ_TOWER_COLLECTOR=0" python -m harmonies_alphazero_b200.build --force > gpurun_out/h_build.log 2>&1
timeout 600 python profiles/tower_bench.py --json gpurun_out/h_bench_nocollector.json > gpurun_out/h_bench2.log 2>&1; tail -3 gpurun_out/h_bench2.log
