#!/bin/bash
mkdir -p gpurun_out
for k in test_tile_layout_roundtrip test_conv_bit_exact_on_integers test_conv_many_tiles test_conv_random_values test_whole_tower; do
  timeout 300 python -m pytest tests/test_gpu_tower.py -x -q -k $k 2>&1 | tail -15
done 2>&1 | tee gpurun_out/c_tests.log
timeout 600 python profiles/tower_bench.py --json gpurun_out/c_tower_bench.json 2>&1 | tail -80 | tee gpurun_out/c_bench.log
