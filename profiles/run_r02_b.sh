#!/bin/bash
# round-2 GPU call B: MN-major descriptor probe, tower parity tests, tower timing vs cuDNN
mkdir -p gpurun_out
( cd profiles
  for v in "112 0 0 1 1 1" "112 0 0 1 1 2" "96 16 16 1 2 1" "112 112 128 1 2 1" "224 8 0 1 1 1" "256 16 256 1 2 1"; do
    timeout 60 ./umma_probe $v
  done ) 2>&1 | tee gpurun_out/b_probe.log
for k in test_tile_layout_roundtrip test_conv_bit_exact_on_integers test_conv_many_tiles test_conv_random_values test_whole_tower; do
  timeout 300 python -m pytest tests/test_gpu_tower.py -x -q -k $k 2>&1 | tail -15
done 2>&1 | tee gpurun_out/b_tests.log
timeout 600 python profiles/tower_bench.py --json gpurun_out/b_tower_bench.json 2>&1 | tail -60 | tee gpurun_out/b_bench.log
