#!/bin/bash
# last evidence run: engine tests, counters for the final sources, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -n 2
timeout 900 bash profiles/regen.sh > /dev/null 2>&1
cp gpurun_out/r02_playout_counters.json gpurun_out/r02_tower_counters.json profiles/
timeout 1200 python bench.py > gpurun_out/bf_bench1.json 2> gpurun_out/bf_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bf_bench1.err
timeout 300 python profiles/positions_1m.py --iters 9 > gpurun_out/r02_positions_1m.json 2> gpurun_out/bf_pos.err; echo "pos rc=$?"
