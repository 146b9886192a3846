#!/bin/bash
# ncu launch list of the engine legs of bench.py (same command, MCTS legs and CPU baselines skipped): the kernel's share of the timed region
mkdir -p gpurun_out
timeout 120 python bench.py --no-mcts --no-cpu-baseline --no-python-reference --steps 5 --warmup 3 > gpurun_out/bg_plain.json 2> gpurun_out/bg_plain.err; echo "plain rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_engine_bench_launches.csv python bench.py --no-mcts --no-cpu-baseline --no-python-reference --steps 5 --warmup 3 > gpurun_out/bg_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_engine_bench_launches.csv
