#!/bin/bash
mkdir -p gpurun_out
bash profiles/regen.sh 2>&1 | tail -12
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/j_bench.err
