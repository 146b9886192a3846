#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/tower_bench.py --json gpurun_out/n_tower_bench.json > gpurun_out/n_bench.log 2>&1; tail -3 gpurun_out/n_bench.log
