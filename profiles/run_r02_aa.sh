#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 3
timeout 1200 python bench.py > gpurun_out/aa_bench1.json 2> gpurun_out/aa_bench1.err; echo "rc=$?"; tail -c 300 gpurun_out/aa_bench1.err
timeout 900 bash profiles/regen.sh
