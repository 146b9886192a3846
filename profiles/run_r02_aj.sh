#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/positions_1m.py 2>&1 | tail -n 1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_score -c 1 -f -o gpurun_out/r02_score_full python profiles/positions_1m.py --iters 1 > gpurun_out/aj_ncu.log 2>&1
ls -la gpurun_out/r02_score_full.ncu-rep
