#!/bin/bash
# ncu --set full of the warp-staged k_legal (and k_hash beside it) at 1 M positions
mkdir -p gpurun_out
timeout 300 python profiles/positions_1m.py --iters 3 > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"k_legal|k_hash" -c 2 -f -o gpurun_out/r02_legal_hash python profiles/positions_1m.py --iters 1 > gpurun_out/be_ncu.log 2>&1; tail -n 1 gpurun_out/be_ncu.log
