#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/playout_case.py > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_playout -s 2 -c 1 -f -o gpurun_out/r02b_playout_full python profiles/playout_case.py > gpurun_out/ba_ncu.log 2>&1; tail -n 2 gpurun_out/ba_ncu.log
