#!/bin/bash
# attribution of the fused tower's time: the same launch with parts switched off (results are garbage, timing only)
mkdir -p gpurun_out
for d in 0 4 8 12 2 1 3 14 15; do echo "debug=$d"; timeout 200 python profiles/tower_trace.py --debug $d --json gpurun_out/ad_$d.json 2>&1 | tail -n 1; done
