#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 1
timeout 900 bash profiles/regen.sh > /dev/null 2>&1
cp gpurun_out/r02_playout_counters.json gpurun_out/r02_tower_counters.json profiles/
timeout 1200 python bench.py > gpurun_out/final4_bench1.json 2> gpurun_out/final4_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/final4_ref.json 2> gpurun_out/final4_ref.err; echo "ref rc=$?"
