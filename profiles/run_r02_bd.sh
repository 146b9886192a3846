#!/bin/bash
# final evidence of the round (1 GPU): build + smoke, GPU suite, counters, bench (both arms), 1 M-position harness, launch list of the bench
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 2
timeout 900 bash profiles/regen.sh > /dev/null 2>&1
cp gpurun_out/r02_playout_counters.json gpurun_out/r02_tower_counters.json profiles/
timeout 1200 python bench.py > gpurun_out/bd_bench1.json 2> gpurun_out/bd_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bd_bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/bd_ref.json 2> gpurun_out/bd_ref.err; echo "ref rc=$?"
timeout 300 python profiles/positions_1m.py --iters 9 > gpurun_out/r02_positions_1m.json 2> gpurun_out/bd_pos.err; echo "pos rc=$?"
