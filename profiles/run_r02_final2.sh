#!/bin/bash
# end-of-round evidence run (1 GPU) after the last kernel changes: smoke, GPU suite, bench line, reference arm, ncu counters
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 1
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 2
timeout 1200 python bench.py > gpurun_out/final2_bench1.json 2> gpurun_out/final2_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final2_bench1.err
timeout 900 bash profiles/regen.sh
timeout 1200 python bench.py > gpurun_out/final2_bench1b.json 2> gpurun_out/final2_bench1b.err; echo "bench rc=$?"
