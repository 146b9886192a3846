#!/bin/bash
# evidence after the k_playout rewrite: GPU suite, counters, bench (both arms), playout A/B against the old path, full ncu capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
timeout 900 bash profiles/regen.sh > /dev/null 2>&1
cp gpurun_out/r02_playout_counters.json gpurun_out/r02_tower_counters.json profiles/
timeout 1200 python bench.py > gpurun_out/bb_bench1.json 2> gpurun_out/bb_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bb_bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/bb_ref.json 2> gpurun_out/bb_ref.err; echo "ref rc=$?"
timeout 300 python profiles/playout_ab.py
HZ_LIB_PATH=profiles/_ab/lib_checked.so timeout 300 python profiles/playout_ab.py
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_playout -s 2 -c 1 -f -o gpurun_out/r02_playout_full python profiles/playout_case.py > gpurun_out/bb_ncu.log 2>&1; tail -n 1 gpurun_out/bb_ncu.log
