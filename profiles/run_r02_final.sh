#!/bin/bash
# end-of-round evidence run (1 GPU): GPU suite, bench line, ncu counters, launch list of the MCTS step, full capture of the tower
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
timeout 1200 python bench.py > gpurun_out/final_bench1.json 2> gpurun_out/final_bench1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench1.err
timeout 600 python bench.py --impl reference > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"
timeout 900 bash profiles/regen.sh
timeout 300 python profiles/mcts_step.py > gpurun_out/final_step.log 2>&1 && tail -n 1 gpurun_out/final_step.log &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_mcts_launches.csv python profiles/mcts_step.py > gpurun_out/final_step_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tower -s 2 -c 1 -f -o gpurun_out/r02_tower_full python profiles/tower_case.py > gpurun_out/final_tower_ncu.log 2>&1
ls -la gpurun_out/r02_tower_full.ncu-rep
