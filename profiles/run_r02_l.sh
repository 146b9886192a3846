#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_golden.py tests/test_gpu_mcts.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/l_tests.log
( cd profiles && timeout 120 ./umma_rate ) 2>&1 | tee gpurun_out/l_rate.log
timeout 300 python profiles/mcts_step.py > gpurun_out/l_step_plain.log 2>&1 && tail -1 gpurun_out/l_step_plain.log &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_mcts_launches.csv python profiles/mcts_step.py > gpurun_out/l_step_ncu.log 2>&1
timeout 300 python profiles/tower_case.py > gpurun_out/l_tower_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tower -s 2 -c 1 -o gpurun_out/r02_tower_full python profiles/tower_case.py > gpurun_out/l_tower_ncu.log 2>&1
ls -la gpurun_out/r02_tower_full.ncu-rep
