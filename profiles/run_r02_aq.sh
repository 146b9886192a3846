#!/bin/bash
mkdir -p gpurun_out
echo "== pair mode"; timeout 600 python -m pytest tests/test_gpu_tower.py tests/test_net_golden.py -q -m gpu -x 2>&1 | tail -n 4
for i in 1 2; do timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
echo "== no pair"; for i in 1 2; do HZ_TOWER_NO_PAIR=1 timeout 300 python profiles/mcts_step.py 2>&1 | tail -n 1; done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:k_tower -s 2 -c 1 --csv --log-file gpurun_out/aq_pair.csv python profiles/tower_case.py > /dev/null 2>&1
HZ_TOWER_NO_PAIR=1 ncu --metrics $M --clock-control none -k regex:k_tower -s 2 -c 1 --csv --log-file gpurun_out/aq_nopair.csv python profiles/tower_case.py > /dev/null 2>&1
