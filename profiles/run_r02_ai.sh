#!/bin/bash
mkdir -p gpurun_out
echo "== discard on"; timeout 900 python -m pytest tests/test_gpu_tower.py -q -m gpu -x 2>&1 | tail -n 2
timeout 600 python profiles/heads_ab.py --rounds 3 2>&1 | grep true
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
ncu --metrics $M --clock-control none -k regex:k_tower -s 2 -c 1 --csv --log-file gpurun_out/ai_on.csv python profiles/tower_case.py > /dev/null 2>&1; grep -E "dram__bytes|gpu__time" gpurun_out/ai_on.csv | cut -d, -f13-15
HZ_NVCC_EXTRA="-DHZ_TOWER_DISCARD=0" python -m harmonies_alphazero_b200.build --force > gpurun_out/z_build.log 2>&1
echo "== discard off"; timeout 600 python profiles/heads_ab.py --rounds 3 2>&1 | grep true
ncu --metrics $M --clock-control none -k regex:k_tower -s 2 -c 1 --csv --log-file gpurun_out/ai_off.csv python profiles/tower_case.py > /dev/null 2>&1; grep -E "dram__bytes|gpu__time" gpurun_out/ai_off.csv | cut -d, -f13-15
