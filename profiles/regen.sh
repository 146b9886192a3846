#!/bin/bash
# Regenerates the ncu-derived evidence bench.py reads (run on a B200 through gpurun, one ncu pass each,
# each preceded by the same command without ncu):  profiles/regen.sh
set -u
mkdir -p gpurun_out
CS=harmonies_alphazero_b200/csrc
M=sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__cycles_active.avg
python profiles/playout_case.py > gpurun_out/playout_plain.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:k_playout -s 2 -c 2 --csv --log-file gpurun_out/r02_playout_counters.csv python profiles/playout_case.py > gpurun_out/playout_ncu.log 2>&1
python profiles/counters_to_json.py gpurun_out/r02_playout_counters.csv gpurun_out/r02_playout_counters.json k_playout $CS/hz_engine.cu $CS/hz_core.cuh $CS/hz_tables.inc > gpurun_out/playout_counters.log 2>&1
T=$M,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum
python profiles/tower_case.py > gpurun_out/tower_plain.log 2>&1 &&
ncu --metrics $T --clock-control none -k regex:k_tower -s 2 -c 2 --csv --log-file gpurun_out/r02_tower_counters.csv python profiles/tower_case.py > gpurun_out/tower_ncu.log 2>&1
python profiles/counters_to_json.py gpurun_out/r02_tower_counters.csv gpurun_out/r02_tower_counters.json k_tower $CS/hz_tower.cu $CS/hz_sm100.cuh > gpurun_out/tower_counters.log 2>&1
tail -n 3 gpurun_out/playout_counters.log gpurun_out/tower_counters.log
# back in the container: cp gpurun_out/r02_playout_counters.json gpurun_out/r02_tower_counters.json profiles/
