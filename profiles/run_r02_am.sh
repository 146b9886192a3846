#!/bin/bash
# r02 ncu --set full captures of the other kernels of the two hot paths (one process each, each preceded by a plain run)
mkdir -p gpurun_out
timeout 300 python profiles/mcts_step.py > gpurun_out/am_step.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_tree_select|k_tree_expand_backup|k_heads_p" -c 3 -f -o gpurun_out/r02_step_kernels python profiles/mcts_step.py --steps 1 > gpurun_out/am_ncu1.log 2>&1
timeout 300 python profiles/playout_case.py > gpurun_out/am_playout.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_playout -s 2 -c 1 -f -o gpurun_out/r02_playout_full python profiles/playout_case.py > gpurun_out/am_ncu2.log 2>&1
ls -la gpurun_out/r02_step_kernels.ncu-rep gpurun_out/r02_playout_full.ncu-rep
