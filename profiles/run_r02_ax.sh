#!/bin/bash
# does a second busy GPU on the same box slow the whole-game loop of the first?  (no NCCL, independent processes)
mkdir -p gpurun_out
P='import json,sys; L=sys.stdin.read().strip().splitlines(); d=json.loads(L[-1]); g=json.loads(L[-3]); print(round(d["stats"]["sims_per_s"]/1e6,2), "M sims/s; search ms/step", round(d["search_ms_per_step"],2), "outside ms/step", round(d["outside_search_ms_per_step"],2), "gap median", round(g["gap_ms_median"],2))'
echo "== alone on GPU 0"; CUDA_VISIBLE_DEVICES=0 timeout 600 python profiles/play_probe.py --games 8192 2>&1 | python -c "$P"
echo "== two at once"
CUDA_VISIBLE_DEVICES=0 timeout 600 python profiles/play_probe.py --games 8192 > gpurun_out/ax_a.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 600 python profiles/play_probe.py --games 8192 > gpurun_out/ax_b.log 2>&1 &
wait
python -c "$P" < gpurun_out/ax_a.log; python -c "$P" < gpurun_out/ax_b.log
nproc; nvidia-smi --query-gpu=index,power.limit,enforced.power.limit --format=csv,noheader
