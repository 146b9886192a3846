#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 3
echo "== PDL on"; timeout 600 python profiles/heads_ab.py --rounds 2 2>&1 | grep true
echo "== PDL off"; HZ_NO_PDL=1 timeout 600 python profiles/heads_ab.py --rounds 2 2>&1 | grep true
echo "== PDL on"; timeout 600 python profiles/heads_ab.py --rounds 2 2>&1 | grep true
echo "== PDL off"; HZ_NO_PDL=1 timeout 600 python profiles/heads_ab.py --rounds 2 2>&1 | grep true
