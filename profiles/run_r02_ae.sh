#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/heads_ab.py --rounds 4 2>&1 | tail -n 8
