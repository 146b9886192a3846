#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/tower_trace.py --json gpurun_out/tower_trace.json 2>&1 | tail -n 2
timeout 300 python profiles/tower_trace.py --debug 14 --json gpurun_out/tower_trace_mmaonly.json 2>&1 | tail -n 1
