#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 600 python profiles/play_probe.py --games 8192 > gpurun_out/v_probe_compact.log 2>&1; tail -3 gpurun_out/v_probe_compact.log
timeout 600 python profiles/play_probe.py --games 8192 --no-compact > gpurun_out/v_probe_plain.log 2>&1; tail -2 gpurun_out/v_probe_plain.log
timeout 600 python profiles/play_probe.py --games 4096 > gpurun_out/v_probe_compact4k.log 2>&1; tail -1 gpurun_out/v_probe_compact4k.log
timeout 900 bash profiles/regen.sh
