#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/streams_probe.py > gpurun_out/x_streams.log 2>&1; cat gpurun_out/x_streams.log | tail -n 12
