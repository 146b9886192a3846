#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/rec_bench1.json 2> gpurun_out/rec_bench1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/rec_ref.json 2> gpurun_out/rec_ref.err; echo "ref rc=$?"
