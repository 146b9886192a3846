#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -n 3
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
