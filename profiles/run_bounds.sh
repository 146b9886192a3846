#!/bin/bash
# The whole GPU suite against a bounds-checked build of the library (-DHZ_DEBUG_BOUNDS: every index into a tree
# arena, hash table, path buffer and the water queue is asserted; a violation traps the launch).  compute-sanitizer
# is closed on this pool; this is the replacement evidence.  Run on a B200 through gpurun; writes profiles-ready text.
set -u
mkdir -p gpurun_out
HZ_NVCC_EXTRA="-DHZ_DEBUG_BOUNDS" python - <<'PY'
import os
from harmonies_alphazero_b200 import build
build.LIB = os.path.join(build.HERE, "libharmonies_b200_dbg.so")
print(build.build())   # stale or missing only: build it in the container first, the .so travels
PY
{
  echo "== GPU suite on the bounds-checked build (nvcc -DHZ_DEBUG_BOUNDS, $(date -u +%Y-%m-%dT%H:%MZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1))"
  echo "== HZ_BOUND sites compiled in: $(grep -c 'HZ_BOUND(' harmonies_alphazero_b200/csrc/hz_mcts.cu harmonies_alphazero_b200/csrc/hz_core.cuh | tr '\n' ' ')"
  HZ_LIB_PATH=$PWD/harmonies_alphazero_b200/libharmonies_b200_dbg.so timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6
  echo "== self-play under the bounds-checked build: 512 whole games, 100 simulations per move, default network"
  HZ_LIB_PATH=$PWD/harmonies_alphazero_b200/libharmonies_b200_dbg.so timeout 600 python profiles/mcts_ab.py --games 512 --moves 1 --play 512 --towers hand 2>&1 | tail -14
} | tee gpurun_out/r02_bounds_check.txt
