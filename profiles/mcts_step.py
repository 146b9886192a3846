"""Profiling harness: a few eager simulation steps of configs[3] (4,096 trees, default net,
bf16) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

    python profiles/mcts_step.py [--games 4096] [--sims-before 40] [--steps 2]
"""

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb  # noqa: E402
from harmonies_alphazero_b200 import net as hznet  # noqa: E402
from harmonies_alphazero_b200 import selfplay as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims-before", type=int, default=40, help="untimed simulations so that trees have realistic depth")
ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16)
cfg = sp.SelfPlayConfig(n_slots=a.games, num_simulations=100, use_cuda_graph=False, seed=77)
drv = sp.BatchedSelfPlay(inf, cfg, device=dev)
states = hb.init_states(a.games, device=dev, seed=77)
hb.playout(states, max_steps=8)
drv.tree.reset(states)
for _ in range(a.sims_before):
    drv._sim_step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    drv._sim_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
drv.tree.check_status()
print("ok")
