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
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16, tower=os.environ.get("HZ_TOWER", "auto"))
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

# CUDA-event timing of the four parts of one simulation step (eager, warm L2), median of 20
g = drv.groups[0]


def timed(fn, reps=20):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400_000)     # keep the GPU busy while the host enqueues: device time, not launch latency
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ms)[len(ms) // 2]


parts = {}
for _ in range(3):
    parts["select_us"] = timed(lambda: g.tree.select(cfg.cpuct, g.board, g.glob, dtype=inf.dtype, channels_last=True, pad40=drv.pad40, tiles=g.tiles), 1)
    if g.tiles:       # hand-written tower: leaves are already in the stem's operand image
        n = g.glob.shape[0]
        parts["tower_us"] = timed(lambda: inf.hand.forward_tiles(g.board, n), 5)
        parts["tower_and_heads_us"] = timed(lambda: inf.forward_tiles(g.board, g.glob, n, out=(g.logits, g.value)), 5)
        parts["heads_us"] = parts["tower_and_heads_us"] - parts["tower_us"]
    else:
        x = inf.tower_out(g.board)
        parts["tower_us"] = timed(lambda: inf.tower_out(g.board), 5)
        parts["heads_us"] = timed(lambda: inf._fused_heads(x, g.glob, (g.logits, g.value)), 5)
    parts["expand_us"] = timed(lambda: g.tree.expand_backup(g.logits, g.value, is_logits=True, noise=g.noise, eps=cfg.dirichlet_epsilon), 1)
parts["step_sum_us"] = parts["select_us"] + parts["tower_us"] + parts["heads_us"] + parts["expand_us"]
parts["tower"] = "hand" if g.tiles else "cudnn"
import json  # noqa: E402

print(json.dumps(parts))
