"""Measurements for the SURVEY.md §8(f) rows built so far (f1 trainer hook / whole-game
self-play, f2 replay ring, f3 arena + greedy tournament), default network, one B200.

    python profiles/next_rows.py [--games 4096] [--sims 100]
"""

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from harmonies_alphazero_b200 import arena  # noqa: E402
from harmonies_alphazero_b200 import net as hznet  # noqa: E402
from harmonies_alphazero_b200 import selfplay as sp  # noqa: E402
from harmonies_alphazero_b200.replay import ReplayRing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=100)
a = ap.parse_args()

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16)
out = {}


def sync_time(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t0


# ---- f1: complete self-play games (trainer.py:62-134,434-541), continuous batching
cfg = sp.SelfPlayConfig(n_slots=min(4096, a.games), num_simulations=a.sims, seed=1)
drv = sp.BatchedSelfPlay(inf, cfg, device=dev)
drv.play(min(64, a.games))                                  # warm-up: graph capture, cuDNN autotune
traj, secs = sync_time(lambda: drv.play(a.games))
ex, conv = sync_time(lambda: traj.to_reference_examples() if len(traj) <= 300000 else None)
out["f1_self_play"] = {
    "games": traj.stats["games"], "examples": len(traj), "seconds": secs, "games_per_s": traj.stats["games"] / secs,
    "examples_per_s": len(traj) / secs, "sims_per_s_live": traj.stats["sims_per_s"], "move_steps": traj.stats["move_steps"],
    "to_reference_examples_s": conv, "note": "whole games incl. refill, trajectory bookkeeping and one host sync per move",
}

# ---- f2: replay ring (buffer.py / trainer.py:136-193)
ring = ReplayRing(50000, device=dev)                         # config.py:87 replay_buffer_size
_, t_ext = sync_time(lambda: ring.extend(traj))
for _ in range(3):
    ring.sample(64)
_, t64 = sync_time(lambda: [ring.sample(64) for _ in range(200)])
_, t4k = sync_time(lambda: [ring.sample(4096) for _ in range(50)])
out["f2_replay_ring"] = {
    "size": len(ring), "bytes_per_example": 128 + 286 + 4, "extend_s": t_ext,
    "sample_batch64_per_s": 200 / t64, "sample_batch4096_examples_per_s": 50 * 4096 / t4k,
    "note": "sample = uniform indices + hz_encode (fp32) + pi normalisation, all on device",
}

# ---- f3: arena (trainer.py:293-431, config.py:67-78: 30 games, 200 sims) and greedy tournament (evaluation.py)
torch.manual_seed(1)
best = hznet.InferenceNet(hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval(), device=dev, dtype=torch.bfloat16)
ecfg = {"num_simulations": 200, "cpuct": 2, "dirichlet_alpha": 0.1, "dirichlet_epsilon": 0, "turns_until_tau0": 0,
        "action_size": 143, "testing": True}
arena.play_match(inf, best, 4, dict(ecfg, num_simulations=8))
r, t_arena = sync_time(lambda: arena.play_match(inf, best, 30, ecfg))
out["f3_arena_30_games_200_sims"] = dict(r, seconds=t_arena)
r, t_big = sync_time(lambda: arena.play_match(inf, best, 1024, dict(ecfg, num_simulations=100)))
out["f3_arena_1024_games_100_sims"] = dict(r, seconds=t_big, games_per_s=1024 / t_big)
r, t_g = sync_time(lambda: arena.play_match(inf, arena.GREEDY, 1024, dict(ecfg, num_simulations=100)))
out["f3_vs_greedy_1024_games_100_sims"] = dict(r, seconds=t_g)
r, t_gg = sync_time(lambda: arena.play_match(arena.GREEDY, arena.GREEDY, 65536, ecfg))
out["f3_greedy_vs_greedy_65536_games"] = dict(r, seconds=t_gg, games_per_s=65536 / t_gg)
print(json.dumps(out))
