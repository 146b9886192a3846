"""A/B of the MCTS simulation step (configs[3]: 4,096 trees, default net, bf16, 100 sims/move) with
the cuDNN tower and the hand-written sm_100a tower: sims/s over a few moves (CUDA-graph replay),
then whole games.   python profiles/mcts_ab.py [--games 4096] [--moves 3] [--play 4096] [--json out]"""

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb  # noqa: E402
from harmonies_alphazero_b200 import net as hznet  # noqa: E402
from harmonies_alphazero_b200 import selfplay as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=100)
ap.add_argument("--moves", type=int, default=3)
ap.add_argument("--play", type=int, default=0, help="also play this many whole games per tower")
ap.add_argument("--towers", default="cudnn,hand")
ap.add_argument("--json", default=None)
a = ap.parse_args()

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
out = {}
for tower in a.towers.split(","):
    torch.manual_seed(0)
    model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
    inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16, tower=tower)
    cfg = sp.SelfPlayConfig(n_slots=a.games, num_simulations=a.sims, seed=77)
    drv = sp.BatchedSelfPlay(inf, cfg, device=dev)
    states = hb.init_states(a.games, device=dev, seed=77)
    hb.playout(states, max_steps=8)
    u01 = torch.rand(a.games, device=dev)

    def one_move():
        drv.search(states)
        hb.apply(states, drv.choose(u01, None))

    one_move(); one_move()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.moves):
        one_move()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    drv.check_status()
    r = {"ms_per_move": ms / a.moves, "us_per_sim_step": 1e3 * ms / a.moves / a.sims, "sims_per_s": a.games * a.sims * a.moves / (ms * 1e-3),
         "tiles_path": bool(drv.groups[0].tiles)}
    if a.play:
        t0 = time.perf_counter()
        traj = drv.play(a.play)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        r["whole_games"] = {"games": traj.stats["games"], "seconds": dt, "games_per_s": traj.stats["games"] / dt,
                            "sims_per_s": traj.stats["sims"] / dt, "move_steps": traj.stats["move_steps"], "examples": len(traj)}
    out[tower] = r
    del drv, inf
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
if a.json:
    json.dump(out, open(a.json, "w"), indent=1)
