"""A/B inside one process: 1x1 head convolutions as a work item of the tower launch vs the separate kernel
(configs[3] step rate, CUDA-graph replay).   python profiles/heads_ab.py [--moves 3] [--rounds 3]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb, net as hznet, selfplay as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--moves", type=int, default=3)
ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
drivers = {}
for fold in (True, False):
    inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16, tower="hand")
    inf.heads_in_tower = fold
    drivers[fold] = sp.BatchedSelfPlay(inf, sp.SelfPlayConfig(n_slots=a.games, num_simulations=100, seed=77), device=dev)
states0 = hb.init_states(a.games, device=dev, seed=77)
hb.playout(states0, max_steps=8)
u01 = torch.rand(a.games, device=dev)
for r in range(a.rounds):
    for fold in (True, False):
        drv = drivers[fold]
        states = states0.clone()

        def one_move():
            drv.search(states)
            hb.apply(states, drv.choose(u01, None))

        one_move()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.moves):
            one_move()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(json.dumps({"round": r, "heads_in_tower": fold, "us_per_sim_step": 1e3 * ms / a.moves / 100,
                          "sims_per_s": a.games * 100 * a.moves / (ms * 1e-3)}), flush=True)
