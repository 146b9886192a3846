"""Two half-batch groups on two streams with the persistent tower held below the SM count, so that the
tree kernels of one half run on the spare SMs under the other half's tower (configs[3] step rate).

    python profiles/streams_probe.py [--games 4096] [--moves 3] [--configs 1:0,2:0,2:140,2:132]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import _lib, batched as hb, net as hznet, selfplay as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--sims", type=int, default=100)
ap.add_argument("--moves", type=int, default=3)
ap.add_argument("--configs", default="1:0,2:0,2:140,2:132,2:124,3:132,4:132")
a = ap.parse_args()
dev = torch.device("cuda", 0)
lib = _lib.load()
torch.manual_seed(0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16, tower="hand")
for c in a.configs.split(","):
    ns, ctas = (int(x) for x in c.split(":"))
    lib.hz_tower_set_max_ctas(ctas)
    cfg = sp.SelfPlayConfig(n_slots=a.games, num_simulations=a.sims, seed=77, n_streams=ns)
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
    print(json.dumps({"streams": ns, "tower_max_ctas": ctas, "ms_per_move": ms / a.moves, "us_per_sim_step": 1e3 * ms / a.moves / a.sims,
                      "sims_per_s": a.games * a.sims * a.moves / (ms * 1e-3)}), flush=True)
    del drv
lib.hz_tower_set_max_ctas(0)
