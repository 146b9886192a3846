"""Where the whole-game self-play loop spends time outside the searches (selfplay.BatchedSelfPlay.play).

    python profiles/play_probe.py [--games 8192] [--sims 100] [--profile]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import net as hznet, selfplay as sp  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192)
ap.add_argument("--sims", type=int, default=100)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--slots", type=int, default=4096)
ap.add_argument("--leaves", type=int, default=1)
ap.add_argument("--no-compact", action="store_true")
a = ap.parse_args()
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16)
cfg = sp.SelfPlayConfig(n_slots=a.slots, num_simulations=a.sims, seed=1, leaves_per_step=a.leaves, compact_live=not a.no_compact)
drv = sp.BatchedSelfPlay(inf, cfg, device=dev)
drv.play(min(64, a.slots))
evs = []
orig = drv.search


def timed_search(states):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(states); e1.record(); evs.append((e0, e1))


drv.search = timed_search
if a.profile:
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        traj = drv.play(a.games)
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=12, max_name_column_width=60))
else:
    traj = drv.play(a.games)
ms = [x.elapsed_time(y) for x, y in evs]
gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(len(evs) - 1)]
print(json.dumps({"gap_ms_mean": sum(gaps) / max(1, len(gaps)), "gap_ms_sorted_tail": [round(g, 2) for g in sorted(gaps)[-8:]],
                  "gap_ms_median": sorted(gaps)[len(gaps) // 2] if gaps else None}))
st = traj.stats
print(json.dumps({"search_ms_every_8th_step": [round(m, 2) for m in ms[::8]]}))
print(json.dumps({"steps": len(ms), "loop_s": st["seconds"], "search_s": sum(ms) / 1e3,
                  "outside_search_ms_per_step": (st["seconds"] - sum(ms) / 1e3) / len(ms) * 1e3,
                  "search_ms_per_step": sum(ms) / len(ms), "slot_utilisation": st["examples"] / (len(ms) * a.slots), "stats": st}))
