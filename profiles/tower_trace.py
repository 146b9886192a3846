"""Raw role timeline of CTA 0 of one fused tower launch (hz_tower_set_trace), after warm-up launches.
    python profiles/tower_trace.py [--boards 4096] [--debug 0] --json gpurun_out/tower_trace.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harmonies_alphazero_b200 import net as hnet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--boards", type=int, default=4096)
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--json", default="gpurun_out/tower_trace.json")
a = ap.parse_args()
torch.manual_seed(0)
model = hnet.AlphaZeroNet.from_config(hnet.DEFAULT_MODEL_CONFIG).eval()
hand = hnet.InferenceNet(model, tower="hand")
ht = hand.hand
B = a.boards
board = torch.zeros((B, 40, 5, 7), dtype=torch.bfloat16, device="cuda").contiguous(memory_format=torch.channels_last)
board[:, :38] = (torch.rand((B, 38, 5, 7), device="cuda") < 0.15).to(torch.bfloat16)
x0 = ht.x0_buffer(B)
ht.to_tiles(board, 40, True, x0)
ht.lib.hz_tower_set_debug(a.debug)
for _ in range(200):
    ht.forward_tiles(x0, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ht.forward_tiles(x0, B)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
tr = torch.zeros(4096, dtype=torch.int64, device="cuda")
if ht.lib.hz_tower_set_trace(tr.data_ptr()) != 0:          # library built without -DHZ_TOWER_TRACE=1: timing only
    ht.lib.hz_tower_set_debug(0)
    print(json.dumps({"us_per_launch_untraced": us, "trace": None}))
    raise SystemExit(0)
ht.forward_tiles(x0, B)
torch.cuda.synchronize()
ht.lib.hz_tower_set_trace(None)
ht.lib.hz_tower_set_debug(0)
t = tr.cpu().tolist()
json.dump({"us_per_launch_untraced": us, "trace": t}, open(a.json, "w"))
print(json.dumps({"us_per_launch_untraced": us, "cycles": t[1] - t[0], "ns": t[3] - t[2]}))
