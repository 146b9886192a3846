"""The headline launch in isolation (for ncu): a few waves of hz_playout at the bench's size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for k in range(4):
    st = hb.init_states(n, seed=1000 + k)
    hb.playout(st)
torch.cuda.synchronize()
print("ok")
