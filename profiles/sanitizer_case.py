import sys; sys.path.insert(0,'.')
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from harmonies_alphazero_b200 import batched as hb, net, selfplay as sp, tree as tr
st = hb.init_states(777, seed=3); hb.playout(st, max_steps=30)
for dt in (torch.float32, torch.bfloat16):
    for cl in (False, True):
        hb.encode(st, dtype=dt, channels_last=cl)
hb.score(st, with_terms=True); hb.legal_mask(st); hb.canon_hash(st, 1); hb.outcome(st); hb.random_actions(st)
torch.manual_seed(0)
m = net.AlphaZeroNet.from_config(net.TEST_MODEL_CONFIG).eval()
inf = net.InferenceNet(m, dtype=torch.bfloat16)
cfg = sp.SelfPlayConfig(n_slots=24, num_simulations=6, use_cuda_graph=False, seed=1)
t = sp.BatchedSelfPlay(inf, cfg).play(30)
print("selfplay", t.stats["games"], len(t))
torch.cuda.synchronize(); print("sanitizer script done")
