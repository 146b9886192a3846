"""Can the tree kernels of one half batch run under the convolution tower of the other half?
Two halves (2,048 trees each), the network on a low-priority stream, select / expand+backup on a
high-priority stream, software-pipelined; compared with the serial full-batch step.

    python profiles/overlap_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb, net as hznet, selfplay as sp  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
model = hznet.AlphaZeroNet.from_config(hznet.DEFAULT_MODEL_CONFIG).eval()
inf = hznet.InferenceNet(model, device=dev, dtype=torch.bfloat16)


def driver(n):
    cfg = sp.SelfPlayConfig(n_slots=n, num_simulations=100, use_cuda_graph=False, seed=77)
    d = sp.BatchedSelfPlay(inf, cfg, device=dev)
    st = hb.init_states(n, device=dev, seed=77)
    hb.playout(st, max_steps=8)
    d.groups[0].prepare(st)
    return d, d.groups[0]


def tree_select(d, g):
    g.tree.select(d.cfg.cpuct, g.board, g.glob, dtype=inf.dtype, channels_last=True, pad40=d.pad40)


def net(g):
    inf(g.board, g.glob, out=(g.logits, g.value))


def tree_expand(d, g):
    g.tree.expand_backup(g.logits, g.value, is_logits=True, noise=g.noise, eps=d.cfg.dirichlet_epsilon)


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2_000_000)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


out = {}
# ---- serial, full batch
dF, gF = driver(4096)


def serial_full():
    tree_select(dF, gF); net(gF); tree_expand(dF, gF)


for _ in range(5):
    serial_full()
out["serial_full_us"] = timed(serial_full, 30)

# ---- two halves, serial
dA, gA = driver(2048)
dB, gB = driver(2048)


def serial_halves():
    tree_select(dA, gA); net(gA); tree_expand(dA, gA)
    tree_select(dB, gB); net(gB); tree_expand(dB, gB)


for _ in range(5):
    serial_halves()
out["serial_halves_us"] = timed(serial_halves, 30)

# ---- pipelined: network on a low-priority stream, tree kernels on a high-priority stream
lo_prio, hi_prio = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
s_net = torch.cuda.Stream(device=dev, priority=0)
s_tree = torch.cuda.Stream(device=dev, priority=-1)
main = torch.cuda.current_stream(dev)


def pipelined(reps):
    s_net.wait_stream(main); s_tree.wait_stream(main)
    ev_sel = {}
    ev_net = {}
    with torch.cuda.stream(s_tree):
        tree_select(dA, gA)
        ev_sel["A"] = torch.cuda.Event(); ev_sel["A"].record(s_tree)
        tree_select(dB, gB)
        ev_sel["B"] = torch.cuda.Event(); ev_sel["B"].record(s_tree)
    for _ in range(reps):
        for name, d, g in (("A", dA, gA), ("B", dB, gB)):
            with torch.cuda.stream(s_net):
                s_net.wait_event(ev_sel[name])
                net(g)
                ev_net[name] = torch.cuda.Event(); ev_net[name].record(s_net)
            with torch.cuda.stream(s_tree):
                s_tree.wait_event(ev_net[name])
                tree_expand(d, g)
                tree_select(d, g)
                ev_sel[name] = torch.cuda.Event(); ev_sel[name].record(s_tree)
    main.wait_stream(s_net); main.wait_stream(s_tree)


pipelined(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(2_000_000)
e0.record()
pipelined(30)
e1.record()
torch.cuda.synchronize()
out["pipelined_two_halves_us"] = e0.elapsed_time(e1) * 1e3 / 30
for d in (dF, dA, dB):
    d.check_status()
print(json.dumps(out))
