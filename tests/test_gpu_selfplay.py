"""GPU: the batched self-play driver (replacement of trainer.py:62-134,434-541) and MCTS
parity with a real network: fp32 net outputs captured once and fed to both sides
(SURVEY.md §8d cfg 4 parity sub-run)."""

import numpy as np
import pytest
import torch

from harmonies_alphazero_b200 import packed as pk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from harmonies_alphazero_b200 import batched, net, selfplay, tree

    batched._lib.load()
    return batched, net, selfplay, tree


def _small_net(net, dtype, seed=0):
    torch.manual_seed(seed)
    m = net.AlphaZeroNet.from_config(net.TEST_MODEL_CONFIG)
    for mod in m.modules():   # non-trivial BatchNorm statistics so that folding is exercised
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.2)
    m.eval()
    return m, net.InferenceNet(m, device="cuda", dtype=dtype)


def test_inference_net_matches_module(mods):
    hb, net, _, _ = mods
    m, inf = _small_net(net, torch.float32)
    st = hb.init_states(300, seed=1)
    hb.playout(st, max_steps=23)
    board, glob = hb.encode(st, channels_last=True)
    with torch.no_grad():
        l_ref, v_ref = m.cuda()(board.contiguous(), glob)
    l, v = inf(board, glob)
    assert torch.allclose(l, l_ref, atol=2e-4, rtol=1e-4) and torch.allclose(v, v_ref.view(-1), atol=1e-4)
    # bf16 path: same function at reduced precision
    _, inf16 = _small_net(net, torch.bfloat16)
    b16, g16 = hb.encode(st, dtype=torch.bfloat16, channels_last=True)
    l16, v16 = inf16(b16, g16)
    assert (torch.softmax(l16, 1) - torch.softmax(l_ref, 1)).abs().max() < 0.05
    assert (v16 - v_ref.view(-1)).abs().max() < 0.1


def test_inference_net_matches_reference_model_golden(mods):
    """the reference's own model (state_dict + outputs in tests/golden/net.npz) through the CUDA
    inference path: fp32 within 1e-4 of the reference's logits, bf16 at bf16 precision"""
    from tests.test_abi_and_host import _golden_net

    _, net, _, _ = mods
    g, m = _golden_net()
    b, gl = torch.from_numpy(g["board"]).cuda(), torch.from_numpy(g["glob"]).cuda()
    inf = net.InferenceNet(m, device="cuda", dtype=torch.float32)
    l, v = inf(b.contiguous(memory_format=torch.channels_last), gl)
    assert np.abs(l.cpu().numpy() - g["logits"]).max() < 1e-4 and np.abs(v.cpu().numpy() - g["value"]).max() < 1e-4
    inf16 = net.InferenceNet(m, device="cuda", dtype=torch.bfloat16)
    l16, v16 = inf16(b.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), gl.to(torch.bfloat16))
    assert np.abs(torch.softmax(l16, 1).cpu().numpy() - g["probs"]).max() < 0.03
    assert np.abs(v16.cpu().numpy() - g["value"]).max() < 0.08


@pytest.mark.parametrize("cfg_name", ["TEST_MODEL_CONFIG", "DEFAULT_MODEL_CONFIG"])
def test_fused_heads_kernel_matches_torch_heads(mods, cfg_name):
    """hz_net_heads (model.py:340-355 in one kernel) vs the same tail computed by torch in
    fp32 from the same bf16 tower output; tolerance: fp32 accumulation-order noise."""
    hb, net, _, _ = mods
    torch.manual_seed(1)
    m = net.AlphaZeroNet.from_config(getattr(net, cfg_name))
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.2)
    m.eval()
    inf = net.InferenceNet(m, device="cuda", dtype=torch.bfloat16)
    assert inf.heads is not None
    C = inf.heads["C"]
    for B in (1, 6, 7, 8, 13, 14, 15, 300, 2077):
        x = (torch.randn((B, C, 5, 7), device="cuda") * 0.7).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        glob = torch.rand((B, 42), device="cuda").to(torch.bfloat16)
        logits, value = inf._fused_heads(x, glob, None)
        mm = m.cuda().float()
        xf, gf = x.float().contiguous(), glob.float()
        with torch.no_grad():
            p = torch.relu(mm.policy_bn(mm.policy_conv(xf))).flatten(1)
            want_l = mm.policy_fc(torch.cat((p, gf), 1))
            v = torch.relu(mm.value_bn(mm.value_conv(xf))).flatten(1)
            want_v = torch.tanh(mm.value_fc2(torch.relu(mm.value_fc1(torch.cat((v, gf), 1))))).view(-1)
        assert torch.allclose(logits, want_l, atol=2e-4, rtol=1e-4), (B, (logits - want_l).abs().max())
        assert torch.allclose(value, want_v, atol=1e-4), (B, (value - want_v).abs().max())
    # the whole forward with fused heads agrees with the torch-head path at bf16 precision
    st = hb.init_states(64, seed=2)
    hb.playout(st, max_steps=20)
    b16, g16 = hb.encode(st, dtype=torch.bfloat16, channels_last=True)
    l1, v1 = inf(b16, g16)
    inf.use_fused_heads = False
    l2, v2 = inf(b16, g16)
    assert (torch.softmax(l1, 1) - torch.softmax(l2, 1)).abs().max() < 0.02 and (v1 - v2).abs().max() < 0.05


def test_search_with_real_net_matches_oracle(mods, oracle):
    """64 games x first 8 moves x 24 simulations, fp32 network, testing=True: the GPU tree
    and the reference-semantics oracle see identical (P, v) per leaf and must produce
    identical visit counts (tolerance: none) and pi within 1e-6."""
    hb, net, _, tr = mods
    _, inf = _small_net(net, torch.float32, seed=3)
    n, sims, cpuct = 64, 24, 2.0
    states = hb.init_states(n, seed=42)
    tree = tr.BatchedMCTS(n, sims)
    board = torch.empty((n, 38, 5, 7), device="cuda").contiguous(memory_format=torch.channels_last)
    glob = torch.empty((n, 42), device="cuda")
    leaf = torch.empty((n, 32), dtype=torch.int32, device="cuda")
    for move in range(8):
        roots = states.cpu().numpy().view(np.uint32).copy()
        skeys = np.array([pk.rand(int(k), move) for k in range(1000, 1000 + n)], dtype=np.uint64)
        tree.reset(states, tr.search_keys_tensor(skeys))
        seen = [dict() for _ in range(n)]
        for s in range(sims):
            tree.select(cpuct, board, glob, leaf, channels_last=True)
            logits, value = inf(board, glob)
            probs = torch.softmax(logits, dim=1).contiguous()       # model.py:104
            lw = leaf.cpu().numpy().view(np.uint32)
            pn, vn = probs.cpu().numpy(), value.cpu().numpy()
            for i in range(n):
                seen[i][pk.canon_hash(lw[i])] = (pn[i], float(vn[i]))
            tree.expand_backup(probs, value)
        tree.check_status()
        N = tree.root_edges()[0].cpu().numpy()
        visits, pi = (x.cpu().numpy() for x in tree.root_policy())
        for i in range(n):
            r = oracle.search(roots[i], skeys[i], sims, cpuct, eval_fn=lambda w, d=seen[i]: d[pk.canon_hash(w)])
            assert np.array_equal(N[i], r["N"]), (move, i)
            tot = r["N"].sum()
            assert np.abs(pi[i] - r["N"] / max(tot, 1)).max() <= 1e-6
        states_before = states.clone()
        hb.apply(states, tree.choose())
        assert not torch.equal(states, states_before)


def test_selfplay_games_and_example_contract(mods, oracle):
    hb, net, sp, _ = mods
    _, inf = _small_net(net, torch.bfloat16, seed=5)
    cfg = sp.SelfPlayConfig(n_slots=96, num_simulations=8, turns_until_tau0=6, seed=11)
    torch.manual_seed(0)
    driver = sp.BatchedSelfPlay(inf, cfg)
    traj = driver.play(150)
    st = traj.stats
    assert st["games"] == 150 and st["examples"] == len(traj) and st["sims"] > 0
    gid = traj.game_id.cpu().numpy(); mv = traj.move_no.cpu().numpy()
    S = traj.states.cpu().numpy().view(np.uint32); V = traj.visits.cpu().numpy(); z = traj.z.cpu().numpy()
    assert sorted(set(gid.tolist())) == list(range(150))
    assert (V.sum(axis=1) == cfg.num_simulations - 1).all()          # sum N = S-1 (MCTS.py:355-381)
    pi = traj.pi().cpu().numpy()
    assert np.allclose(pi.sum(axis=1), 1.0, atol=1e-6)
    # visits only on legal actions of the recorded state
    legal = oracle.legal_mask(S)
    for i in range(0, len(S), 37):
        acts = set(pk.mask_to_actions(legal[i]))
        assert set(np.nonzero(V[i])[0].tolist()) <= acts
    # per game: consecutive move numbers from 0, z = +-outcome by mover, successive states are
    # one legal move apart (boards/hand/piles evolve legally)
    for g in range(0, 150, 7):
        idx = np.nonzero(gid == g)[0]
        idx = idx[np.argsort(mv[idx])]
        assert mv[idx].tolist() == list(range(len(idx)))
        movers = (S[idx, 22] >> 24) & 1
        zz = z[idx]
        assert set(np.unique(np.abs(zz)).tolist()) <= {0.0, 1.0}
        if np.abs(zz).max() > 0:
            assert (zz[movers == 0] == zz[movers == 0][0]).all() and (zz[movers == 1] == -zz[movers == 0][0]).all()
        for a, b in zip(idx[:-1], idx[1:]):
            cand_actions = pk.mask_to_actions(legal[a])
            nxt, stt = oracle.apply(np.repeat(S[a][None], len(cand_actions), 0), np.array(cand_actions, dtype=np.int16))
            same_board = [(nxt[k, :18] == S[b, :18]).all() and (nxt[k, 20] >> 16) == (S[b, 20] >> 16) for k in range(len(cand_actions))]
            assert any(same_board)
    # reference example contract (trainer.py:531-538)
    ex = traj.to_reference_examples()
    assert len(ex) == len(traj)
    b, gl, p, zz = ex[0]
    assert b.shape == (38, 5, 7) and gl.shape == (42,) and p.shape == (143,) and zz.shape == (1,)
    assert b.dtype == gl.dtype == p.dtype == zz.dtype == torch.float32 and not b.is_cuda
    ob, og = oracle.encode(S[:1])
    assert np.array_equal(b.numpy(), ob[0]) and np.array_equal(gl.numpy(), og[0])
    # buffer.save_buffer pickles the deque of examples: each tensor must own its storage
    import pickle

    assert len(pickle.dumps(ex[:20], pickle.HIGHEST_PROTOCOL)) < 20 * 8000
    assert b.untyped_storage().nbytes() == 38 * 5 * 7 * 4
    tail = traj.to_reference_examples(last=7)
    assert len(tail) == 7 and all(torch.equal(x, y) for x, y in zip(tail[-1], ex[-1])) and torch.equal(tail[0][2], ex[-7][2])


def test_selfplay_reproduces_the_reference_worker(mods):
    """f1: tests/golden/selfplay.npz was written by the reference's own self_play_worker
    (trainer.py:434-541) playing whole games with the fake evaluator and the library's draw
    streams.  BatchedSelfPlay.play with the same evaluator must produce the same examples:
    identical states before every search (bit-exact), pi within 1e-6, identical z."""
    from tests.conftest import load_golden

    hb, net, sp, _ = mods
    g = load_golden("selfplay")
    n_games = int(g["game"].max()) + 1
    cfg = sp.SelfPlayConfig(n_slots=n_games, num_simulations=int(g["sims"]), cpuct=float(g["cpuct"]), testing=True,
                            seed=int(g["seed"]), use_cuda_graph=False)
    traj = sp.BatchedSelfPlay(sp.SyntheticEvaluator(), cfg).play(n_games)
    assert len(traj) == len(g["z"])
    S, P, Z = traj.states.cpu().numpy().view(np.uint32), traj.pi().cpu().numpy(), traj.z.cpu().numpy()
    G, M = traj.game_id.cpu().numpy(), traj.move_no.cpu().numpy()
    for game in range(n_games):
        idx = np.nonzero(G == game)[0]
        idx = idx[np.argsort(M[idx], kind="stable")]
        ref = np.nonzero(g["game"] == game)[0]
        assert len(idx) == len(ref)
        assert np.array_equal(S[idx][:, :28], g["states"][ref][:, :28])
        assert np.abs(P[idx] - g["pi"][ref]).max() <= 1e-6
        assert np.array_equal(Z[idx], g["z"][ref])
    board, glob = hb.encode(torch.from_numpy(g["states"][:4].view(np.int32)).cuda())     # the worker's state tensors
    assert np.array_equal(board.cpu().numpy(), g["board0"]) and np.array_equal(glob.cpu().numpy(), g["glob0"])


def test_replay_ring_matches_reference_deque_semantics(mods, oracle):
    """f2: packed GPU ring == deque(maxlen) of the reference tuples (buffer.py, trainer.py:127)"""
    hb, net, sp, _ = mods
    from collections import deque

    from harmonies_alphazero_b200.replay import ReplayRing

    _, inf = _small_net(net, torch.bfloat16, seed=4)
    cfg = sp.SelfPlayConfig(n_slots=32, num_simulations=5, testing=True, seed=21)
    ring, ref = ReplayRing(500), deque(maxlen=500)
    for it in range(3):                                   # 3 x 6 games ~ 1100 examples > capacity
        traj = sp.BatchedSelfPlay(inf, sp.SelfPlayConfig(n_slots=32, num_simulations=5, testing=True, seed=21 + it)).play(6)
        ring.extend(traj)
        ref.extend(traj.to_reference_examples())
    assert len(ring) == len(ref) == 500
    got = ring.to_reference_buffer()
    assert got.maxlen == 500 and len(got) == 500
    import pickle

    assert len(pickle.dumps(got, pickle.HIGHEST_PROTOCOL)) < 500 * 8000      # buffer.save_buffer stays ~6 KB per example
    for k in (0, 1, 250, 499):
        for a, b in zip(got[k], ref[k]):
            assert torch.equal(a, b)
    board, glob, pi, z = ring.sample(64)
    assert board.shape == (64, 38, 5, 7) and glob.shape == (64, 42) and pi.shape == (64, 143) and z.shape == (64, 1)
    assert board.is_cuda and torch.allclose(pi.sum(1), torch.ones(64, device="cuda"), atol=1e-6)
    seen = sum(b.shape[0] for b, _, _, _ in ring.epoch(128))
    assert seen == 500
    ring2 = ReplayRing(500)
    ring2.load_state_dict(ring.state_dict())
    for a, b in zip(ring2.to_reference_buffer()[123], got[123]):
        assert torch.equal(a, b)
    del cfg


def test_arena_match(mods):
    """candidate-vs-best evaluation (trainer.py:293-431) batched over all eval games"""
    hb, net, _, _ = mods
    from harmonies_alphazero_b200 import arena

    _, cand = _small_net(net, torch.bfloat16, seed=1)
    _, best = _small_net(net, torch.bfloat16, seed=2)
    cfg = {"num_simulations": 6, "cpuct": 2, "dirichlet_alpha": 0.1, "dirichlet_epsilon": 0,
           "turns_until_tau0": 0, "action_size": 143, "testing": True}
    r1 = arena.play_match(cand, best, 9, cfg, seed=5)
    assert r1["candidate_wins"] + r1["best_wins"] + r1["draws"] == 9 and 0.0 <= r1["win_rate"] <= 1.0
    assert arena.play_match(cand, best, 9, cfg, seed=5) == r1          # deterministic in testing mode
    # swapping the roles mirrors the tally
    r2 = arena.play_match(best, cand, 9, cfg, seed=5)
    assert r2["games"] == 9 and r2["candidate_wins"] + r2["best_wins"] + r2["draws"] == 9
    # run_tournament (evaluation.py:7-65): network vs the 1-ply greedy agent; an untrained net
    # with 6 simulations loses to greedy play; greedy vs greedy is a legal, complete match
    r3 = arena.play_match(cand, arena.GREEDY, 10, cfg, seed=7)
    assert r3["candidate_wins"] + r3["best_wins"] + r3["draws"] == 10 and r3["best_wins"] >= 8
    r4 = arena.play_match(arena.GREEDY, arena.GREEDY, 6, cfg, seed=7)
    assert r4["candidate_wins"] + r4["best_wins"] + r4["draws"] == 6
    # virtual-loss searches (3 simulations in flight per tree): a complete, deterministic match
    r5 = arena.play_match(cand, best, 6, cfg, seed=5, leaves=3)
    assert r5["candidate_wins"] + r5["best_wins"] + r5["draws"] == 6
    assert arena.play_match(cand, best, 6, cfg, seed=5, leaves=3) == r5


def test_selfplay_with_virtual_loss(mods):
    """leaves_per_step > 1: whole games complete, sum N = S - K (the K descents of the first
    step all end at the unexpanded root), examples keep the reference contract"""
    hb, net, sp, _ = mods
    _, inf = _small_net(net, torch.bfloat16, seed=6)
    cfg = sp.SelfPlayConfig(n_slots=48, num_simulations=12, leaves_per_step=4, seed=13)
    torch.manual_seed(1)
    traj = sp.BatchedSelfPlay(inf, cfg).play(60)
    assert traj.stats["games"] == 60
    assert (traj.visits.sum(dim=1) == 12 - 4).all()
    assert torch.allclose(traj.pi().sum(dim=1), torch.ones(len(traj), device="cuda"), atol=1e-6)
    with pytest.raises(ValueError):
        sp.BatchedSelfPlay(inf, sp.SelfPlayConfig(n_slots=8, num_simulations=10, leaves_per_step=4))


def test_selfplay_is_independent_of_slot_count_in_testing_mode(mods):
    """deterministic mode: the same game ids give the same trajectories whatever the batch
    shape (keys are per game id, search keys per (game, move))."""
    hb, net, sp, _ = mods

    class ElementwiseNet:   # row-wise elementwise function: bitwise independent of the batch shape
        dtype, device = torch.float32, torch.device("cuda")

        def __call__(self, board, glob):
            a = torch.arange(143, device=glob.device, dtype=torch.float32)
            x = glob[:, 36:42].repeat(1, 24)[:, :143] * 3.0 + glob[:, 30:36].repeat(1, 24)[:, :143]
            logits = torch.sin(a * 0.37 + x * 5.0) + board[:, 36, 2, 3].unsqueeze(1) * torch.cos(a)
            return logits, torch.tanh(glob[:, 36] * 3.0 - glob[:, 41] * 2.0 - glob[:, 33])

    inf = ElementwiseNet()
    out = []
    for slots in (16, 40):
        cfg = sp.SelfPlayConfig(n_slots=slots, num_simulations=6, testing=True, seed=3, use_cuda_graph=(slots == 40))
        t = sp.BatchedSelfPlay(inf, cfg).play(24)
        order = torch.argsort(t.game_id * 1000 + t.move_no.to(torch.int64))
        out.append((t.states[order].cpu(), t.visits[order].cpu(), t.z[order].cpu()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
