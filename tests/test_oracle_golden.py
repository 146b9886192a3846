"""CPU: pins the C oracle (oracle/hz_oracle.c) to golden vectors produced by the unmodified
Python reference (oracle/gen_golden.py).  Integer/byte results must be bit-exact."""

import numpy as np

from tests.conftest import load_golden
from harmonies_alphazero_b200 import packed as pk
from harmonies_alphazero_b200.constants import TILE_TYPES, coordinate_to_index_map

CANON = slice(0, 24)  # boards, piles, hand, bag, meta, scores


def _canon_eq(a, b):
    return np.array_equal(a[:, CANON], b[:, CANON])


def test_engine_traces_explicit_draws(oracle):
    """harmonies_engine.py:145-329 — replay every golden game with its recorded draws."""
    g = load_golden("engine")
    before, after = g["before"], g["after"]
    new, status = oracle.apply(before, g["action"], g["draw"])
    assert (status == 0).all()
    assert _canon_eq(new, after)
    assert np.array_equal(new[:, 27], after[:, 27])  # move counter
    assert np.array_equal(oracle.legal_mask(before), g["legal"])


def test_engine_traces_stream_draws(oracle):
    """stream-mode games: the oracle's own draw source must reproduce the reference run with
    the same patched draw function, including rng event counters."""
    g = load_golden("engine")
    sel = g["game_of"] >= int(g["n_python"])
    new, status = oracle.apply(g["before"][sel], g["action"][sel])
    assert (status == 0).all()
    assert np.array_equal(new[:, :28], g["after"][sel][:, :28])


def test_init_and_full_playouts(oracle):
    """__init__ (:66-79) + the playout policy: final scores/winner of the stream games."""
    g = load_golden("engine")
    npy = int(g["n_python"])
    keys = g["keys"][npy:]
    init = oracle.init_states(len(keys), keys=keys)
    first = g["before"][g["starts"][npy:-1]]
    assert np.array_equal(init[:, :28], first[:, :28])
    final, steps, total = oracle.playout(init, n_threads=3)
    assert np.array_equal(steps, g["lengths"][npy:])
    assert total == int(g["lengths"][npy:].sum())
    f = [pk.unpack_fields(w) for w in final]
    assert [x["final_scores"] for x in f] == g["final_scores"][npy:].tolist()
    assert [x["winner"] for x in f] == g["winner"][npy:].tolist()
    over, oc = oracle.outcome(final)
    assert over.all()
    assert oc.tolist() == [{0: 1, 1: -1, -1: 0}[w] for w in g["winner"][npy:].tolist()]
    # random_actions reproduces the recorded action at every step
    sel = g["game_of"] >= npy
    assert np.array_equal(oracle.random_actions(g["before"][sel]), g["action"][sel])


def test_scoring_and_legal_synthetic(oracle):
    """harmonies_engine.py:357-523 on synthetic boards incl. unreachable stacks."""
    g = load_golden("scoring")
    sc, terms = oracle.score(g["states"])
    assert np.array_equal(terms, g["terms"])
    assert np.array_equal(sc, g["totals"])
    assert np.array_equal(oracle.legal_mask(g["states"]), g["legal"])


def test_encode(oracle):
    """process_game_state.py:15-137 — bit-exact fp32."""
    g = load_golden("encode")
    b, gl = oracle.encode(g["states"])
    assert np.array_equal(b.view(np.uint32), g["board"].view(np.uint32))
    assert np.array_equal(gl.view(np.uint32), g["glob"].view(np.uint32))


def test_canonical_equivalence(oracle):
    """hash partition == the reference's __eq__ classes (harmonies_engine.py:81-118)."""
    g = load_golden("equiv")
    # mode 0: __eq__ classes; mode 1: classes of Python's hash(state), which MCTS.py keys
    # its node dict by (hash(-1) == hash(-2) merges aliased coordinates)
    assert len(set(g["hcls"].tolist())) < len(set(g["cls"].tolist()))
    for mode, cls in ((pk.KEY_EXACT, g["cls"]), (pk.KEY_REFERENCE, g["hcls"])):
        h = oracle.canon_hash(g["states"], mode)
        by_cls, by_hash = {}, {}
        for hi, ci in zip(h.tolist(), cls.tolist()):
            assert by_cls.setdefault(ci, hi) == hi
            assert by_hash.setdefault(hi, ci) == ci
        # python host implementation agrees with C
        idx = list(range(0, len(h), 37))
        for k in idx:
            assert pk.canon_hash(g["states"][k], mode) == int(h[k])


def test_apply_error_codes(oracle):
    """ValueError sites of apply_move (harmonies_engine.py:220,234,242,246,281,296)."""
    g = load_golden("engine")
    b = g["before"]
    phase = (b[:, 22] >> 25) & 7
    choose = b[phase == 0][:1]
    place = b[phase == 1][:1]
    term = g["after"][((g["after"][:, 22] >> 25) & 7) == 4][:1]
    n_piles = int((choose[0, 22] >> 16) & 0xFF)
    for st, a, code in [
        (choose, n_piles, 1),
        (choose, -1, 1),
        (choose, 40, 1),
        (place, 2, 2),
        (place, 143, 3),
        (term, 0, 6),
        (term, 77, 6),
    ]:
        new, status = oracle.apply(st, [a])
        assert status[0] == code, (a, code, status)
        assert np.array_equal(new, st)
    # tile not in hand / illegal stack
    f = pk.unpack_fields(place[0])
    missing = [t for t in range(6) if TILE_TYPES[t] not in f["tiles_in_hand"]][0]
    new, status = oracle.apply(place, [5 + 23 * missing])
    assert status[0] == 4 and np.array_equal(new, place)
    s = pk.pack_fields([{(0, 0): ["water"]}, {}], {"water": 5}, [], 0, ["water", "plant", "wood"], "place_tile_1")
    hexi = coordinate_to_index_map[(0, 0)]
    for t in (0, 1, 2):
        new, status = oracle.apply(s[None], [5 + 23 * t + hexi])
        assert status[0] == 5


def test_weird_states(oracle):
    """states only reachable through initial_state (harmonies_engine.py:67-68): ragged piles,
    leftover hands, nearly empty bags, stray flags — legal moves, results and error sites"""
    g = load_golden("weird")
    assert np.array_equal(oracle.legal_mask(g["states"]), g["legal"])
    new, status = oracle.apply(g["states"], g["action"])
    assert np.array_equal(status, g["status"])
    assert np.array_equal(new[:, :28], g["after"][:, :28])


def test_greedy_agent(oracle):
    """choose_move_greedy (evaluation.py:137-196)"""
    g = load_golden("greedy")
    assert np.array_equal(oracle.greedy_actions(g["states"]), g["action"])


def test_mcts_golden(oracle):
    """MCTS.py:63-441 — visit counts, W, priors, node/edge counts and the chosen move of
    the reference searches (synthetic evaluator, injected noise/uniform)."""
    g = load_golden("mcts")
    for i in range(len(g["sims"])):
        noise = None if g["testing"][i] else g["noise"][i]
        r = oracle.search(
            g["root"][i], g["skey"][i], int(g["sims"][i]), float(g["cpuct"][i]), noise, float(g["eps"][i])
        )
        assert r["rc"] == 0
        assert np.array_equal(r["N"], g["N"][i]), i
        assert np.array_equal(r["W"], g["W"][i]), i
        assert np.array_equal(r["P"].view(np.uint32), g["P"][i].view(np.uint32)), i
        assert r["n_nodes"] == g["n_nodes"][i] and r["n_edges"] == g["n_edges"][i], i
        tot = r["N"].sum()
        if tot:
            pi = r["N"] / tot
            assert np.abs(pi - g["pi"][i]).max() <= 1e-6
        expl = (not g["testing"][i]) and g["move_no"][i] < g["tau0"][i]
        assert oracle.choose(r["N"], float(g["choice_u"][i]), expl) == g["action"][i], i


def test_mcts_python_callback(oracle):
    """the evaluator callback path gives the same tree as the built-in synthetic one"""
    g = load_golden("mcts")
    i = 3
    args = (g["root"][i], g["skey"][i], int(g["sims"][i]), float(g["cpuct"][i]), None, 0.0)
    r = oracle.search(*args, eval_fn=lambda w: pk.fake_eval(pk.canon_hash(w)))
    r2 = oracle.search(*args)
    assert np.array_equal(r["N"], r2["N"]) and np.array_equal(r["W"], r2["W"])


def test_selfplay_worker_golden(oracle):
    """the reference worker's games (tests/golden/selfplay.npz): the oracle's search from each
    recorded state, with the library's default search key, gives the recorded pi; applying the
    first-max-N move leads to the next recorded state; z follows the final outcome"""
    g = load_golden("selfplay")
    S, PI, Z, G = g["states"], g["pi"], g["z"], g["game"]
    sims, cpuct = int(g["sims"]), float(g["cpuct"])
    for game in range(int(G.max()) + 1):
        idx = np.nonzero(G == game)[0]
        for j, i in enumerate(idx):
            w = S[i]
            key = int(w[24]) | (int(w[25]) << 32)
            r = oracle.search(w, pk.rand(key ^ pk.SEARCH_SALT, int(w[27])), sims, cpuct)
            N = r["N"].astype(np.float64)
            assert np.abs(N / N.sum() - PI[i]).max() <= 1e-6
            nxt, st = oracle.apply(w[None], np.array([int(np.argmax(r["N"]))], dtype=np.int16))
            assert st[0] == 0
            if j + 1 < len(idx):
                assert np.array_equal(nxt[0][:28], S[idx[j + 1]][:28])
            else:
                over, oc = oracle.outcome(nxt)
                assert over[0]
                mover = (S[idx] [:, 22] >> 24) & 1
                want = np.where(mover == 0, float(oc[0]), -float(oc[0])) if oc[0] != 0 else np.zeros(len(idx))
                assert np.array_equal(Z[idx], want.astype(np.float32))


def test_oracle_search_with_the_reference_models_priors(oracle):
    """tests/golden/mcts_real.npz: the reference's MCTS.py with the reference's default-size network
    (real fp32 softmax priors).  The C oracle, fed the recorded (priors, value) table in call order,
    reproduces visit counts, W bit for bit, priors, node and edge counts."""
    g = load_golden("mcts_real")
    sims = int(g["sims"])
    for i in range(len(g["root"])):
        calls = [0]

        def ev(w, i=i):
            k = calls[0]
            calls[0] += 1
            return g["table_p"][i, k], float(g["table_v"][i, k])

        noise = None if g["testing"][i] else g["noise"][i]
        r = oracle.search(g["root"][i], g["skey"][i], sims, float(g["cpuct"][i]), noise=noise, eps=float(g["eps"][i]), eval_fn=ev)
        assert calls[0] == int(g["n_eval"][i])
        assert np.array_equal(r["N"], g["N"][i]) and np.array_equal(r["W"], g["W"][i])
        assert np.array_equal(r["P"].view(np.uint32), g["P"][i].view(np.uint32))
        assert (r["n_nodes"], r["n_edges"]) == (int(g["n_nodes"][i]), int(g["n_edges"][i]))
