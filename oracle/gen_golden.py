"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py).  Run here (the dev container);
the fixtures are committed because the reference cannot travel to the GPU box.

    python oracle/gen_golden.py            # writes tests/golden/{engine,scoring,encode,equiv,mcts}.npz

Every array is produced by reference code (harmonies_engine.py, process_game_state.py,
MCTS.py); this script only chooses inputs, injects the deterministic draw source and
serialises states with harmonies_alphazero_b200.packed (format conversion, no game logic).
"""

import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_harness as rh  # noqa: E402
from harmonies_alphazero_b200 import packed as pk  # noqa: E402
from harmonies_alphazero_b200.constants import TILE_TYPES, sorted_coords  # noqa: E402

OUT = os.environ.get("HZ_GOLDEN_OUT", os.path.join(os.path.dirname(HERE), "tests", "golden"))


def legal_mask(state, gai):
    return pk.actions_to_mask([gai(m) for m in state.get_legal_moves()])


def playout_pick(key, moves, n_legal):
    return ((pk.rand(key ^ pk.PLAYOUT_SALT, moves) >> 32) * n_legal) >> 32


# --------------------------------------------------------------------------------------
def gen_engine(n_python=24, n_stream=40):
    ref = rh.load_reference()
    gai = ref["pgs"].get_action_index
    before, after, legal, action, draw, game_of = [], [], [], [], [], []
    starts, finals, winners, keys, lengths = [0], [], [], [], []
    g = 0
    for mode in ["python"] * n_python + ["stream"] * n_stream:
        if mode == "python":
            key = 0
            state = rh.new_game_python(seed=g)
            picker = random.Random(10_000 + g)
        else:
            key = pk.rand(0xB200, g)
            state = rh.new_game_stream(key)
        moves = 0
        while not state.is_game_over():
            lm = state.get_legal_moves()
            assert lm, "reference produced a stuck state"
            if mode == "python":
                mv = lm[picker.randrange(len(lm))]
                ev = 0
            else:
                mv = lm[playout_pick(key, moves, len(lm))]
                ev = rh.ctx.event
            before.append(pk.pack_state(state, rng_key=key, rng_event=ev, moves=moves))
            legal.append(legal_mask(state, gai))
            action.append(gai(mv))
            rh.ctx.recorded = []
            state = state.apply_move(mv)
            rec, rh.ctx.recorded = rh.ctx.recorded, None
            assert len(rec) <= 1
            draw.append(rec[0] if rec else pk.NO_DRAW)
            moves += 1
            ev = rh.ctx.event if mode == "stream" else 0
            after.append(pk.pack_state(state, rng_key=key, rng_event=ev, moves=moves))
            game_of.append(g)
        starts.append(len(action))
        finals.append(list(state.final_scores))
        winners.append(state.winner)
        keys.append(key)
        lengths.append(moves)
        g += 1
    np.savez_compressed(
        os.path.join(OUT, "engine.npz"),
        before=np.array(before, dtype=np.uint32),
        after=np.array(after, dtype=np.uint32),
        legal=np.array(legal, dtype=np.uint32),
        action=np.array(action, dtype=np.int16),
        draw=np.array(draw, dtype=np.uint16),
        game_of=np.array(game_of, dtype=np.int32),
        starts=np.array(starts, dtype=np.int64),
        final_scores=np.array(finals, dtype=np.int16),
        winner=np.array(winners, dtype=np.int8),
        keys=np.array(keys, dtype=np.uint64),
        lengths=np.array(lengths, dtype=np.int32),
        n_python=np.int32(n_python),
    )
    print(f"engine: {g} games, {len(action)} steps, lengths {min(lengths)}..{max(lengths)}")
    return np.array(before, dtype=np.uint32), np.array(after, dtype=np.uint32)


# --------------------------------------------------------------------------------------
def synth_state(G, rng, kind):
    """Synthetic positions (BASELINE.json configs[2]): full boards with stacks of height
    ~{1,1,2,3} and uniform types (incl. unreachable stacks); 'sparse' leaves holes and
    'watery'/'fieldy'/'stony' skew the type distribution to grow large components."""
    weights = {
        "full": [1, 1, 1, 1, 1, 1],
        "sparse": [1, 1, 1, 1, 1, 1],
        "watery": [6, 1, 1, 1, 1, 1],
        "fieldy": [1, 1, 1, 1, 1, 6],
        "stony": [1, 1, 1, 6, 3, 1],
    }[kind]
    boards = []
    for p in (0, 1):
        b = {}
        occ = 1.0 if kind == "full" else rng.choice([0.35, 0.6, 0.85, 1.0])
        for c in sorted_coords:
            if rng.random() > occ:
                continue
            h = rng.choice([1, 1, 2, 3])
            b[c] = [rng.choices(TILE_TYPES, weights)[0] for _ in range(h)]
        boards.append(b)
    hand = [rng.choice(TILE_TYPES) for _ in range(3)]
    piles = [[rng.choice(TILE_TYPES) for _ in range(3)] for _ in range(rng.choice([4, 4, 3, 0]))]
    bag = {t: rng.randrange(0, 20) for t in ["water", "plant", "wood", "stone", "field", "building"]}
    return G(
        initial_state={
            "player_boards": boards,
            "tile_bag": bag,
            "available_piles": piles,
            "current_player": rng.randrange(2),
            "tiles_in_hand": hand,
            "turn_phase": rng.choice(["place_tile_1", "place_tile_2", "place_tile_3"]),
            "game_over": False,
            "winner": None,
            "final_scores": [0, 0],
        }
    )


def gen_scoring(n=6000):
    ref = rh.load_reference()
    G, gai = ref["G"], ref["pgs"].get_action_index
    rng = random.Random(20240)
    kinds = ["full"] * 3 + ["sparse", "watery", "fieldy", "stony"]
    states, terms, totals, legal = [], [], [], []
    for i in range(n):
        s = synth_state(G, rng, kinds[i % len(kinds)])
        if i % 3 == 0:  # hands with repeated / single types exercise the distinct-type rule
            s.tiles_in_hand = s.tiles_in_hand[: rng.choice([1, 2, 3])]
        states.append(pk.pack_state(s))
        tt = []
        for p in (0, 1):
            b = s.player_boards[p]
            tt.append(
                [
                    s._score_grass(b, p),
                    s._score_mountains(b, p),
                    s._score_fields(b, p),
                    s._score_buildings(b, p),
                    s._score_water(b, p),
                ]
            )
        terms.append(tt)
        totals.append([s.calculate_score_for_player(0), s.calculate_score_for_player(1)])
        legal.append(legal_mask(s, gai))
    np.savez_compressed(
        os.path.join(OUT, "scoring.npz"),
        states=np.array(states, dtype=np.uint32),
        terms=np.array(terms, dtype=np.int16),
        totals=np.array(totals, dtype=np.int16),
        legal=np.array(legal, dtype=np.uint32),
    )
    t = np.array(terms)
    print(f"scoring: {n} positions, max terms {t.max(axis=(0, 1))}, max total {np.max(totals)}")
    return np.array(states, dtype=np.uint32)


# --------------------------------------------------------------------------------------
def gen_encode(trace_states, synth_states, n_trace=400, n_synth=100):
    ref = rh.load_reference()
    G, cst = ref["G"], ref["pgs"].create_state_tensors
    rng = np.random.default_rng(7)
    pick = list(rng.choice(len(trace_states), n_trace, replace=False))
    # make sure terminal states (phase game_over) are in
    term = [i for i in range(len(trace_states)) if ((trace_states[i][22] >> 25) & 7) == 4][:24]
    sel = [trace_states[i] for i in pick + term] + list(
        synth_states[rng.choice(len(synth_states), n_synth, replace=False)]
    )
    boards, globs = [], []
    for w in sel:
        f = pk.unpack_fields(w)
        for k in ("rng_key", "rng_event", "moves"):
            f.pop(k)
        b, gl = cst(G(initial_state=f))
        boards.append(b.numpy())
        globs.append(gl.numpy())
    np.savez_compressed(
        os.path.join(OUT, "encode.npz"),
        states=np.array(sel, dtype=np.uint32),
        board=np.array(boards, dtype=np.float32),
        glob=np.array(globs, dtype=np.float32),
    )
    print(f"encode: {len(sel)} states")


# --------------------------------------------------------------------------------------
def gen_equiv(n_roots=12, depth=3, cap=6000, n_sparse=3000):
    """States with the reference's own equivalence classes, two ways:
    cls  = classes of __eq__ (harmonies_engine.py:115-118), i.e. tuple equality;
    hcls = classes of hash(state) (harmonies_engine.py:112-113) — what MCTS.py keys its node
           dict by (MCTS.py:14,177,185).  They differ because hash(-1) == hash(-2)."""
    ref = rh.load_reference()
    G = ref["G"]
    states, cls, hcls = [], [], []
    seen, hseen = {}, {}
    rng = random.Random(99)

    def add(c):
        states.append(pk.pack_state(c))
        cls.append(seen.setdefault(c, len(seen)))
        hcls.append(hseen.setdefault(hash(c), len(hseen)))

    for r in range(n_roots):
        key = pk.rand(0xE9, r)
        s = rh.new_game_stream(key)
        for _ in range(rng.choice([1, 5, 9, 21])):
            lm = s.get_legal_moves()
            s = s.apply_move(lm[rng.randrange(len(lm))])
        rh.ctx.mode, rh.ctx.key, rh.ctx.sim = "tree", key ^ 0x77, 0
        frontier = [s]
        for d in range(depth):
            nxt = []
            for st in frontier:
                lm = st.get_legal_moves()
                rng.shuffle(lm)
                for mv in lm[:12]:
                    rh.ctx.sim = 0  # same (sim, action) -> same draw: transpositions survive
                    c = st.apply_move(mv)
                    add(c)
                    nxt.append(c)
            rng.shuffle(nxt)
            frontier = nxt[:40]
        if len(states) > cap:
            break
    # sparse boards drawn from a small pool of stacks, so that alias collisions
    # (coordinates differing by -1 vs -2) and true duplicates are both frequent
    stacks = [["water"], ["wood"], ["wood", "plant"], ["stone", "stone"], ["field"]]
    for i in range(n_sparse):
        boards = []
        for p in (0, 1):
            k = rng.choice([0, 1, 1, 2, 2, 3, 4, 6])
            b = {c: list(rng.choice(stacks[: rng.choice([1, 2, 5])])) for c in rng.sample(sorted_coords, k)}
            boards.append(b)
        add(
            G(
                initial_state={
                    "player_boards": boards,
                    "tile_bag": {"water": 3, "plant": 3, "wood": 3, "stone": 3, "field": 3, "building": 3},
                    "available_piles": [["water", "wood", "field"]],
                    "current_player": i % 2,
                    "tiles_in_hand": ["plant"],
                    "turn_phase": "place_tile_3",
                    "game_over": False,
                    "winner": None,
                    "final_scores": [0, 0],
                }
            )
        )
    np.savez_compressed(
        os.path.join(OUT, "equiv.npz"),
        states=np.array(states, dtype=np.uint32),
        cls=np.array(cls, dtype=np.int32),
        hcls=np.array(hcls, dtype=np.int32),
    )
    print(f"equiv: {len(states)} states, {len(seen)} __eq__ classes, {len(hseen)} hash() classes")


# --------------------------------------------------------------------------------------
def gen_mcts():
    ref = rh.load_reference()
    gai = ref["pgs"].get_action_index
    rng = random.Random(4242)
    nrng = np.random.default_rng(4242)
    cases = []
    # roots: positions along stream-mode playout games
    for g in range(10):
        key = pk.rand(0x3C75, g)
        s = rh.new_game_stream(key)
        traj = [(s, 0, rh.ctx.event)]
        moves = 0
        while not s.is_game_over():
            lm = s.get_legal_moves()
            s = s.apply_move(lm[playout_pick(key, moves, len(lm))])
            moves += 1
            traj.append((s, moves, rh.ctx.event))
        L = moves
        wanted = sorted({0, 1, 2, 3, 4, 9, 22, 38, L - 9, L - 6, L - 4, L - 3, L - 2, L - 1} & set(range(L)))
        for m in rng.sample(wanted, 6 if g else len(wanted)):
            cases.append((traj[m][0], key, m, traj[m][2]))
        if g == 0:
            cases.append((traj[L][0], key, L, traj[L][2]))  # terminal root
    # non-standard roots (tests/golden/weird.npz): stuck positions, ragged piles, nearly empty
    # bags -> trees with dead-end leaves, bag-empty endings and multi-pile replenishes
    wpath = os.path.join(OUT, "weird.npz")
    if os.path.exists(wpath):
        ws = np.load(wpath)["states"]
        for i in range(0, len(ws), len(ws) // 14):
            f = pk.unpack_fields(ws[i])
            key, ev, mv = f.pop("rng_key"), f.pop("rng_event"), f.pop("moves")
            cases.append((ref["G"](initial_state=f), key, mv, ev))
    out = {k: [] for k in "root skey sims cpuct testing eps noise choice_u move_no tau0 N W P pi action n_nodes n_edges".split()}
    t0 = time.time()
    for i, (state, key, m, ev) in enumerate(cases):
        sims = [16, 40, 100, 64, 200][i % 5] if i % 11 else 320
        variant = i % 4
        cfg = {
            "num_simulations": sims,
            "cpuct": [2, 1.0, 2, 1.25][variant],
            "dirichlet_alpha": 0.4,
            "dirichlet_epsilon": [0.25, 0.0, 0.25, 0.4][variant],
            "fpu_value": 0.25,
            "turns_until_tau0": [15, 0, 15, 200][variant],
            "action_size": 143,
            "testing": variant == 1,
        }
        noise = nrng.gamma(0.4, size=143).astype(np.float32) + np.float32(1e-6)
        u = float(np.float32(nrng.random()))
        skey = pk.rand(key ^ 0x5EA7C4, m)
        use_noise = not cfg["testing"]
        mv, pi, info = rh.run_search(
            state, skey, cfg, m, noise=noise if use_noise else None, choice_u=u
        )
        out["root"].append(pk.pack_state(state, rng_key=key, rng_event=ev, moves=m))
        out["skey"].append(skey)
        out["sims"].append(sims)
        out["cpuct"].append(float(cfg["cpuct"]))
        out["testing"].append(int(cfg["testing"]))
        out["eps"].append(cfg["dirichlet_epsilon"])
        out["noise"].append(noise)
        out["choice_u"].append(u)
        out["move_no"].append(m)
        out["tau0"].append(cfg["turns_until_tau0"])
        out["N"].append(info["N"])
        out["W"].append(info["W"])
        out["P"].append(info["P"])
        out["pi"].append(pi)
        out["action"].append(-1 if mv is None else gai(mv))
        out["n_nodes"].append(info["n_nodes"])
        out["n_edges"].append(info["n_edges"])
    dt = time.time() - t0
    np.savez_compressed(
        os.path.join(OUT, "mcts.npz"),
        root=np.array(out["root"], dtype=np.uint32),
        skey=np.array(out["skey"], dtype=np.uint64),
        sims=np.array(out["sims"], dtype=np.int32),
        cpuct=np.array(out["cpuct"], dtype=np.float64),
        testing=np.array(out["testing"], dtype=np.uint8),
        eps=np.array(out["eps"], dtype=np.float64),
        noise=np.array(out["noise"], dtype=np.float32),
        choice_u=np.array(out["choice_u"], dtype=np.float32),
        move_no=np.array(out["move_no"], dtype=np.int32),
        tau0=np.array(out["tau0"], dtype=np.int32),
        N=np.array(out["N"], dtype=np.int32),
        W=np.array(out["W"], dtype=np.float64),
        P=np.array(out["P"], dtype=np.float32),
        pi=np.array(out["pi"], dtype=np.float64),
        action=np.array(out["action"], dtype=np.int16),
        n_nodes=np.array(out["n_nodes"], dtype=np.int32),
        n_edges=np.array(out["n_edges"], dtype=np.int32),
    )
    tot = int(np.sum(out["sims"]))
    print(f"mcts: {len(cases)} searches, {tot} sims in {dt:.1f}s ({tot / dt:.0f} sims/s reference+fake eval)")


# --------------------------------------------------------------------------------------
_ERR = [("Invalid pile index", 1), ("Invalid move format", 2), ("Invalid coordinate", 3),
        ("not found in hand", 4), ("Cannot place", 5), ("Invalid turn phase", 6)]


def gen_weird(n=1500, per_state=8):
    """States the standard game never reaches but the reference code accepts (via
    initial_state, harmonies_engine.py:67-68): 0-5 piles of 1-3 tiles, hands of 0-3 tiles in
    any phase, nearly empty bags (partial piles, several piles drawn in one replenish, bag-empty
    end trigger), arbitrary stacks, stray `game_over` flags.  For each state: the reference's
    legal moves and, for a handful of actions (legal or not), the resulting state or the
    ValueError site — with the library's stream draws (key/event of the state)."""
    ref = rh.load_reference()
    G, gai = ref["G"], ref["pgs"].get_action_index
    rng = random.Random(777)
    phases = ["choose_pile", "place_tile_1", "place_tile_2", "place_tile_3", "game_over"]
    states, legal, actions, status, after = [], [], [], [], []
    for i in range(n):
        boards = []
        for p in (0, 1):
            occ = rng.choice([0.0, 0.3, 0.7, 0.9, 1.0])
            b = {}
            for c in sorted_coords:
                if rng.random() < occ:
                    b[c] = [rng.choice(TILE_TYPES) for _ in range(rng.choice([1, 1, 2, 3]))]
            boards.append(b)
        phase = rng.choice(phases)
        over = phase == "game_over"
        fields = {
            "player_boards": boards,
            "tile_bag": {t: rng.choice([0, 0, 1, 2, 4]) for t in ["water", "plant", "wood", "stone", "field", "building"]},
            "available_piles": [[rng.choice(TILE_TYPES) for _ in range(rng.choice([1, 2, 3, 3]))] for _ in range(rng.randrange(6))],
            "current_player": rng.randrange(2),
            "tiles_in_hand": [rng.choice(TILE_TYPES) for _ in range(rng.randrange(4))],
            "turn_phase": phase,
            "game_over": over or rng.random() < 0.3,
            "winner": rng.choice([0, 1, -1]) if over else None,
            "final_scores": [rng.randrange(40), rng.randrange(40)] if over else [0, 0],
        }
        key, event, moves = pk.rand(0xC0DE, i), rng.randrange(1, 30), rng.randrange(100)
        base = G(initial_state=fields)
        w = pk.pack_state(base, rng_key=key, rng_event=event, moves=moves)
        lm = base.get_legal_moves()
        legal_idx = [gai(m) for m in lm]
        cand = set(rng.sample(legal_idx, min(len(legal_idx), per_state // 2)))
        while len(cand) < per_state:
            cand.add(rng.randrange(0, 143))
        for a in sorted(cand):
            rh.ctx.mode, rh.ctx.key, rh.ctx.event, rh.ctx.k = "stream", key, event, 0
            try:
                nxt = base.apply_move(pk.action_to_move(a))
                st = 0
                aw = pk.pack_state(nxt, rng_key=key, rng_event=rh.ctx.event, moves=moves + 1)
            except ValueError as e:
                st = next(code for msg, code in _ERR if msg in str(e))
                aw = w
            states.append(w); legal.append(pk.actions_to_mask(legal_idx)); actions.append(a); status.append(st); after.append(aw)
    rh.ctx.mode = "python"
    np.savez_compressed(
        os.path.join(OUT, "weird.npz"),
        states=np.array(states, dtype=np.uint32), legal=np.array(legal, dtype=np.uint32),
        action=np.array(actions, dtype=np.int16), status=np.array(status, dtype=np.uint8),
        after=np.array(after, dtype=np.uint32),
    )
    st = np.array(status)
    print(f"weird: {len(states)} (state, action) pairs, status histogram {np.bincount(st, minlength=7).tolist()}")


def gen_greedy(n_trace=700, n_weird=300):
    """choose_move_greedy (evaluation.py:137-196) on trace states and non-standard states."""
    import contextlib
    import io

    ref = rh.load_reference()
    sys.path.insert(0, rh.REF_ROOT)
    try:
        import evaluation as ev
    finally:
        sys.path.remove(rh.REF_ROOT)
    G, gai = ref["G"], ref["pgs"].get_action_index
    rng = np.random.default_rng(11)
    tr = np.load(os.path.join(OUT, "engine.npz"))["before"]
    wd = np.load(os.path.join(OUT, "weird.npz"))["states"]
    sel = np.concatenate([tr[rng.choice(len(tr), n_trace, replace=False)], wd[rng.choice(len(wd), n_weird, replace=False)]])
    acts = []
    rh.ctx.mode = "python"
    for w in sel:
        f = pk.unpack_fields(w)
        for k in ("rng_key", "rng_event", "moves"):
            f.pop(k)
        with contextlib.redirect_stdout(io.StringIO()):
            r = ev.choose_move_greedy(G(initial_state=f))
        acts.append(-1 if r is None else gai(r[0]))
    np.savez_compressed(os.path.join(OUT, "greedy.npz"), states=sel.astype(np.uint32), action=np.array(acts, dtype=np.int16))
    print(f"greedy: {len(sel)} states, {sum(a < 0 for a in acts)} without a legal move")


def gen_net(n=16):
    """a17: the reference's own AlphaZeroModel (model.py:277-357) with the test-size
    configuration (config.py:103-113), randomised BatchNorm statistics, eval mode: its
    state_dict, ``n`` inputs encoded by the reference from real positions, and its outputs
    (logits, tanh value, and ModelManager.predict's unmasked softmax, model.py:81-110)."""
    import importlib

    import torch

    rh.load_reference()
    if rh.REF_ROOT not in sys.path:
        sys.path.insert(0, rh.REF_ROOT)
    model_mod, cfgm = importlib.import_module("model"), importlib.import_module("config")
    pgs = importlib.import_module("process_game_state")
    mc = dict(cfgm.test_model_config) if hasattr(cfgm, "test_model_config") else dict(cfgm.model_config_default)
    torch.manual_seed(1234)
    ref = model_mod.AlphaZeroModel(
        input_channels=mc["input_channels"], cnn_filters=mc["cnn_filters"], board_size=mc["board_size"],
        action_size=mc["action_size"], global_feature_size=mc["global_feature_size"],
        value_hidden_dim=mc["value_head_hidden_dim"], num_res_blocks=mc["num_res_blocks"])
    for m in ref.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.3); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    ref.eval()
    boards, globs = [], []
    for g in range(n):
        st = rh.new_game_stream(1000 + g)
        for _ in range(3 + 4 * g):
            if st.is_game_over():
                break
            moves = sorted(st.get_legal_moves(), key=pgs.get_action_index)
            st = st.apply_move(moves[(7 * g + 3) % len(moves)])
        b, gl = pgs.create_state_tensors(st)
        boards.append(b); globs.append(gl)
    B, G = torch.stack(boards), torch.stack(globs)
    with torch.no_grad():
        logits, value = ref(B, G)
    out = {"cfg_" + k: np.array(v) for k, v in mc.items() if isinstance(v, (int, float))}
    out.update({"sd_" + k: v.numpy() for k, v in ref.state_dict().items()})
    out.update(board=B.numpy(), glob=G.numpy(), logits=logits.numpy(), value=value.numpy().reshape(-1),
               probs=torch.softmax(logits, dim=1).numpy())
    np.savez_compressed(os.path.join(OUT, "net.npz"), **out)
    print(f"net: {sum(v.numel() for v in ref.state_dict().values())} values in the state_dict, {n} positions")


def _reference_model(mc):
    import importlib

    rh.load_reference()
    if rh.REF_ROOT not in sys.path:
        sys.path.insert(0, rh.REF_ROOT)
    model_mod = importlib.import_module("model")
    return model_mod.AlphaZeroModel(
        input_channels=mc["input_channels"], cnn_filters=mc["cnn_filters"], board_size=mc["board_size"],
        action_size=mc["action_size"], global_feature_size=mc["global_feature_size"],
        value_hidden_dim=mc["value_head_hidden_dim"], num_res_blocks=mc["num_res_blocks"])


def _positions(n, pgs):
    boards, globs = [], []
    for g in range(n):
        st = rh.new_game_stream(2000 + g)
        for _ in range(2 + 3 * g):
            if st.is_game_over():
                break
            moves = sorted(st.get_legal_moves(), key=pgs.get_action_index)
            st = st.apply_move(moves[(5 * g + 1) % len(moves)])
        b, gl = pgs.create_state_tensors(st)
        boards.append(b); globs.append(gl)
    import torch

    return torch.stack(boards), torch.stack(globs)


def gen_net_default(n=24, seed=20260):
    """a17 / f4 at the DEFAULT size (config.py:18-29: 128 filters x 8 blocks, the network every
    throughput number uses): the reference's own AlphaZeroModel with weights from
    oracle/synth_weights.fill_ (numpy PCG64, so the fixture needs no 10 MB state_dict), eval mode;
    stored: the inputs (encoded by the reference from real positions), logits, value, softmax, and
    the tower output (what the hand-written sm_100a tower must reproduce)."""
    import importlib

    import torch

    try:
        from oracle import synth_weights
    except ImportError:          # run as a script: oracle/ itself is on sys.path
        import synth_weights

    rh.load_reference()
    if rh.REF_ROOT not in sys.path:
        sys.path.insert(0, rh.REF_ROOT)
    cfgm, pgs = importlib.import_module("config"), importlib.import_module("process_game_state")
    mc = dict(cfgm.model_config_default)
    ref = synth_weights.fill_(_reference_model(mc), seed).eval()
    B, G = _positions(n, pgs)
    with torch.no_grad():
        logits, value = ref(B, G)
        x = torch.relu(ref.bn(ref.conv(B)))
        for blk in ref.residual_blocks:
            x = blk(x)
    np.savez_compressed(os.path.join(OUT, "net_default.npz"), seed=np.int64(seed), board=B.numpy(), glob=G.numpy(), logits=logits.numpy(),
                        value=value.numpy().reshape(-1), probs=torch.softmax(logits, dim=1).numpy(),
                        tower_abs_max=np.float32(x.abs().max()), tower_sample=x[:4].numpy().astype(np.float16))
    print(f"net_default: {sum(v.numel() for v in ref.state_dict().values())} synthetic values (seed {seed}), {n} positions, "
          f"|logits| max {float(logits.abs().max()):.3f}")


def gen_mcts_real_priors(n_searches=16, sims=100, seed=20261):
    """a12-a16 with REAL priors: the reference's MCTS.py searching with the reference's default-size
    AlphaZeroModel behind the reference's ModelManager.predict (fp32 softmax, unmasked, model.py:81-110).
    Every (priors, value) the network returned is recorded in call order; the GPU tree is then fed the
    same table and must reproduce N, W, P, pi and the chosen move."""
    import importlib

    import torch

    try:
        from oracle import synth_weights
    except ImportError:          # run as a script: oracle/ itself is on sys.path
        import synth_weights

    ref = rh.load_reference()
    if rh.REF_ROOT not in sys.path:
        sys.path.insert(0, rh.REF_ROOT)
    cfgm = importlib.import_module("config")
    model_mod = importlib.import_module("model")
    import contextlib
    import io

    torch.set_num_threads(1)
    with contextlib.redirect_stdout(io.StringIO()):
        mm = model_mod.ModelManager(dict(cfgm.model_config_default), dict(cfgm.training_config_default, device="cpu"))
    synth_weights.fill_(mm.model, seed)
    mm.model.eval()

    class Recording:
        def __init__(self):
            self.rows = []

        def predict(self, b, g):
            p, v = mm.predict(b, g)
            self.rows.append((np.asarray(p, dtype=np.float32).copy(), np.float32(v)))
            return p, v

    gai = ref["pgs"].get_action_index
    nrng = np.random.default_rng(seed)
    out = {k: [] for k in "root skey testing eps noise choice_u move_no tau0 cpuct N W P pi action n_nodes n_edges n_eval".split()}
    table_p, table_v = [], []
    t0 = time.time()
    for i in range(n_searches):
        key = pk.rand(0x7EA1, i)
        s = rh.new_game_stream(key)
        moves = 0
        for _ in range([0, 1, 3, 4, 9, 17, 26, 33, 41, 47, 52, 55, 57, 58, 59, 60][i % 16]):
            if s.is_game_over():
                break
            lm = s.get_legal_moves()
            s = s.apply_move(lm[playout_pick(key, moves, len(lm))])
            moves += 1
        while s.is_game_over():          # never a terminal root here
            s = rh.new_game_stream(key + 1)
            moves = 0
        testing = i % 3 == 1
        cfg = {"num_simulations": sims, "cpuct": 2, "dirichlet_alpha": 0.4, "dirichlet_epsilon": 0.0 if testing else 0.25,
               "fpu_value": 0.25, "turns_until_tau0": 15, "action_size": 143, "testing": testing}
        noise = nrng.gamma(0.4, size=143).astype(np.float32) + np.float32(1e-6)
        u = float(np.float32(nrng.random()))
        skey = pk.rand(key ^ 0x5EA7C4, moves)
        rec = Recording()
        ev = rh.ctx.event
        mv, pi, info = rh.run_search(s, skey, cfg, moves, noise=None if testing else noise, choice_u=u, manager=rec)
        P = np.zeros((sims, 143), dtype=np.float32)
        V = np.zeros(sims, dtype=np.float32)
        for j, (p, v) in enumerate(rec.rows):
            P[j], V[j] = p, v
        table_p.append(P); table_v.append(V)
        for k, val in (("root", pk.pack_state(s, rng_key=key, rng_event=ev, moves=moves)), ("skey", skey), ("testing", int(testing)),
                       ("eps", cfg["dirichlet_epsilon"]), ("noise", noise), ("choice_u", u), ("move_no", moves), ("tau0", 15),
                       ("cpuct", 2.0), ("N", info["N"]), ("W", info["W"]), ("P", info["P"]), ("pi", pi),
                       ("action", -1 if mv is None else gai(mv)), ("n_nodes", info["n_nodes"]), ("n_edges", info["n_edges"]),
                       ("n_eval", len(rec.rows))):
            out[k].append(val)
    np.savez_compressed(
        os.path.join(OUT, "mcts_real.npz"), sims=np.int32(sims), seed=np.int64(seed),
        root=np.array(out["root"], dtype=np.uint32), skey=np.array(out["skey"], dtype=np.uint64),
        testing=np.array(out["testing"], dtype=np.uint8), eps=np.array(out["eps"], dtype=np.float64),
        noise=np.array(out["noise"], dtype=np.float32), choice_u=np.array(out["choice_u"], dtype=np.float32),
        move_no=np.array(out["move_no"], dtype=np.int32), tau0=np.array(out["tau0"], dtype=np.int32),
        cpuct=np.array(out["cpuct"], dtype=np.float64), N=np.array(out["N"], dtype=np.int32), W=np.array(out["W"], dtype=np.float64),
        P=np.array(out["P"], dtype=np.float32), pi=np.array(out["pi"], dtype=np.float64), action=np.array(out["action"], dtype=np.int16),
        n_nodes=np.array(out["n_nodes"], dtype=np.int32), n_edges=np.array(out["n_edges"], dtype=np.int32),
        n_eval=np.array(out["n_eval"], dtype=np.int32), table_p=np.array(table_p, dtype=np.float32), table_v=np.array(table_v, dtype=np.float32))
    print(f"mcts_real: {n_searches} searches x {sims} sims with the reference model's fp32 softmax priors, {time.time() - t0:.1f}s")


def gen_selfplay(n_games=3, sims=12, seed=4242):
    """f1: the reference's own ``self_play_worker`` (trainer.py:434-541), unmodified, run for
    ``n_games`` whole games.  Only its collaborators are pinned down: ModelManager is the fake
    evaluator (priors/value = exact function of the leaf hash), HarmoniesGameState() draws from
    the library's stream of game key rand(seed, game), and every search draws in-tree with the
    library's default search key rand(key ^ SEARCH_SALT, move number).  testing=True (no root
    noise, greedy move), as config.py:136-150.  Stored per example: the packed state before the
    search, pi (float32, as the worker stores it) and z."""
    import importlib

    ref = rh.load_reference()
    if rh.REF_ROOT not in sys.path:
        sys.path.insert(0, rh.REF_ROOT)
    tr = importlib.import_module("trainer")
    cfg = {"num_simulations": sims, "cpuct": 2, "dirichlet_alpha": 0.4, "dirichlet_epsilon": 0.25,
           "fpu_value": 0.25, "turns_until_tau0": 15, "action_size": 143, "testing": True}

    class FakeManager(rh.FakeModelManager):
        class _M:
            def load_state_dict(self, sd):
                pass

            def eval(self):
                pass

        def __init__(self, *a, **k):
            self.model = self._M()

    recorded = []
    real_search = tr.get_best_action_and_pi
    game_key = [0]

    def search(state, manager, mcts_config, move_number):
        recorded.append(pk.pack_state(state, rng_key=game_key[0], rng_event=rh.ctx.event, moves=move_number))
        saved = (rh.ctx.mode, rh.ctx.key, rh.ctx.event, rh.ctx.k)
        rh.ctx.mode, rh.ctx.key, rh.ctx.sim = "tree", pk.rand(game_key[0] ^ pk.SEARCH_SALT, move_number), -1
        try:
            return real_search(state, manager, mcts_config, move_number)
        finally:
            rh.ctx.mode, rh.ctx.key, rh.ctx.event, rh.ctx.k = saved

    tr.ModelManager, tr.get_best_action_and_pi = FakeManager, search
    states, pis, zs, gids, boards, globs = [], [], [], [], [], []
    t0 = time.time()
    try:
        for g in range(n_games):
            game_key[0] = pk.rand(seed, g)
            rh.ctx.mode, rh.ctx.key, rh.ctx.event, rh.ctx.k = "stream", game_key[0], 0, 0
            recorded.clear()
            data = tr.self_play_worker(({}, {}, {"device": "cpu"}, cfg, "cpu"))
            assert data and len(data) == len(recorded)
            for w, (b, gl, pi, z) in zip(recorded, data):
                states.append(w); pis.append(pi.numpy()); zs.append(float(z.item())); gids.append(g)
                boards.append(b.numpy()); globs.append(gl.numpy())
    finally:
        tr.get_best_action_and_pi = real_search
    np.savez_compressed(
        os.path.join(OUT, "selfplay.npz"), states=np.array(states, dtype=np.uint32), pi=np.array(pis, dtype=np.float32),
        z=np.array(zs, dtype=np.float32), game=np.array(gids, dtype=np.int32), seed=np.uint64(seed), sims=np.int32(sims),
        cpuct=np.float32(cfg["cpuct"]), board0=np.array(boards[:4], dtype=np.float32), glob0=np.array(globs[:4], dtype=np.float32))
    print(f"selfplay: {n_games} games, {len(states)} examples, {time.time() - t0:.1f}s")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "net_default" in (sys.argv[1:] or ["net_default"]):
        gen_net_default()
    if "mcts_real" in (sys.argv[1:] or ["mcts_real"]):
        gen_mcts_real_priors()
    if "net" in (sys.argv[1:] or ["net"]):
        gen_net()
    if "selfplay" in (sys.argv[1:] or ["selfplay"]):
        gen_selfplay()
    if "greedy" in (sys.argv[1:] or ["greedy"]):
        gen_greedy()
    which = sys.argv[1:] or ["engine", "scoring", "encode", "equiv", "mcts", "weird"]
    if "weird" in which:
        gen_weird()
    tb = ta = ss = None
    if "engine" in which or "encode" in which:
        tb, ta = gen_engine()
    if "scoring" in which or "encode" in which:
        ss = gen_scoring()
    if "encode" in which:
        term = ta[((ta[:, 22] >> 25) & 7) == 4]
        gen_encode(np.concatenate([tb, term]), ss)
    if "equiv" in which:
        gen_equiv()
    if "mcts" in which:
        gen_mcts()
