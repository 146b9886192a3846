"""TEST INFRASTRUCTURE — drives the *unmodified* reference (read-only at /root/reference)
under controlled determinism so that golden vectors can be generated from it.

Only ``oracle/gen_golden.py`` and (when /root/reference exists) CPU tests import this.
Nothing here ships in the product path and nothing here runs on the GPU box.

Determinism controls:
  * ``loggers`` is replaced by a stub before import (loggers.py:22-35 opens files under the
    read-only tree at import time);
  * ``HarmoniesGameState.get_legal_moves`` is wrapped to return moves in ascending
    ``get_action_index`` order (the reference's order is ``list(set(...))``,
    harmonies_engine.py:164,203, i.e. PYTHONHASHSEED-dependent);
  * ``_draw_tiles`` (harmonies_engine.py:120-130) is wrapped to either record Python's
    ``random`` draws or to use the library's counter-based draw source
    (include/harmonies_b200.h) keyed per game stream or per (simulation, action) in a tree.
"""

import logging
import os
import random
import sys
import types

import numpy as np

REF_ROOT = os.environ.get("HZ_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

from harmonies_alphazero_b200 import packed as pk  # noqa: E402
from harmonies_alphazero_b200.constants import TILE_TYPES  # noqa: E402


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "harmonies_engine.py"))


class Ctx:
    """Mutable draw context shared by the wrappers."""

    mode = "python"  # "python" | "stream" | "tree"
    key = 0
    event = 0
    k = 0
    sim = -1
    recorded = None  # list of pile multiset codes drawn since last clear
    leaf_hash = None
    noise = None  # vector returned by np.random.dirichlet
    choice_u = None  # uniform used by np.random.choice


ctx = Ctx()
_loaded = {}


def load_reference():
    """Import the reference modules (once) with the wrappers installed."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    stub = types.ModuleType("loggers")
    for n in ["logger_main", "logger_mcts", "logger_model", "logger_tourney", "logger_memory"]:
        lg = logging.getLogger("hzref." + n)
        lg.addHandler(logging.NullHandler())
        lg.propagate = False
        lg.setLevel(logging.CRITICAL + 1)
        lg.disabled = True
        setattr(stub, n, lg)
    sys.modules["loggers"] = stub
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_ROOT)
    try:
        import harmonies_engine as he
        import process_game_state as pgs
        import MCTS as mcts
    finally:
        sys.path.remove(REF_ROOT)
    G = he.HarmoniesGameState

    orig_legal = G.get_legal_moves
    orig_draw = G._draw_tiles
    orig_replenish = G._replenish_piles
    orig_apply = G.apply_move

    def legal_sorted(self):
        return sorted(orig_legal(self), key=pgs.get_action_index)

    def draw_tiles(self, num_tiles):
        if ctx.mode == "python":
            drawn = orig_draw(self, num_tiles)
        else:
            bag = [self.tile_bag[t] for t in TILE_TYPES]
            z = pk.rand(ctx.key, ctx.event * 8 + ctx.k)
            ctx.k += 1
            idx = pk.draw_pile(bag, z, num_tiles)
            drawn = [TILE_TYPES[t] for t in idx]
            for t in drawn:
                self.tile_bag[t] -= 1
        if ctx.recorded is not None and drawn:
            ctx.recorded.append(pk.multiset_code(drawn))
        return drawn

    def replenish(self):
        ctx.k = 0
        orig_replenish(self)
        if ctx.mode == "stream":
            ctx.event += 1

    def apply_move(self, move):
        if ctx.mode == "tree":
            ctx.event = (ctx.sim << 8) | pgs.get_action_index(move)
        return orig_apply(self, move)

    G.get_legal_moves = legal_sorted
    G._draw_tiles = draw_tiles
    G._replenish_piles = replenish
    G.apply_move = apply_move

    # --- MCTS instrumentation: simulation counter, leaf hash, injected randomness
    orig_move_to_leaf = mcts.MCTS.move_to_leaf

    def move_to_leaf(self):
        ctx.sim += 1
        return orig_move_to_leaf(self)

    mcts.MCTS.move_to_leaf = move_to_leaf
    orig_cst = mcts.create_state_tensors

    def create_state_tensors(state):
        ctx.leaf_hash = pk.canon_hash(pk.pack_state(state))
        return orig_cst(state)

    mcts.create_state_tensors = create_state_tensors

    _loaded.update(he=he, pgs=pgs, mcts=mcts, G=G)
    return _loaded


class FakeModelManager:
    """ModelManager.predict stand-in (model.py:81-110): priors/value are an exact function
    of the leaf's canonical hash (packed.fake_eval)."""

    def predict(self, board_tensor, global_features_tensor):
        return pk.fake_eval(ctx.leaf_hash)


def new_game_python(seed):
    """Fresh reference game whose draws come from Python's seeded global ``random``."""
    ref = load_reference()
    ctx.mode = "python"
    random.seed(seed)
    return ref["G"]()


def new_game_stream(key):
    """Fresh reference game whose draws come from the library's stream ``key``."""
    ref = load_reference()
    ctx.mode, ctx.key, ctx.event, ctx.k = "stream", key, 0, 0
    return ref["G"]()


def run_search(state, search_key, mcts_config, move_number, noise=None, choice_u=None, manager=None):
    """Reference get_best_action_and_pi (MCTS.py:272-441) with in-tree draws keyed by
    (simulation, action), the fake evaluator and injected Dirichlet noise / choice uniform.
    Returns (move, pi[143] float64, info)."""
    ref = load_reference()
    saved = (ctx.mode, ctx.key, ctx.event, ctx.k)
    ctx.mode, ctx.key, ctx.sim = "tree", search_key, -1
    captured = {}
    orig_dirichlet, orig_choice = np.random.dirichlet, np.random.choice
    OrigMCTS = ref["mcts"].MCTS

    class CapturingMCTS(OrigMCTS):
        def __init__(self, root, cfg):
            super().__init__(root, cfg)
            captured["mcts"] = self

    legal_idx = [ref["pgs"].get_action_index(m) for m in state.get_legal_moves()]

    def dirichlet(alpha, size=None):
        # noise is indexed by action; normalise over the legal moves with a sequential
        # fp64 sum in ascending action order (the library's definition)
        assert len(alpha) == len(legal_idx)
        g = [float(np.float32(noise[a])) for a in legal_idx]
        s = 0.0
        for x in g:
            s += x
        return np.array([x / s for x in g], dtype=np.float64)

    def choice(n, p=None):
        # integer restatement of the inverse-CDF draw: first i with cumN[i] > u*sumN
        counts = [e.stats["N"] for e in captured["mcts"].root.edges.values()]
        tot = sum(counts)
        acc = 0
        for i, c in enumerate(counts):
            acc += c
            if acc > choice_u * tot:
                return i
        return len(counts) - 1

    ref["mcts"].MCTS = CapturingMCTS
    if noise is not None:
        np.random.dirichlet = dirichlet
    if choice_u is not None:
        np.random.choice = choice
    try:
        move, pi = ref["mcts"].get_best_action_and_pi(
            state.clone(), manager if manager is not None else FakeModelManager(), mcts_config, move_number
        )
    finally:
        ref["mcts"].MCTS = OrigMCTS
        np.random.dirichlet, np.random.choice = orig_dirichlet, orig_choice
        ctx.mode, ctx.key, ctx.event, ctx.k = saved
    tree = captured["mcts"]
    gai = ref["pgs"].get_action_index
    N = np.zeros(143, dtype=np.int32)
    W = np.zeros(143, dtype=np.float64)
    P = np.zeros(143, dtype=np.float32)
    for a, e in tree.root.edges.items():
        i = gai(a)
        N[i], W[i], P[i] = e.stats["N"], e.stats["W"], e.stats["P"]
    n_edges = sum(len(n.edges) for n in tree.tree.values())
    info = {"N": N, "W": W, "P": P, "n_nodes": len(tree.tree), "n_edges": n_edges}
    return move, np.asarray(pi, dtype=np.float64), info
