"""Drop-in for the reference's ``MCTS.py`` search entry point:

    get_best_action_and_pi(game_state, model_manager, mcts_config, game_move_number)
        -> (move | None, np.ndarray[action_size] float64)

Same arguments, config keys (num_simulations, cpuct, dirichlet_alpha, dirichlet_epsilon,
turns_until_tau0, action_size, testing), return convention and fallbacks as MCTS.py:272-441.
The tree itself (Node/Edge/MCTS, move_to_leaf, expand_leaf, back_fill — MCTS.py:8-264)
lives in flat GPU arrays behind hz_tree_*; this function is the B = 1 use of it, calling
``model_manager.predict(board, global)`` once per simulation with a non-terminal leaf, like
the reference (MCTS.py:297-304).  Randomness comes from the same generators the reference uses
(np.random for Dirichlet noise and the exploratory move, ``random`` for the fallback move
and, here, for the key of the in-tree draw stream).

For throughput use selfplay.BatchedSelfPlay: thousands of trees per call, the network batched
over all leaves.
"""

import logging
import random

import numpy as np
import torch

from . import batched as hb
from . import packed as pk
from .process_game_state import get_action_index, _words
from .tree import BatchedMCTS, search_keys_tensor

logger_mcts = logging.getLogger("harmonies_b200.mcts")

_trees = {}


def _tree_for(sims, key_mode):
    k = (int(sims), key_mode)
    if k not in _trees:
        _trees[k] = BatchedMCTS(1, int(sims), key_mode=key_mode)
    return _trees[k]


def get_best_action_and_pi(game_state, model_manager, mcts_config, game_move_number, key_mode=hb.KEY_REFERENCE):
    sims = int(mcts_config["num_simulations"])
    action_size = int(mcts_config.get("action_size", 143))
    testing = bool(mcts_config.get("testing", False))
    dev = torch.device("cuda", torch.cuda.current_device())
    pi_target = np.zeros(action_size, dtype=int)

    if sims > 0:
        tree = _tree_for(sims, key_mode)
        root = hb.states_from_numpy(_words(game_state).reshape(1, 32), dev)
        tree.reset(root, search_keys_tensor([random.getrandbits(64)], dev))
        board = torch.empty((1, 38, 5, 7), dtype=torch.float32, device=dev)
        glob = torch.empty((1, 42), dtype=torch.float32, device=dev)
        policy = torch.zeros((1, 143), dtype=torch.float32, device=dev)
        value = torch.zeros(1, dtype=torch.float32, device=dev)
        leaf = torch.empty((1, 32), dtype=torch.int32, device=dev)
        noise, eps = None, 0.0
        if not testing:                                          # MCTS.py:308-326
            eps = float(mcts_config["dirichlet_epsilon"])
            g = np.random.gamma(float(mcts_config["dirichlet_alpha"]), size=143).astype(np.float32)
            noise = torch.from_numpy(np.maximum(g, np.float32(1e-30))).view(1, 143).to(dev)
        for _ in range(sims):                                    # MCTS.py:291
            tree.select(float(mcts_config["cpuct"]), board, glob, leaf_states=leaf)
            # a terminal leaf is not evaluated by the network (MCTS.py:297,333-341): the kernel backs up
            # the game outcome itself; is_game_over <=> winner bits of the meta byte are set
            if ((int(leaf[0, 22].item()) >> 29) & 3) == 0:
                p, v = model_manager.predict(board[0].cpu(), glob[0].cpu())   # MCTS.py:302 (batch 1, as the reference)
                policy.copy_(torch.as_tensor(np.asarray(p, dtype=np.float32)).view(1, -1)[:, :143])
                value.fill_(float(v))
            tree.expand_backup(policy, value, noise=noise, eps=eps)
        tree.check_status()
        visits = tree.root_policy()[0][0].cpu().numpy()
    else:
        visits = np.zeros(143, dtype=np.int64)

    n = min(action_size, 143)
    pi_target[:n] = visits[:n]
    total_visits = int(visits.sum())
    # the reference iterates the root's edges: they exist once the root has been expanded (>= 1
    # simulation on a non-terminal root) and are its legal moves, visited or not, in edge order
    legal_moves = game_state.get_legal_moves()
    root_expanded = sims > 0 and not game_state.is_game_over()
    visit_counts = [(m, int(visits[get_action_index(m)])) for m in legal_moves] if root_expanded else []
    if total_visits > 0:
        pi_target = pi_target / total_visits                     # MCTS.py:378-381
    else:
        logger_mcts.warning("MCTS root had zero total visits after simulations.")
        # MCTS.py:386-392 writes 1/len(legal) into the INTEGER array allocated at :356, which truncates
        # to 0: the reference returns an all-zero int vector here, and so does this
        for m in legal_moves:
            pi_target[get_action_index(m)] = 1.0 / len(legal_moves)

    best_action = None
    exploratory = (not testing) and game_move_number < mcts_config["turns_until_tau0"]   # MCTS.py:399-402
    if exploratory:
        if total_visits > 0 and visit_counts:
            probs = np.array([vc[1] for vc in visit_counts], dtype=float) / total_visits
            best_action = visit_counts[np.random.choice(len(visit_counts), p=probs)][0]    # MCTS.py:411
    else:
        max_visits = -1
        for m, vcount in visit_counts:                           # first max, MCTS.py:420-423 (0 > -1: an
            if vcount > max_visits:                              # unvisited first edge wins when nothing was visited)
                max_visits, best_action = vcount, m
    if best_action is None:                                      # MCTS.py:425-439
        logger_mcts.warning("MCTS could not select a best action; falling back to a random legal move.")
        if legal_moves:
            best_action = random.choice(legal_moves)
        else:
            logger_mcts.error("MCTS failed, and no legal moves exist. Game should have ended.")
            return None, pi_target
    return best_action, pi_target


__all__ = ["get_best_action_and_pi"]
_ = pk
