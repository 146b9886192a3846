"""Batch-size-1 round trips to the CUDA engine for the object-level drop-in API
(harmonies_engine.py / process_game_state.py / MCTS.py of this package).  Host code only
moves 128-byte records; all rules run in the kernels."""

import numpy as np
import torch

from . import batched as hb
from . import packed as pk

_device = None


def device():
    global _device
    if _device is None:
        hb._lib.load()   # raises if the CUDA library is not built
        if not torch.cuda.is_available():
            raise hb._lib.HarmoniesLibraryError("no CUDA device: the engine has no CPU fallback")
        _device = torch.device("cuda", torch.cuda.current_device())
    return _device


def _to_dev(words):
    return hb.states_from_numpy(np.asarray(words, dtype=np.uint32).reshape(1, 32), device())


def new_game(key):
    st = hb.init_states(1, device=device(), keys=np.array([key], dtype=np.uint64))
    return hb.states_to_numpy(st)[0]


def legal_actions(words):
    return pk.mask_to_actions(hb.states_to_numpy(hb.legal_mask(_to_dev(words)))[0])


def apply(words, action, draw=None):
    st = _to_dev(words)
    a = torch.tensor([action], dtype=torch.int16, device=device())
    d = None if draw is None else torch.from_numpy(np.array([draw], dtype=np.uint16).view(np.int16)).to(device())
    status = hb.apply(st, a, d)
    return hb.states_to_numpy(st)[0], int(status[0])


def end_turn(words):
    st = _to_dev(words)
    hb.end_turn(st)
    return hb.states_to_numpy(st)[0]


def replenish(words):
    st = _to_dev(words)
    hb.replenish_piles(st)
    return hb.states_to_numpy(st)[0]


def draw_tiles(words, count):
    st = _to_dev(words)
    t = hb.draw_tiles(st, count).cpu().numpy()[0]
    return hb.states_to_numpy(st)[0], [int(x) for x in t[: int(t[15])]]


def score_terms(words):
    """int[2][5]: grass, mountains, fields, buildings, water per player"""
    return hb.score(_to_dev(words), with_terms=True)[1].cpu().numpy()[0]


def scores(words):
    return hb.score(_to_dev(words)).cpu().numpy()[0]


def encode(words):
    b, g = hb.encode(_to_dev(words))
    return b[0].cpu(), g[0].cpu()
