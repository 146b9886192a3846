"""Drop-in for the reference's ``process_game_state.py``: same function names, shapes, dtypes
and errors.  The tensors are produced by the hz_encode CUDA kernel (bit-exact fp32 against
process_game_state.py:15-137); ``get_action_index`` is host-side index arithmetic."""

import torch

from .constants import NUM_HEXES, NUM_PILES, TILE_TYPES, coordinate_to_index_map
from . import packed as pk


def _words(game_state):
    if hasattr(game_state, "_pack"):
        return game_state._pack()
    return pk.pack_state(game_state)   # any object with the reference's attributes


def create_state_tensors(game_state):
    """-> (board float32[38,5,7], global float32[42]) CPU tensors (process_game_state.py:15-16)."""
    from . import _single

    return _single.encode(_words(game_state))


def create_board_tensor(game_state):
    return create_state_tensors(game_state)[0]


def create_global_features(game_state):
    return create_state_tensors(game_state)[1]


def get_action_index(action, hand_tiles=None):
    """Move -> flat index 0..142 (process_game_state.py:156-177), same ValueErrors."""
    if isinstance(action, int):
        if 0 <= action < NUM_PILES:
            return action
        raise ValueError(f"Invalid pile index action: {action}")
    if isinstance(action, tuple) and len(action) == 2:
        tile_type, coord = action
        if tile_type not in TILE_TYPES:
            raise ValueError(f"Invalid tile type in action: {tile_type}")
        if coord not in coordinate_to_index_map:
            raise ValueError(f"Invalid coordinate in action: {coord}")
        return NUM_PILES + TILE_TYPES.index(tile_type) * NUM_HEXES + coordinate_to_index_map[coord]
    raise ValueError(f"Invalid action format: {action}")


__all__ = ["create_state_tensors", "create_board_tensor", "create_global_features", "get_action_index"]
_ = torch
