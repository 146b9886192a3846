"""Drop-in for the reference's ``harmonies_engine.py`` (object-level boundary b1 of
SURVEY.md §8): same class name, attributes, method names, argument meaning and error
behaviour — but every rule is evaluated by the CUDA kernels through the C ABI, at batch
size 1.  There is no CPU fallback: without the CUDA library these methods raise.

The Python attributes are the source of truth (the reference's tests and UIs read and
mutate them directly, tests/test_harmonies_engine.py:87-163): each call packs them into the
128-byte record, runs the kernel, and unpacks the result into a NEW object (value
semantics of apply_move, harmonies_engine.py:210-211).

Throughput work does not go through this class — see batched.py / selfplay.py.
"""

import copy
import logging
import random

import numpy as np

from .constants import *  # noqa: F401,F403  (the reference module re-exports its constants the same way)
from .constants import TILE_TYPES, VALID_HEXES, NUM_HEXES
from . import packed as pk

logger_main = logging.getLogger("harmonies_b200.main")

PLAYER_BOARD_HEX_COUNT = NUM_HEXES
WATER_SCORES = {1: 0, 2: 2, 3: 5, 4: 8, 5: 11, 6: 15}

_MOVE_ERRORS = {
    1: "Invalid pile index: {move}",
    2: "Invalid move format for placement phase: {move}. Expected (tile_type, (q, r))",
    3: "Invalid coordinate: {move}",
    4: "Illegal move attempted: Tile '{tile}' not found in hand {hand}",
    5: "Illegal move attempted in apply_move: Cannot place {tile} on {coord} with stack {stack}",
    6: "Invalid turn phase: {phase}",
    7: "Replay draw not available in the bag",
}


def get_water_score(length):
    """harmonies_engine.py:21-27 (table lookup; the scoring itself runs on the GPU)."""
    if length <= 0:
        return 0
    return WATER_SCORES[length] if length in WATER_SCORES else WATER_SCORES[6] + (length - 6) * 4


def get_neighbors(coord):
    """harmonies_engine.py:31-43."""
    if coord not in VALID_HEXES:
        return []
    q, r = coord
    return [(q + dq, r + dr) for dq, dr in AXIAL_DIRECTIONS if (q + dq, r + dr) in VALID_HEXES]  # noqa: F405


def _dev():
    from . import _single

    return _single


class HarmoniesGameState:
    def __init__(self, initial_state=None):
        if initial_state:
            self.__dict__.update(initial_state)          # harmonies_engine.py:67-68
            self.__dict__.setdefault("_rng_key", random.getrandbits(64))
            self.__dict__.setdefault("_rng_event", 0)
            self.__dict__.setdefault("_moves", 0)
        else:
            # harmonies_engine.py:70-79: empty boards, full bag, five piles drawn.  The draw
            # stream is keyed from Python's global `random`, so random.seed() makes games
            # reproducible as it does for the reference.
            key = random.getrandbits(64)
            self._load(_dev().new_game(key))

    def _load(self, words):
        f = pk.unpack_fields(words)
        self._rng_key, self._rng_event, self._moves = f.pop("rng_key"), f.pop("rng_event"), f.pop("moves")
        self.__dict__.update(f)

    def _pack(self):
        return pk.pack_fields(
            self.player_boards, self.tile_bag, self.available_piles, self.current_player, self.tiles_in_hand,
            self.turn_phase, self.game_over, self.winner, self.final_scores,
            rng_key=self.__dict__.get("_rng_key", 0), rng_event=self.__dict__.get("_rng_event", 0),
            moves=self.__dict__.get("_moves", 0),
        )

    # ---- identity (pure attribute logic, harmonies_engine.py:81-118) ------------------------
    def get_canonical_tuple(self):
        boards = tuple(
            tuple((coord, tuple(stack)) for coord, stack in sorted(self.player_boards[p].items())) for p in (0, 1)
        )
        return (
            self.current_player,
            self.turn_phase,
            tuple(sorted(self.tiles_in_hand)),
            tuple(tuple(sorted(pile)) for pile in self.available_piles),
            tuple(sorted(self.tile_bag.items())),
            boards[0],
            boards[1],
        )

    def __hash__(self):
        return hash(self.get_canonical_tuple())

    def __eq__(self, other):
        if not isinstance(other, HarmoniesGameState):
            return NotImplemented
        return self.get_canonical_tuple() == other.get_canonical_tuple()

    # ---- rules (GPU) ---------------------------------------------------------------------------
    def get_current_player(self):
        return self.current_player

    def get_legal_moves(self):
        """harmonies_engine.py:145-208.  Order: ascending action index (the reference's order
        is PYTHONHASHSEED-dependent)."""
        if self.turn_phase != "choose_pile" and not str(self.turn_phase).startswith("place_tile"):
            logger_main.warning(f"get_legal_moves called during unexpected phase: {self.turn_phase}")
            return []
        return [pk.action_to_move(a) for a in _dev().legal_actions(self._pack())]

    def apply_move(self, move):
        """harmonies_engine.py:210-298: returns a NEW state; ValueError on any illegal input."""
        phase = self.turn_phase
        if phase == "choose_pile":
            if not isinstance(move, int) or isinstance(move, bool) or not (0 <= move < len(self.available_piles)):
                raise ValueError(f"Invalid pile index: {move}")
            action = move
        elif str(phase).startswith("place_tile"):
            if not (isinstance(move, tuple) and len(move) == 2 and isinstance(move[0], str)
                    and move[0] in TILE_TYPES and isinstance(move[1], tuple)):
                raise ValueError(f"Invalid move format for placement phase: {move}. Expected (tile_type, (q, r))")
            tile, coord = move
            if coord not in VALID_HEXES:
                raise ValueError(f"Invalid coordinate: {coord}")
            action = 5 + 23 * TILE_TYPES.index(tile) + coordinate_to_index_map[coord]  # noqa: F405
        else:
            raise ValueError(f"Invalid turn phase: {phase}")
        words, status = _dev().apply(self._pack(), action)
        if status != 0:
            tile, coord = (move if isinstance(move, tuple) else (None, None))
            raise ValueError(_MOVE_ERRORS[status].format(
                move=move, tile=tile, coord=coord, hand=self.tiles_in_hand, phase=phase,
                stack=self.player_boards[self.current_player].get(coord)))
        new = HarmoniesGameState.__new__(HarmoniesGameState)
        new._load(words)
        return new

    # ---- the reference's private helpers (its GUI, harnesses and tests call them directly) -----
    def _adopt(self, words):
        """Mutate THIS object into the state the kernel produced (the reference's private helpers
        work in place)."""
        self._load(words)

    def _draw_tiles(self, num_tiles):
        """harmonies_engine.py:120-130: draws from (and decrements) the bag; returns tile names."""
        words, tiles = _dev().draw_tiles(self._pack(), int(num_tiles))
        f = pk.unpack_fields(words)
        self.tile_bag = f["tile_bag"]
        self._rng_event = f["rng_event"]
        return [TILE_TYPES[t] for t in tiles]

    def _replenish_piles(self):
        """harmonies_engine.py:132-137."""
        f = pk.unpack_fields(_dev().replenish(self._pack()))
        self.tile_bag, self.available_piles, self._rng_event = f["tile_bag"], f["available_piles"], f["rng_event"]

    def _get_top_tile(self, board, coord):
        return board.get(coord, [None])[-1]                        # harmonies_engine.py:142-143

    def _end_turn_actions(self):
        """harmonies_engine.py:301-329, in place: replenish, end-of-game triggers, turn switch or
        final scoring for the player who has just finished a turn."""
        self._adopt(_dev().end_turn(self._pack()))

    def _calculate_final_scores(self):                             # harmonies_engine.py:344-346
        self.final_scores[0] = self.calculate_score_for_player(0)
        self.final_scores[1] = self.calculate_score_for_player(1)

    def _determine_winner(self):                                   # harmonies_engine.py:348-354
        a, b = self.final_scores
        self.winner = 0 if a > b else 1 if b > a else -1

    def _score_term(self, board, term):
        """One scoring term of ``board`` (any dict coord -> stack) on the GPU: the board is packed as
        player 0 of a scratch state."""
        words = pk.pack_fields([board, {}], {t: 0 for t in TILE_TYPES}, [], 0, [], "choose_pile", False, None, [0, 0])
        return int(_dev().score_terms(words)[0][term])

    def _score_grass(self, board, player):                         # harmonies_engine.py:369-385
        return self._score_term(board, 0)

    def _score_mountains(self, board, player):                     # :392-413
        return self._score_term(board, 1)

    def _score_fields(self, board, player):                        # :424-443
        return self._score_term(board, 2)

    def _score_buildings(self, board, player):                     # :454-469
        return self._score_term(board, 3)

    def _score_water(self, board, player):                         # :480-523
        return self._score_term(board, 4)

    def is_game_over(self):
        return self.game_over and self.winner is not None          # harmonies_engine.py:332-333

    def get_game_outcome(self):
        if not self.is_game_over():
            return None
        return 1 if self.winner == 0 else -1 if self.winner == 1 else 0   # :335-342

    def calculate_score_for_player(self, player_id):
        """harmonies_engine.py:357-367 (grass, mountains, fields, buildings, water)."""
        return int(_dev().scores(self._pack())[player_id])

    def clone(self):
        return copy.deepcopy(self)

    def __str__(self):
        s = "--- Harmonies State (Grid: 5-4-5-4-5 rows) ---\n"
        s += f"Player Turn: {self.current_player}, Phase: {self.turn_phase}\n"
        s += f"Game Over: {self.is_game_over()}, Winner: {self.winner}, Scores: {self.final_scores}\n"
        s += f"Bag: {dict(sorted(self.tile_bag.items()))}\n"
        s += f"Available Piles: {self.available_piles}\n"
        s += f"Player {self.current_player} Hand: {self.tiles_in_hand}\n"
        for p in (0, 1):
            s += f"Player {p} Board ({len(self.player_boards[p])}/{PLAYER_BOARD_HEX_COUNT} hexes):\n"
            s += f"  { {str(c): st for c, st in sorted(self.player_boards[p].items())} }\n"
        return s + "---------------------------------------------\n"


__all__ = ["HarmoniesGameState", "VALID_HEXES", "TILE_TYPES", "get_neighbors", "get_water_score", "PILE_SIZE", "NUM_PILES"]  # noqa: F405
_ = np  # numpy is part of the reference module's namespace too
