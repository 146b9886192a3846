"""Game constants of the simplified 2-player Harmonies variant.

Same names and values as the reference's ``constants.py:1-52`` so that code written against
the reference (``from constants import *``) keeps working against this package.  The hex
grid is derived from the row pattern instead of being listed by hand.
"""

TILE_TYPES = ["water", "plant", "wood", "stone", "building", "field"]
WATER, PLANT, WOOD, STONE, BUILDING, FIELD = TILE_TYPES

# 5-4-5-4-5 rows, axial (q, r); row r starts at q = -1 - (r + 2 + 1) // 2 ... derived below
_ROW_LEN = {-2: 5, -1: 4, 0: 5, 1: 4, 2: 5}
_ROW_Q0 = {-2: -1, -1: -1, 0: -2, 1: -2, 2: -3}
VALID_HEXES = {(q0 + i, r) for r, q0 in _ROW_Q0.items() for i in range(_ROW_LEN[r])}

AXIAL_DIRECTIONS = [(1, 0), (-1, 0), (0, 1), (0, -1), (1, -1), (-1, 1)]
BOARD_SIZE = (5, 7)

# NB: dict order (field before building) is the reference's (constants.py:41) and is the
# order random.sample sees the flattened bag in; type-indexed code uses TILE_TYPES order.
INITIAL_BAG = {WATER: 23, PLANT: 19, WOOD: 21, STONE: 23, FIELD: 19, BUILDING: 15}
NUM_PILES = 5
PILE_SIZE = 3
NUM_HEXES = 23
EMPTY_HEX_END_THRESHOLD = 2

sorted_coords = sorted(VALID_HEXES)
coordinate_to_index_map = {coord: index for index, coord in enumerate(sorted_coords)}

INPUT_CHANNELS = 38
GLOBAL_FEATURE_SIZE = 42
ACTION_SIZE = 143

# ---- derived tables used by the packed format (include/harmonies_b200.h) -------------
INITIAL_BAG_BY_TYPE = [INITIAL_BAG[t] for t in TILE_TYPES]          # [23,19,21,23,15,19]
TYPE_INDEX = {t: i for i, t in enumerate(TILE_TYPES)}
PHASES = ["choose_pile", "place_tile_1", "place_tile_2", "place_tile_3", "game_over"]
PHASE_INDEX = {p: i for i, p in enumerate(PHASES)}


def neighbor_indices(i):
    """Hex indices adjacent to hex index ``i`` (harmonies_engine.py:31-43)."""
    q, r = sorted_coords[i]
    out = []
    for dq, dr in AXIAL_DIRECTIONS:
        j = coordinate_to_index_map.get((q + dq, r + dr))
        if j is not None:
            out.append(j)
    return sorted(out)


NEIGHBOR_MASKS = [sum(1 << j for j in neighbor_indices(i)) for i in range(NUM_HEXES)]
# (y, x) cell of hex i in the 5x7 plane: y = r + 2, x = q + 3 (process_game_state.py:9-12,36-37)
HEX_CELL = [(r + 2, q + 3) for (q, r) in sorted_coords]
