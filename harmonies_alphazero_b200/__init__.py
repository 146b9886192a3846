"""B200-native batched Harmonies engine + MCTS self-play core.

Drop-in for the hot path of IllyaArtemchuk/Harmonies-Alphazero (harmonies_engine.py,
process_game_state.py, MCTS.py) behind the reference's own object API, implemented as
hand-written sm_100a CUDA kernels reached through the C ABI in include/harmonies_b200.h.
Submodules are imported lazily so that host-only utilities (constants, packed) work on a
machine without the CUDA library; everything that computes game logic needs it.
"""

__all__ = ["constants", "packed"]
__version__ = "0.1.0"
