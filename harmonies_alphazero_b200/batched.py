"""Tensor-level API of the batched engine: thin wrappers that pass torch CUDA tensors'
device pointers and the current stream to the C ABI (include/harmonies_b200.h).

States are ``torch.int32[n, 32]`` CUDA tensors holding the packed 128-byte records
(int32 is a bit-level view of the uint32 words).  PyTorch is plumbing only: memory,
streams, and later torch.distributed.
"""

import numpy as np
import torch

from . import _lib

F32, BF16 = 0, 1
NCHW, NHWC = 0, 1
KEY_EXACT, KEY_REFERENCE = 0, 1
NO_DRAW = 0xFFFF


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _check_states(states):
    if not (states.is_cuda and states.dtype == torch.int32 and states.dim() == 2 and states.shape[1] == 32):
        raise TypeError("states must be a CUDA int32 tensor of shape [n, 32]")
    if not states.is_contiguous():
        raise ValueError("states must be contiguous")
    return states.shape[0]


def _ptr(t):
    return None if t is None else t.data_ptr()


def states_from_numpy(words, device="cuda"):
    """np.uint32[n, 32] -> CUDA int32[n, 32]."""
    a = np.ascontiguousarray(words, dtype=np.uint32).reshape(-1, 32)
    return torch.from_numpy(a.view(np.int32).copy()).to(device)


def states_to_numpy(states):
    return states.detach().cpu().numpy().view(np.uint32)


def init_states(n, device="cuda", keys=None, seed=0, first_id=0):
    """n new games (HarmoniesGameState.__init__, harmonies_engine.py:66-79)."""
    lib = _lib.load()
    dev = torch.device(device)
    states = torch.empty((n, 32), dtype=torch.int32, device=dev)
    if keys is not None:
        keys = torch.as_tensor(np.asarray(keys, dtype=np.uint64).view(np.int64)).to(dev)
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_init_states(_ptr(states), n, _ptr(keys), seed, first_id, _stream(states)), "hz_init_states")
    return states


def legal_mask(states, out=None):
    """int32[n, 5]: 143-bit legal-action masks (get_legal_moves, harmonies_engine.py:145-208)."""
    lib = _lib.load()
    n = _check_states(states)
    out = torch.empty((n, 5), dtype=torch.int32, device=states.device) if out is None else out
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_legal_mask(_ptr(states), n, _ptr(out), _stream(states)), "hz_legal_mask")
    return out


def apply(states, actions, draws=None, status=None):
    """In-place apply_move (harmonies_engine.py:210-329).  actions int16[n]; draws optional
    int16[n] (bit view of uint16 pile codes, 0xFFFF = none).  Returns status uint8[n]."""
    lib = _lib.load()
    n = _check_states(states)
    if actions.dtype != torch.int16 or actions.shape[0] != n:
        raise TypeError("actions must be int16[n]")
    if draws is not None and (draws.dtype != torch.int16 or draws.shape[0] != n):
        raise TypeError("draws must be int16[n] (uint16 bit pattern)")
    status = torch.empty(n, dtype=torch.uint8, device=states.device) if status is None else status
    with torch.cuda.device(states.device):
        _lib.check(
            lib.hz_apply(_ptr(states), n, _ptr(actions), _ptr(draws), _ptr(status), _stream(states)), "hz_apply"
        )
    return status


def score(states, with_terms=False):
    """int16[n, 2] scores (calculate_score_for_player, harmonies_engine.py:357-523);
    with_terms also returns int16[n, 2, 5] (grass, mountains, fields, buildings, water)."""
    lib = _lib.load()
    n = _check_states(states)
    sc = torch.empty((n, 2), dtype=torch.int16, device=states.device)
    tm = torch.empty((n, 2, 5), dtype=torch.int16, device=states.device) if with_terms else None
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_score(_ptr(states), n, _ptr(sc), _ptr(tm), _stream(states)), "hz_score")
    return (sc, tm) if with_terms else sc


def encode(states, dtype=torch.float32, channels_last=False, board=None, glob=None):
    """create_state_tensors (process_game_state.py:15-137) for a batch.

    Returns (board, glob): board is logically [n, 38, 5, 7]; with channels_last the memory
    layout is NHWC (torch.channels_last strides)."""
    lib = _lib.load()
    n = _check_states(states)
    code = {torch.float32: F32, torch.bfloat16: BF16}[dtype]
    if board is None:
        board = torch.empty(
            (n, 38, 5, 7), dtype=dtype, device=states.device,
            memory_format=torch.channels_last if channels_last else torch.contiguous_format,
        )
    if glob is None:
        glob = torch.empty((n, 42), dtype=dtype, device=states.device)
    with torch.cuda.device(states.device):
        _lib.check(
            lib.hz_encode(_ptr(states), n, _ptr(board), _ptr(glob), code, NHWC if channels_last else NCHW, _stream(states)),
            "hz_encode",
        )
    return board, glob


def canon_hash(states, mode=KEY_EXACT):
    """int64[n] (bit view of uint64) node keys (harmonies_engine.py:81-113)."""
    lib = _lib.load()
    n = _check_states(states)
    out = torch.empty(n, dtype=torch.int64, device=states.device)
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_canon_hash(_ptr(states), n, mode, _ptr(out), _stream(states)), "hz_canon_hash")
    return out


def outcome(states):
    """(over uint8[n], outcome int8[n]) — is_game_over / get_game_outcome (:332-342)."""
    lib = _lib.load()
    n = _check_states(states)
    over = torch.empty(n, dtype=torch.uint8, device=states.device)
    oc = torch.empty(n, dtype=torch.int8, device=states.device)
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_outcome(_ptr(states), n, _ptr(over), _ptr(oc), _stream(states)), "hz_outcome")
    return over, oc


def random_actions(states, out=None):
    """int16[n]: the uniform-random playout policy's action for each game (-1: none)."""
    lib = _lib.load()
    n = _check_states(states)
    out = torch.empty(n, dtype=torch.int16, device=states.device) if out is None else out
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_random_actions(_ptr(states), n, _ptr(out), _stream(states)), "hz_random_actions")
    return out


def greedy_actions(states, out=None):
    """int16[n]: the 1-ply greedy agent's action (choose_move_greedy, evaluation.py:137-196)."""
    lib = _lib.load()
    n = _check_states(states)
    out = torch.empty(n, dtype=torch.int16, device=states.device) if out is None else out
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_greedy_actions(_ptr(states), n, _ptr(out), _stream(states)), "hz_greedy_actions")
    return out


def playout(states, max_steps=1000, steps=None, total=None):
    """Fused random playout in place.  Returns (steps int32[n], total int64[1]) tensors."""
    lib = _lib.load()
    n = _check_states(states)
    steps = torch.empty(n, dtype=torch.int32, device=states.device) if steps is None else steps
    total = torch.zeros(1, dtype=torch.int64, device=states.device) if total is None else total
    with torch.cuda.device(states.device):
        _lib.check(
            lib.hz_playout(_ptr(states), n, max_steps, _ptr(steps), _ptr(total), _stream(states)), "hz_playout"
        )
    return steps, total


class HostPlayout:
    """Playouts on HOST buffers: ``run(states_in, states_out)`` takes pinned int32[n,32] host
    tensors, plays every game to the end on the GPU and returns the final records in
    ``states_out``.  The batch is cut into chunks that travel on separate CUDA streams, so the
    H2D copy of one chunk, the kernel of another and the D2H copy of a third overlap (PCIe is
    full duplex); device staging buffers are allocated once."""

    def __init__(self, n, device="cuda", chunks=4, max_steps=1000):
        self.n, self.max_steps = int(n), int(max_steps)
        self.device = torch.device(device)
        self.chunks = max(1, min(int(chunks), self.n))
        self.bounds = [self.n * c // self.chunks for c in range(self.chunks + 1)]
        self.dev_states = torch.empty((self.n, 32), dtype=torch.int32, device=self.device)
        self.steps = torch.empty(self.n, dtype=torch.int32, device=self.device)
        self.total = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.chunks)]

    def run(self, states_in, states_out):
        if not (states_in.is_pinned() and states_out.is_pinned()):
            raise ValueError("host buffers must be pinned")
        if states_in.shape != (self.n, 32) or states_out.shape != (self.n, 32) or states_in.dtype != torch.int32:
            raise TypeError("host buffers must be int32[n, 32]")
        main = torch.cuda.current_stream(self.device)
        for c, st in enumerate(self.streams):
            lo, hi = self.bounds[c], self.bounds[c + 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                d = self.dev_states[lo:hi]
                d.copy_(states_in[lo:hi], non_blocking=True)
                playout(d, self.max_steps, steps=self.steps[lo:hi], total=self.total)
                states_out[lo:hi].copy_(d, non_blocking=True)
        for st in self.streams:
            main.wait_stream(st)
        main.synchronize()
        return states_out

    def run_keys(self, keys_in, results_out):
        """Fresh games from their 64-bit keys: ``keys_in`` pinned int64[n] (one draw-stream key
        per game: HarmoniesGameState.__init__ is hz_init_states on the device), plays them to
        the end and fills ``results_out`` pinned int32[n, 3] with (meta word incl. winner,
        final_scores word, number of actions).  12 bytes back per game instead of 128."""
        if not (keys_in.is_pinned() and results_out.is_pinned()):
            raise ValueError("host buffers must be pinned")
        if keys_in.shape != (self.n,) or keys_in.dtype != torch.int64 or results_out.shape != (self.n, 3):
            raise TypeError("keys int64[n], results int32[n, 3]")
        lib = _lib.load()
        main = torch.cuda.current_stream(self.device)
        if not hasattr(self, "dev_keys"):
            self.dev_keys = torch.empty(self.n, dtype=torch.int64, device=self.device)
            self.cols = torch.tensor([22, 23, 27], device=self.device)
        self.dev_keys.copy_(keys_in, non_blocking=True)
        with torch.cuda.device(self.device):
            _lib.check(lib.hz_init_states(_ptr(self.dev_states), self.n, _ptr(self.dev_keys), 0, 0, main.cuda_stream), "hz_init_states")
        playout(self.dev_states, self.max_steps, steps=self.steps, total=self.total)
        results_out.copy_(self.dev_states.index_select(1, self.cols), non_blocking=True)
        main.synchronize()
        return results_out


def launch_count():
    return int(_lib.load().hz_launch_count())
