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


def end_turn(states, draws=None, status=None):
    """In-place _end_turn_actions (harmonies_engine.py:301-329) for the player to move of every state."""
    lib = _lib.load()
    n = _check_states(states)
    status = torch.empty(n, dtype=torch.uint8, device=states.device) if status is None else status
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_end_turn(_ptr(states), n, _ptr(draws), _ptr(status), _stream(states)), "hz_end_turn")
    return status


def replenish_piles(states):
    """In-place _replenish_piles (harmonies_engine.py:132-137)."""
    lib = _lib.load()
    n = _check_states(states)
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_replenish_piles(_ptr(states), n, _stream(states)), "hz_replenish_piles")


def draw_tiles(states, count):
    """In-place _draw_tiles(count) (harmonies_engine.py:120-130), count <= 15.  Returns uint8[n,16]:
    tile types in draw order, [:,15] = number drawn."""
    lib = _lib.load()
    n = _check_states(states)
    out = torch.zeros((n, 16), dtype=torch.uint8, device=states.device)
    with torch.cuda.device(states.device):
        _lib.check(lib.hz_draw_tiles(_ptr(states), n, int(count), _ptr(out), _stream(states)), "hz_draw_tiles")
    return out


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


def playout_keys(keys, results=None, total=None, max_steps=1000, n=None, seed=0, first_id=0, device=None):
    """Fresh games from their keys, played to the end in one launch (hz_playout_keys).
    ``keys`` int64[n] on the device (or None: keys derived from (seed, first_id + i), then ``n``
    and ``device`` are required).  Returns (results int32[n, 3], total int64[1]): per game the
    meta word (phase, winner code), the final-score word and the number of actions played."""
    lib = _lib.load()
    if keys is not None:
        if keys.dtype != torch.int64 or keys.dim() != 1 or not keys.is_cuda or not keys.is_contiguous():
            raise TypeError("keys must be a contiguous int64[n] CUDA tensor")
        n, device = keys.shape[0], keys.device
    elif n is None or device is None:
        raise ValueError("n and device are required when keys is None")
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    results = torch.empty((n, 3), dtype=torch.int32, device=device) if results is None else results
    if results.shape != (n, 3) or results.dtype != torch.int32 or not results.is_contiguous() or results.device != device:
        raise TypeError("results must be a contiguous int32[n, 3] tensor on the keys' device")
    total = torch.zeros(1, dtype=torch.int64, device=device) if total is None else total
    with torch.cuda.device(device):
        _lib.check(
            lib.hz_playout_keys(_ptr(keys) if keys is not None else None, n, seed, first_id, max_steps, _ptr(results),
                                _ptr(total), torch.cuda.current_stream(device).cuda_stream),
            "hz_playout_keys",
        )
    return results, total


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
        self._key_buffers(1)
        main = torch.cuda.current_stream(self.device)
        self.dev_keys[0].copy_(keys_in, non_blocking=True)
        playout_keys(self.dev_keys[0], self.dev_res[0], self.total, self.max_steps)
        results_out.copy_(self.dev_res[0], non_blocking=True)
        main.synchronize()
        return results_out

    def _key_buffers(self, depth):
        have = len(getattr(self, "dev_keys", []))
        if have < depth:
            self.dev_keys = getattr(self, "dev_keys", []) + [
                torch.empty(self.n, dtype=torch.int64, device=self.device) for _ in range(depth - have)]
            self.dev_res = getattr(self, "dev_res", []) + [
                torch.empty((self.n, 3), dtype=torch.int32, device=self.device) for _ in range(depth - have)]

    def run_keys_many(self, keys_batches, results_batches, depth=3):
        """A stream of ``run_keys`` batches, pipelined: while the kernel of batch k runs on the
        calling stream, the keys of batch k+1 travel host->device on a copy stream and the results
        of batch k-1 travel device->host on another (PCIe is full duplex), through ``depth``
        device buffer pairs.  Every batch is copied in and read back; the call returns when the
        last result is on the host.  ``results_batches[k]`` may alias an earlier entry once the
        caller has consumed it (at least ``depth`` batches later)."""
        self._check_batches(keys_batches, results_batches, (self.n,), torch.int64, (self.n, 3))
        depth = max(1, min(int(depth), len(keys_batches)))
        self._key_buffers(depth)
        return self._pipeline(
            keys_batches, results_batches, depth, self.dev_keys, self.dev_res,
            lambda slot: playout_keys(self.dev_keys[slot], self.dev_res[slot], self.total, self.max_steps))

    def run_many(self, states_batches, out_batches, depth=3):
        """``run`` for a stream of batches of 128-byte records, pipelined like ``run_keys_many``
        (in place on ``depth`` device staging buffers; PCIe-bound: 128 B each way per game)."""
        self._check_batches(states_batches, out_batches, (self.n, 32), torch.int32, (self.n, 32))
        depth = max(1, min(int(depth), len(states_batches)))
        if len(getattr(self, "dev_many", [])) < depth:
            self.dev_many = [self.dev_states] + [torch.empty_like(self.dev_states) for _ in range(depth - 1)]
        return self._pipeline(
            states_batches, out_batches, depth, self.dev_many, self.dev_many,
            lambda slot: playout(self.dev_many[slot], self.max_steps, steps=self.steps, total=self.total))

    def _check_batches(self, ins, outs, in_shape, in_dtype, out_shape):
        if len(ins) != len(outs):
            raise ValueError("one output buffer per input buffer")
        for a, b in zip(ins, outs):
            if not (a.is_pinned() and b.is_pinned()):
                raise ValueError("host buffers must be pinned")
            if a.shape != in_shape or a.dtype != in_dtype or b.shape != out_shape or b.dtype != torch.int32:
                raise TypeError(f"inputs {in_dtype}{list(in_shape)}, outputs int32{list(out_shape)}")

    def _pipeline(self, ins, outs, depth, dev_in, dev_out, launch):
        """The pipelined stream of batches.  A caller that streams through the SAME pinned buffers again
        and again (the normal case: a ring of host buffers) gets the whole pipeline — every copy, kernel
        and cross-stream dependency — replayed as ONE CUDA graph from the third call on (first call
        eager, second call captured): the host then issues one launch per call instead of ~10 API calls
        per batch, which is what keeps the end-to-end rate up when 8 ranks share one host's cores."""
        key = (id(launch.__code__), depth, tuple(a.data_ptr() for a in ins), tuple(b.data_ptr() for b in outs),
               tuple(t.data_ptr() for t in dev_in[:depth]))
        graphs = self.__dict__.setdefault("_graphs", {})
        g = graphs.get(key)
        if g is not None and g is not False:
            g.replay()
            torch.cuda.current_stream(self.device).synchronize()
            return outs
        if g is False and self.use_graphs:          # second call with these buffers: capture
            torch.cuda.synchronize(self.device)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self._pipeline_body(ins, outs, depth, dev_in, dev_out, launch)
            graphs[key] = gr
            gr.replay()
            torch.cuda.current_stream(self.device).synchronize()
            return outs
        graphs[key] = False
        self._pipeline_body(ins, outs, depth, dev_in, dev_out, launch)
        torch.cuda.current_stream(self.device).synchronize()
        return outs

    use_graphs = True

    def _pipeline_body(self, ins, outs, depth, dev_in, dev_out, launch):
        if not hasattr(self, "copy_in"):
            self.copy_in, self.copy_out = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        self.copy_in.wait_stream(main)
        self.copy_out.wait_stream(main)
        computed = [None] * depth      # event: the kernel that used slot's device buffers finished
        drained = [None] * depth       # event: slot's output buffer has been copied to the host
        for k, (a, b) in enumerate(zip(ins, outs)):
            slot = k % depth
            with torch.cuda.stream(self.copy_in):
                if computed[slot] is not None:
                    self.copy_in.wait_event(computed[slot])
                if dev_in is dev_out and drained[slot] is not None:
                    self.copy_in.wait_event(drained[slot])      # in-place buffers: wait for the read-back too
                dev_in[slot].copy_(a, non_blocking=True)
                arrived = torch.cuda.Event()
                arrived.record(self.copy_in)
            main.wait_event(arrived)
            if drained[slot] is not None:
                main.wait_event(drained[slot])
            launch(slot)
            computed[slot] = torch.cuda.Event()
            computed[slot].record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(computed[slot])
                b.copy_(dev_out[slot], non_blocking=True)
                drained[slot] = torch.cuda.Event()
                drained[slot].record(self.copy_out)
        main.wait_stream(self.copy_out)
        main.wait_stream(self.copy_in)
        return outs


def launch_count():
    return int(_lib.load().hz_launch_count())
