"""configs[2] harness: 1 M synthetic end-game positions (full boards, stacks of height
~{1,1,2,3}, uniform tile types incl. unreachable stacks) through hz_legal_mask, hz_score,
hz_canon_hash and hz_encode.  Prints one JSON line with CUDA-event times and achieved
algorithmic GB/s; the timed launches sit between cudaProfilerStart/Stop for
`ncu --profile-from-start off`.

    python profiles/positions_1m.py [--n 1000000] [--iters 5] [--encode-n 262144]
"""

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb  # noqa: E402


def synth_positions(n, dev, seed=31337):
    g = torch.Generator(device=dev).manual_seed(seed)
    w = torch.zeros((n, 32), dtype=torch.int64, device=dev)
    shifts = torch.arange(23, device=dev, dtype=torch.int64)
    for p in range(2):
        h = torch.tensor([1, 1, 2, 3], device=dev)[torch.randint(0, 4, (n, 23), device=dev, generator=g)]
        for lvl in range(3):
            t = torch.randint(1, 7, (n, 23), device=dev, generator=g)
            code = torch.where(h > lvl, t, torch.zeros_like(t))
            for b in range(3):
                w[:, p * 9 + lvl * 3 + b] = (((code >> b) & 1) << shifts).sum(dim=1)
    hand = torch.randint(0, 6, (n, 3), device=dev, generator=g)
    w[:, 20] = (1 << (2 * hand)).sum(dim=1) << 16
    w[:, 21] = 0x05050505
    w[:, 22] = 0x0505 | ((torch.randint(0, 2, (n,), device=dev, generator=g) | 2) << 24)
    # int64 -> uint32 bit pattern -> int32
    return torch.where(w >= 2**31, w - 2**32, w).to(torch.int32).contiguous()


def timeit(fn, iters, flush):
    ms = []
    for _ in range(iters):
        flush.fill_(1)       # evict everything ...
        flush.max()          # ... then a read pass, so the timed kernel does not pay for write-backs of dirty flush lines
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sorted(ms)[len(ms) // 2]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--encode-n", type=int, default=262144)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    st = synth_positions(a.n, dev)
    n, m = a.n, min(a.encode_n, a.n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    mask = torch.empty((n, 5), dtype=torch.int32, device=dev)
    sub = st[:m].contiguous()
    b32 = torch.empty((m, 38, 5, 7), dtype=torch.float32, device=dev); g32 = torch.empty((m, 42), dtype=torch.float32, device=dev)
    b16 = torch.empty((m, 38, 5, 7), dtype=torch.bfloat16, device=dev).contiguous(memory_format=torch.channels_last)
    g16 = torch.empty((m, 42), dtype=torch.bfloat16, device=dev)
    runs = {
        "legal_mask": (lambda: hb.legal_mask(st, out=mask), n, 148),
        "score": (lambda: hb.score(st), n, 96),
        "canon_hash_ref": (lambda: hb.canon_hash(st, 1), n, 100),
        "encode_f32_nchw": (lambda: hb.encode(sub, board=b32, glob=g32), m, 128 + 5488),
        "encode_bf16_nhwc": (lambda: hb.encode(sub, dtype=torch.bfloat16, channels_last=True, board=b16, glob=g16), m, 128 + 2744),
    }
    for fn, _, _ in runs.values():
        fn()
    torch.cuda.synchronize()
    out = {}
    torch.cuda.profiler.start()
    for name, (fn, units, bpu) in runs.items():
        ms = timeit(fn, a.iters, flush)
        out[name] = {"ms": ms, "units_per_s": units / (ms * 1e-3), "algorithmic_GBps": units * bpu / (ms * 1e-3) / 1e9,
                     "frac_of_hbm_6455.6": units * bpu / (ms * 1e-3) / 1e9 / 6455.6, "bytes_per_unit": bpu, "units": units}
    torch.cuda.profiler.stop()
    # write-only reference: the encoders only store, and a pure store stream does not reach the
    # read+write copy figure of MEASURED_PEAKS.json; torch.fill_ of the same bytes is the yardstick
    for name, buf in (("fill_same_bytes_as_encode_f32", b32), ("fill_same_bytes_as_encode_bf16", b16)):
        ms = timeit(lambda: buf.fill_(1.0), a.iters, flush)
        nbytes = buf.numel() * buf.element_size()
        out[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9}
    print(json.dumps(out))
