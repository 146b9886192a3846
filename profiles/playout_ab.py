"""k_playout timed alone (CUDA events, L2 flushed between waves): the bench's engine leg without the rest of bench.py.
HZ_LIB_PATH selects the build (A/B of kernel variants in one gpurun call).  Prints us per 65,536-game wave (median of 15)
and a checksum of all results (equal checksums = identical games)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from harmonies_alphazero_b200 import batched as hb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
keys = torch.randint(0, 2**62, (n,), dtype=torch.int64, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
res = torch.zeros(n, 3, dtype=torch.int32, device=dev)
total = torch.zeros(1, dtype=torch.int64, device=dev)
times = []
for it in range(20):
    flush.fill_(it & 255)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hb.playout_keys(keys, results=res, total=total)
    b.record()
    torch.cuda.synchronize()
    if it >= 5:
        times.append(a.elapsed_time(b) * 1e3)
times.sort()
t2 = []
chk2 = 0
for it in range(20):
    st = hb.init_states(n, seed=1000 + it)
    flush.fill_(it & 255)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hb.playout(st)
    b.record()
    torch.cuda.synchronize()
    chk2 ^= int(st.to(torch.int64).sum().item())
    if it >= 5:
        t2.append(a.elapsed_time(b) * 1e3)
t2.sort()
print(f"  records in HBM (hz_playout): median {t2[len(t2) // 2]:.2f} us  min {t2[0]:.2f} us  checksum {chk2}")
chk = int(res.to(torch.int64).sum().item()) ^ int(total.item())
print(f"{os.environ.get('HZ_LIB_PATH', 'default')}: median {times[len(times) // 2]:.2f} us  min {times[0]:.2f} us  steps/wave {int(total.item()) // 20}  checksum {chk}")
