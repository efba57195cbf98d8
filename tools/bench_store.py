"""Developer tool: timing of one K^T (K z) product pair, recomputed (2 x kmv) against stored blocks (kmv_store + ktv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
a = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 6
b, p, d = 50000, 3000, 20
A = torch.randn(a, p, device=dev) * 0.2
B = torch.randn(b, p, device=dev) * 0.2
PA, PB = nat.prep(A, None, None, 1.0), nat.prep(B, None, None, 1.0)
del A, B
dp = 20
Z = torch.randn(b, dp, device=dev)
t = torch.zeros(a, dp, device=dev)
w = torch.zeros(b, dp, device=dev)
store = torch.empty(nat.kstore_bytes(a, b), dtype=torch.uint8, device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


fl = 2.0 * a * b * (p + dp)
t_fwd = timed(lambda: nat.kmv(1, 0.1, PA, PB, Z, t))
t_tr = timed(lambda: nat.kmv(1, 0.1, PB, PA, t, w))
t_st = timed(lambda: nat.kmv(1, 0.1, PA, PB, Z, t, kstore=store))
t_ktv = timed(lambda: nat.ktv(store, a, b, t, w))
print(f"a={a} b={b} p={p} d={d}: kmv fwd {t_fwd:.2f} ms ({fl / t_fwd / 1e9:.0f} TFLOP/s), kmv transposed {t_tr:.2f} ms "
      f"({fl / t_tr / 1e9:.0f}), kmv fwd + store {t_st:.2f} ms ({fl / t_st / 1e9:.0f} TFLOP/s, {4.0 * a * b / t_st / 1e6:.0f} GB/s "
      f"of stores), ktv {t_ktv:.2f} ms ({4.0 * a * b / t_ktv / 1e6:.0f} GB/s); pair: {t_fwd + t_tr:.2f} -> {t_st + t_ktv:.2f} ms")
for ch in (8, 16, 32, 64):
    nat.set_option("SAB_KTV_CHUNK", ch)
    print(f"  SAB_KTV_CHUNK={ch}: ktv {timed(lambda: nat.ktv(store, a, b, t, w)):.2f} ms")
