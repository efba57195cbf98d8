"""Developer first-light checks on a B200 (not part of the test-suite): python tools/gpu_check.py <what>"""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")


def ref_kernel(A, B, kid, kscale):
    A = A.double(); B = B.double()
    d2 = ((A * A).sum(1)[:, None] + (B * B).sum(1)[None, :] - 2 * A @ B.T).clamp_min(0)
    if kid == 0:
        return torch.exp(-kscale * d2)
    r = kscale * d2.sqrt()
    if kid == 1:
        return torch.exp(-r)
    if kid == 2:
        s = math.sqrt(3) * r
        return (1 + s) * torch.exp(-s)
    s = math.sqrt(5) * r
    return (1 + s + s * s / 3) * torch.exp(-s)


def relerr(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max())


def check_prep():
    for n, p in [(1000, 500), (257, 100), (5, 3), (4096, 3000), (300, 67)]:
        X = torch.randn(n, p, device=dev) * 2 + 1
        shift = X.mean(0).contiguous()
        scale = (torch.rand(p, device=dev) + 0.5).contiguous()
        P = nat.prep(X, shift, scale, 0.25)
        torch.cuda.synchronize()
        ref = ((X - shift) * scale * 0.25).half()
        ok_d = torch.equal(P.data[:, :p], ref)
        pad0 = bool((P.data[:, p:] == 0).all())
        nref = (ref.float() ** 2).sum(1)
        e = relerr(P.norms[:n], nref.double())
        print(f"prep n={n} p={p}: data_equal={ok_d} pad_zero={pad0} norm_relerr={e:.2e} "
              f"norm_pad_zero={bool((P.norms[n:] == 0).all())}")


def make_ops(a, b, p, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    A = torch.randn(a, p, device=dev, generator=g)
    B = torch.randn(b, p, device=dev, generator=g)
    return A, B


def check_kdense():
    for (a, b, p) in [(128, 256, 64), (128, 256, 128), (300, 500, 100), (1000, 700, 500), (2000, 2000, 1000)]:
        A, B = make_ops(a, b, p)
        sigma = math.sqrt(2 * p)
        PA = nat.prep(A, None, None, 1.0)
        PB = nat.prep(B, None, None, 1.0)
        for kid in range(4):
            kscale = 1 / (2 * sigma * sigma) if kid == 0 else 1 / sigma
            out = torch.full((a, b), float("nan"), device=dev)
            nat.kdense(kid, kscale, PA, PB, out)
            torch.cuda.synchronize()
            ref = ref_kernel(PA.data[:, :p].float(), PB.data[:, :p].float(), kid, kscale)
            ref_true = ref_kernel(A, B, kid, kscale)
            print(f"kdense a={a} b={b} p={p} kid={kid}: err_vs_rounded={relerr(out, ref):.2e} "
                  f"err_vs_fp32_inputs={relerr(out, ref_true):.2e} nan={int(torch.isnan(out).sum())}")
    # symmetric
    A, _ = make_ops(700, 1, 300)
    PA = nat.prep(A, None, None, 1.0)
    out = torch.empty(700, 700, device=dev)
    nat.kdense(1, 0.05, PA, PA, out, symmetric=True)
    torch.cuda.synchronize()
    print("kdense symmetric diag max dev from 1:", float((out.diagonal() - 1).abs().max()),
          "asym:", float((out - out.T).abs().max()))


def check_kmv():
    for (a, b, p, d) in [(128, 256, 64, 4), (300, 500, 100, 7), (2000, 500, 500, 10), (500, 2000, 500, 10),
                         (5000, 3000, 1000, 20), (1000, 40000, 256, 20), (40000, 1000, 256, 32)]:
        A, B = make_ops(a, b, p)
        sigma = math.sqrt(2 * p)
        PA = nat.prep(A, None, None, 1.0)
        PB = nat.prep(B, None, None, 1.0)
        dp = nat.round_up(d, 4)
        V = torch.zeros(b, dp, device=dev)
        V[:, :d] = torch.randn(b, d, device=dev)
        for kid in range(4):
            kscale = 1 / (2 * sigma * sigma) if kid == 0 else 1 / sigma
            out = torch.zeros(a, dp, device=dev)
            nat.kmv(kid, kscale, PA, PB, V, out)
            torch.cuda.synchronize()
            Kref = ref_kernel(PA.data[:, :p].float(), PB.data[:, :p].float(), kid, kscale)
            ref = Kref @ V.double()
            ref_true = ref_kernel(A, B, kid, kscale) @ V.double()
            print(f"kmv a={a} b={b} p={p} d={d} kid={kid}: err_vs_rounded={relerr(out, ref):.2e} "
                  f"err_vs_fp32_inputs={relerr(out, ref_true):.2e}")


def check_linalg():
    for M, dp in [(128, 4), (256, 8), (1024, 12), (2048, 20)]:
        g = torch.Generator(device=dev).manual_seed(1)
        Z = torch.randn(M, 64, device=dev, generator=g)
        K = torch.exp(-torch.cdist(Z.double(), Z.double()) / 8.0) + 1e-3 * torch.eye(M, device=dev, dtype=torch.float64)
        A = K.float().contiguous()
        invd = torch.empty(M // 128, 128, 128, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        nat.potrf(A, invd, info)
        torch.cuda.synchronize()
        Lref = torch.linalg.cholesky(K)
        L = torch.tril(A).double()
        print(f"potrf M={M}: info={int(info)} relerr L={relerr(L, Lref):.2e} "
              f"recon={float((L @ L.T - K).abs().max()):.2e}")
        G = torch.empty(M, M, device=dev)
        nat.ltl(A, G, 1.0 / M, 1e-3)
        torch.cuda.synchronize()
        Gref = Lref.T @ Lref / M + 1e-3 * torch.eye(M, device=dev, dtype=torch.float64)
        print(f"ltl M={M}: relerr={relerr(torch.tril(G), torch.tril(Gref)):.2e}")
        B = torch.randn(M, dp, device=dev, generator=g)
        F = nat.trsm_pack(A, invd)
        for tr in (False, True):
            X = B.clone()
            nat.trsm(F, X, tr)
            torch.cuda.synchronize()
            ref = torch.linalg.solve_triangular(Lref.T if tr else Lref, B.double(), upper=tr)
            print(f"trsm M={M} dp={dp} trans={tr}: relerr={relerr(X, ref):.2e}")
        P = torch.randn(M, dp, device=dev, generator=g); Q = torch.randn(M, dp, device=dev, generator=g)
        out = torch.empty(dp, device=dev)
        cgs = nat.CgScratch(dev if isinstance(dev, torch.device) else torch.device(dev), 2, dp)
        nat.coldot(P, Q, out, cgs)
        print(f"coldot relerr={relerr(out, (P.double() * Q.double()).sum(0)):.2e}")
        X = torch.randn(M, dp, device=dev, generator=g); R = torch.randn(M, dp, device=dev, generator=g)
        rs_old = torch.rand(dp, device=dev) + 0.5; pAp = torch.rand(dp, device=dev) + 0.5
        rs_new = torch.empty(dp, device=dev)
        X0, R0 = X.clone(), R.clone()
        nat.cg_update(P, Q, X, R, rs_old, pAp, 1e-7, rs_new, cgs)
        al = rs_old / (pAp + 1e-7)
        print(f"cg_update X={relerr(X, (X0 + P * al).double()):.2e} R={relerr(R, (R0 - Q * al).double()):.2e} "
              f"rs={relerr(rs_new, ((R0 - Q * al).double() ** 2).sum(0)):.2e}")
        P0 = P.clone()
        nat.cg_direction(R, P, rs_new, rs_old, 1e-7, cgs)
        print(f"cg_direction={relerr(P, (R + P0 * (rs_new / (rs_old + 1e-7))).double()):.2e}")
        Y = Q.clone()
        nat.axpby(0.5, P, 2.0, Y)
        print(f"axpby={relerr(Y, (0.5 * P + 2 * Q).double()):.2e}")


def bench_kmv():
    sms, smem = nat.device_info()
    print("device:", torch.cuda.get_device_name(0), "sms", sms, "smem", smem)
    for (a, b, p, d, tag) in [(100000, 10000, 2000, 10, "C2 fwd"), (10000, 100000, 2000, 10, "C2 trans"),
                              (148 * 128 * 4, 50000, 3000, 20, "C3 fwd slice"),
                              (50000, 148 * 128 * 4, 3000, 20, "C3 trans slice")]:
        A = torch.randn(a, p, device=dev) * 0.2
        B = torch.randn(b, p, device=dev) * 0.2
        PA = nat.prep(A, None, None, 1.0); PB = nat.prep(B, None, None, 1.0)
        del A, B
        dp = nat.round_up(d, 4)
        V = torch.randn(b, dp, device=dev)
        out = torch.zeros(a, dp, device=dev)
        for kid in (1, 2):
            for _ in range(2):
                nat.kmv(kid, 0.1, PA, PB, V, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                nat.kmv(kid, 0.1, PA, PB, V, out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            fl = 2.0 * a * b * (p + d)
            print(f"{tag} kid={kid} a={a} b={b} p={p} d={d}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s "
                  f"(stages={os.environ.get('SAB_STAGES', 'auto')} vbufs={os.environ.get('SAB_VBUFS', 'auto')})")
        del PA, PB, V, out
        torch.cuda.empty_cache()


def bench_linalg():
    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    for M in [int(x) for x in os.environ.get("BENCH_M", "2048,16384,32768,50048").split(",")]:
        g = torch.Generator(device=dev).manual_seed(1)
        Z = torch.randn(M, 64, device=dev, generator=g)
        A = torch.cdist(Z, Z)
        A.mul_(-1.0 / 8.0).exp_()
        A.diagonal().add_(1e-5 * M)
        del Z
        invd = torch.empty(M // 128, 128, 128, device=dev)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        t1 = timed(lambda: nat.potrf(A, invd, info))
        G = torch.empty(M, M, device=dev)
        t2 = timed(lambda: nat.ltl(A, G, 1.0 / M, 1e-6))
        info2 = torch.zeros(1, dtype=torch.int32, device=dev)
        invd2 = torch.empty(M // 128, 128, 128, device=dev)
        t5 = timed(lambda: nat.potrf(G, invd2, info2))
        B = torch.randn(M, 20, device=dev)
        B0 = B.clone()
        holder = {}
        t6 = timed(lambda: holder.__setitem__("F", nat.trsm_pack(A, invd)))
        F = holder["F"]
        F2 = nat.trsm_pack(G, invd2)
        B2 = torch.empty_like(B)
        def best(fn, reps=5):
            fn()
            return min(timed(fn) for _ in range(reps))

        t3 = best(lambda: nat.trsm(F, B.copy_(B0), False))
        t4 = best(lambda: nat.trsm(F, B.copy_(B0), True))
        t7 = best(lambda: nat.trsm_pair(F, F2, B.copy_(B0), B2, None, 0.0, False))
        t8 = best(lambda: nat.trsm_pair(F, F2, B.copy_(B0), B2, None, 0.0, True))
        print(f"M={M}: potrf {t1*1e3:.1f} ms ({M**3/3/t1/1e12:.1f} TFLOP/s) ltl {t2*1e3:.1f} ms "
              f"({M**3/3/t2/1e12:.1f} TFLOP/s) potrf(G) {t5*1e3:.1f} ms pack {t6*1e3:.1f} ms trsm fwd {t3*1e3:.2f} ms "
              f"({M*M*2/t3/1e9:.0f} GB/s) bwd {t4*1e3:.2f} ms pair fwd {t7*1e3:.2f} ms ({M*M*4/t7/1e9:.0f} GB/s) "
              f"bwd {t8*1e3:.2f} ms info={int(info)},{int(info2)}", flush=True)
        del A, G, invd, invd2, B, F, F2


def check_accum():
    """How far is |a|^2+|a|^2-2a.a from zero (relative to |a|^2+|a|^2)?  tcgen05 accumulation resolution."""
    import numpy as np
    nat.set_option("SAB_SNAP", "-1")   # disable the snap: look at the raw computed distances
    from sobolev_alignment_b200 import synthetic as syn
    for p in (64, 256, 512, 1024, 2048, 3000, 4096):
        for name, A in (("normal", torch.randn(1024, p, device=dev)),
                        ("positive", torch.rand(1024, p, device=dev) + 0.5),
                        ("logcounts_centred", None)):
            if A is None:
                Xn = torch.from_numpy(syn.log_counts_numpy(1024, p, seed=p)).to(dev)
                A = (Xn - Xn.mean(0)) * 8.0
            PA = nat.prep(A.contiguous(), None, None, 1.0)
            out = torch.empty(1024, 1024, device=dev)
            # gaussian with kscale 1: K = exp(-d2)  =>  d2 = -log K on the diagonal
            nat.kdense(0, 1.0, PA, PA, out)
            torch.cuda.synchronize()
            d2 = -torch.log(out.diagonal().double())
            nn = 2 * PA.norms[:1024].double()
            ratio = (d2 / nn)
            print(f"accum p={p:5d} {name:18s}: d2/nn  max={float(ratio.max()):.3e} mean={float(ratio.mean()):.3e} "
                  f"min={float(ratio.min()):.3e}  per-MMA-step={float(ratio.max()) / (nat.round_up(p, 64) / 16):.3e}")


if __name__ == "__main__":
    what = sys.argv[1]
    t0 = time.time()
    {"prep": check_prep, "kdense": check_kdense, "kmv": check_kmv, "linalg": check_linalg,
     "bench_kmv": bench_kmv, "bench_linalg": bench_linalg, "accum": check_accum}[what]()
    print(f"[{what}] done in {time.time() - t0:.1f}s")
