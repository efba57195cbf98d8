#!/usr/bin/env python
"""
bench.py -- KRR (Falkon-path) fit benchmark on B200, BASELINE.json metric:
"KRR fit seconds at n=1M, M=50k; fused K_nM.v TFLOP/s vs peak, 1/2/4/8 B200".

A step = one complete `Falkon.fit` (centre selection, operand preparation, Nystrom preconditioner,
RHS, 20 preconditioned-CG iterations with the periodic residual recomputation, alpha) on BASELINE
config 3: n = 1e6 artificial log10-count profiles, p = 3000 genes, M = 50k centres, d = 20 latents,
Laplacian kernel sigma = 10, penalization 1e-6.  With N GPUs the n rows are sharded (strong scaling),
centres / preconditioner / CG state replicated, one NCCL all-reduce of the M x d partial sums per
K^T(K v) product.

  value   fit seconds with X, Y already resident in HBM (fp32) when the timed region starts
  e2e     the same fit through the reference-facing API (`KRRApprox(method="falkon").fit`) with X, Y
          in pinned HOST memory: H2D staging of the inputs and D2H of alpha inside the timed region
  roofline  the fused `kmv` kernel: algorithmic flops (2ab(p+d)) / CUDA-event time of every launch
            inside the timed fits, against the measured bf16 tensor peak
  cpu_baseline  float32 CPU port of the same algorithm (oracle/falkon_oracle.py) timed on this box's
            host cores on a bounded sample and extrapolated to the full fit

`--impl reference` times that CPU port as the reference arm (the reference's own Falkon path cannot
be installed here: `falkon` is un-vendored, un-pinned and there is no network).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (n, p, M, d, kernel, kernel_params, penalization)
    "C3": (1_000_000, 3000, 50_000, 20, "laplacian", {"sigma": 10.0}, 1e-6),
    "C2": (100_000, 2000, 10_000, 10, "laplacian", {"sigma": 10.0}, 1e-6),
    "C1": (2_000, 500, 500, 10, "laplacian", {"sigma": 10.0}, 1e-3),
    "C4": (2_000_000, 3000, 100_000, 20, "matern", {"sigma": 10.0, "nu": 1.5}, 1e-6),
}
MAXITER = 20


def kmv_traffic(cfg_name, world):
    """
    dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernel, from the committed ncu
    capture (profiles/kmv_traffic.json: {"<workload>@<gpus>": {"bytes_per_launch": ..., "source": ...}}); None
    when no capture exists for this workload -- the number is never made up in this file.
    """
    path = os.path.join(ROOT, "profiles", "kmv_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        ent = json.load(f).get(f"{cfg_name}@{world}")
    return (ent["bytes_per_launch"], ent["source"]) if ent else (None, None)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_sustained": d["bf16_tflops_sustained"], "bf16_burst": d["bf16_tflops"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def reference_arm(args, cfg_name):
    """CPU port of the same path on the host cores, bounded sample, extrapolated to the full fit."""
    from oracle import falkon_oracle as fo
    from sobolev_alignment_b200 import synthetic as syn

    n, p, M, d, kernel, kparams, lam = WORKLOADS[cfg_name]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sigma, nu = kparams["sigma"], kparams.get("nu")
    full = n * M * p <= 4e12 and not args.force_sample
    # sample sized for ~10-30 s of CPU work per step at ~0.5-1 TFLOP/s
    ns = n if full else int(min(n, max(1024, 8e12 / (2.0 * M * p))))
    ns = max(256, ns // 256 * 256) if not full else n
    X = torch.from_numpy(syn.log_counts_numpy(ns, p, seed=0))
    rng = np.random.default_rng(2)
    C = torch.from_numpy(syn.log_counts_numpy(M, p, seed=3)) if not full else X[np.sort(rng.choice(n, M, replace=False))]
    v = torch.randn(M, d)
    times = []
    if full:
        Y = torch.from_numpy(syn.targets_numpy(X.numpy(), d, seed=1))
        for i in range(args.warmup + args.steps):
            t = time.perf_counter()
            fo.falkon_fit_f32(X, Y, C, kernel, sigma, lam, nu=nu, maxiter=MAXITER)
            if i >= args.warmup:
                times.append(time.perf_counter() - t)
        fit_s = float(np.mean(times))
        sample = f"complete float32 CPU fit (n={n}, all rows), {len(times)} timed runs"
    else:
        for i in range(min(args.warmup, 1) + args.steps):
            t = time.perf_counter()
            fo.dmmv_pass_f32(X, C, v, None, kernel, sigma, nu)
            if i >= min(args.warmup, 1):
                times.append(time.perf_counter() - t)
        t_pass = float(np.mean(times))
        flops_pass = 2.0 * ns * M * p + 4.0 * ns * M * d
        rate = flops_pass / t_pass
        n_pass = 1 + MAXITER + MAXITER // 10
        precond = (2.0 * M * M * p + M ** 3) / rate
        fit_s = t_pass * (n / ns) * n_pass + precond
        sample = (f"one float32 K^T(K v) pass over {ns} of {n} rows (full M={M}, p={p}): {t_pass:.2f} s = "
                  f"{rate / 1e12:.2f} TFLOP/s; extrapolated x(n/ns) x {n_pass} passes + preconditioner flops "
                  f"(2M^2p + M^3) at that rate")
    line = {
        "impl": "reference", "metric": f"krr_fit_seconds_{cfg_name}", "value": fit_s, "unit": "s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": fit_s * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg_name, args.gpus),
        "cpu_baseline": {"value": fit_s, "unit": "s", "cores": threads, "kind": "port", "sample": sample,
                         "extrapolated": not full},
        "e2e": {"value": fit_s, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def workload_config(cfg_name, gpus):
    n, p, M, d, kernel, kparams, lam = WORKLOADS[cfg_name]
    return {"workload": f"BASELINE {cfg_name}: KRRApprox Falkon-path fit, n={n}, p={p}, M={M}, d={d}, "
                        f"{kernel} {kparams}, penalization={lam}, maxiter={MAXITER}",
            "n": n, "p": p, "M": M, "d": d, "kernel": kernel, "row_shards": gpus,
            "l2": "inputs larger than L2 (X fp16 6 GB, centres 300 MB at C3)"}


# --------------------------------------------------------------------------------------------------
# secondary legs: extra keys of the one JSON line (the C3 fit stays the headline); every leg is bounded
# and fenced by try/except so that it can never cost the headline its line
def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def leg_c2(dev, args):
    """BASELINE config 2 (n=1e5, p=2000, M=1e4, d=10) fit, device-resident, + the COMPLETE float32 CPU fit beside it."""
    from oracle import falkon_oracle as fo
    from sobolev_alignment_b200 import KRRApprox, _native as nat, synthetic as syn
    from sobolev_alignment_b200.falkon import Falkon, FalkonOptions

    n, p, M, d, kernel, kparams, lam = WORKLOADS["C2"]
    X = syn.log_counts_torch(n, p, seed=2000, device=dev)
    Y = syn.targets_torch(X, d, seed=7)
    kobj = KRRApprox.falkon_kernel[kernel](**kparams)

    def fit():
        return Falkon(kernel=kobj, penalty=lam, M=M, maxiter=MAXITER, options=FalkonOptions(center_seed=11)).fit(X, Y)

    for _ in range(2):
        fit()
    nat.STATS.reset()
    nat.STATS.time_kmv = True
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    reps = 5
    for _ in range(reps):
        est = fit()
    e1.record()
    torch.cuda.synchronize()
    nat.STATS.time_kmv = False
    fl, ks, kn = nat.STATS.kmv_summary()
    out = {"fit_s": e0.elapsed_time(e1) / reps / 1e3, "kmv_tflops": fl / ks / 1e12 if ks else None,
           "phases_s": est.solver_stats_["times"], "workload": workload_config("C2", 1)["workload"]}
    if not args.no_cpu_baseline:
        Xc, Yc = X.cpu(), Y.cpu()
        idx = np.sort(np.random.default_rng(11).choice(n, size=M, replace=False))
        torch.set_num_threads(os.cpu_count() or 1)
        t = time.perf_counter()
        fo.falkon_fit_f32(Xc, Yc, Xc[idx], kernel, kparams["sigma"], lam, maxiter=MAXITER)
        out["cpu_fit_s"] = time.perf_counter() - t
        out["cpu"] = f"complete float32 CPU port fit, all {n} rows, {os.cpu_count()} cores, not extrapolated"
    return out


def leg_transform(est, Xh, label):
    """`transform` throughput (row a9): K(X, C) alpha for host rows, staging inside the timed region."""
    est.predict(Xh[:131072])                       # warm-up: predictor (prepared centres), staging buffers
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = est.predict(Xh)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    n, p = Xh.shape
    M, d = est.alpha_.shape
    return {"rows_per_s": n / dt, "seconds": dt, "n": n, "M": M, "d": d, "input": label,
            "tflops_incl_staging": 2.0 * n * M * (p + d) / dt / 1e12, "h2d_bytes": n * p * 4, "d2h_bytes": int(out.numel() * 4)}


def leg_memmap_e2e(one_fit, Xh, Yh, est_for_transform):
    """
    The e2e fit with the input the reference REALLY hands over: torch.from_numpy of a read-only np.memmap
    (pageable, file-backed; sobolev_alignment.py:613-616,922-925) -- the backend has to stage it through its own
    pinned buffers.  The file lives on /dev/shm when that has room (page-cache speed), else in the temp dir.
    """
    import shutil
    import tempfile
    import warnings

    n, p = Xh.shape
    need = n * p * 4 + (1 << 30)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else tempfile.gettempdir()
    if shutil.disk_usage(base).free < need:
        return {"skipped": f"no room for a {need >> 30} GiB memmap under {base}"}
    d = tempfile.mkdtemp(prefix="sab_bench_", dir=base)
    try:
        path = os.path.join(d, "X.npy")
        mm = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(n, p))
        step = 1 << 16
        for s0 in range(0, n, step):
            mm[s0: s0 + step] = Xh[s0: s0 + step].numpy()
        mm.flush()
        del mm
        ro = np.load(path, mmap_mode="r")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)       # "non-writable array": it is never written
            Xm = torch.from_numpy(ro)
        Ym = Yh.clone()                                        # pageable
        e0, e1 = _events()
        torch.cuda.synchronize()
        t = time.perf_counter()
        e0.record()
        est = one_fit(Xm, Ym, through_krr_approx=True)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t
        out = {"value": e0.elapsed_time(e1) / 1e3, "unit": "s", "wall_s": wall, "input": f"read-only np.memmap under {base}",
               "phases_s": est.solver_stats_["times"], "h2d_bytes_per_step": int(n * p * 4 + Ym.numel() * 4)}
        out["transform"] = leg_transform(est_for_transform, Xm, f"read-only np.memmap under {base}")
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def leg_c5(dev):
    """
    BASELINE config 5: source + target KRR fits with M = 50k centres each, then the cross-kernel products
    M_X, M_Y, M_XY (alpha^T K beta through the fused kernel, no M x M matrix) and the principal-vector cosines.
    n = 250 000 artificial samples per side (BASELINE.json does not fix n for this config), p = 3000, d = 20.
    """
    from sobolev_alignment_b200 import KRRApprox, synthetic as syn
    from sobolev_alignment_b200.alignment import EncoderComparison

    n, p, M, d = 250_000, 3000, 50_000, 20
    krr, fit_s = {}, {}
    for i, side in enumerate(("source", "target")):
        X = syn.log_counts_torch(n, p, seed=3000 + i, device=dev)
        Y = syn.targets_torch(X, d, seed=7)
        clf = KRRApprox(method="falkon", kernel="laplacian", kernel_params={"sigma": 10.0}, M=M, penalization=1e-6,
                        maxiter=MAXITER, falkon_options={"center_seed": 21 + i})
        e0, e1 = _events()
        e0.record()
        clf.fit(X, Y)
        e1.record()
        torch.cuda.synchronize()
        fit_s[side] = e0.elapsed_time(e1) / 1e3
        krr[side] = clf
        del X, Y
    torch.cuda.synchronize()
    t = time.perf_counter()
    cmp_ = EncoderComparison(krr["source"], krr["target"]).compare().principal_vectors()
    torch.cuda.synchronize()
    align_s = time.perf_counter() - t
    return {"fit_s": fit_s, "cross_kernels_and_cosines_s": align_s, "total_s": sum(fit_s.values()) + align_s,
            "top_principal_cosines": [float(x) for x in cmp_.principal_angles[:5]],
            "workload": f"two KRRApprox Falkon-path fits (n={n}, p={p}, M={M}, d={d}, laplacian sigma 10) + M_X/M_Y/M_XY + SVD"}


def leg_grid(dev, single_fit_s):
    """
    The model-selection grid of krr_model_selection.py (4 nu x 11 penalizations, :24-52,260-266) at config-2 size:
    KRRGrid (operands prepared once, T per nu, A per lambda, multi-lambda CG) against 44 independent fits.
    """
    from sobolev_alignment_b200 import synthetic as syn
    from sobolev_alignment_b200.model_selection import KRRGrid

    n, p, M, d, kernel, kparams, lam = WORKLOADS["C2"]
    X = syn.log_counts_torch(n, p, seed=2000, device=dev)
    Y = syn.targets_torch(X, d, seed=7)
    torch.cuda.synchronize()
    t = time.perf_counter()
    grid = KRRGrid(X, Y, M=M, sigma=10.0, kernel="matern", center_seed=11)
    out = grid.fit(nus=(0.5, 1.5, 2.5, float("inf")), penalizations=tuple(np.logspace(-8, 2, 11)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    return {"grid_s": dt, "points": len(out), "independent_fits_s": None if single_fit_s is None else 44 * single_fit_s,
            "per_kernel": {str(k): v for k, v in grid.times["kernels"].items()}, "prepare_s": grid.times["prepare"],
            "workload": f"4 nu x 11 penalizations, matern sigma 10, n={n}, p={p}, M={M}, d={d} (device-resident inputs)"}


def leg_baseline_b():
    """BASELINE.md Baseline B: the reference's OWN KRRApprox(method='sklearn') at config 1 (exact KRR, CPU)."""
    import importlib.util

    from sobolev_alignment_b200 import synthetic as syn

    path = os.path.join(ROOT, "baseline", "_ref", "sobolev_alignment", "krr_approx.py")
    if not os.path.exists(path):
        return {"skipped": "unmodified reference not installed (oracle/install_reference.sh)"}
    spec = importlib.util.spec_from_file_location("ref_krr_approx_bench", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n, p, M, d, kernel, kparams, lam = WORKLOADS["C1"]
    Xn = syn.log_counts_numpy(n, p, seed=0)
    X, Y = torch.from_numpy(Xn), torch.from_numpy(syn.targets_numpy(Xn, d, seed=1, linear=True))
    times = []
    for _ in range(3):
        clf = mod.KRRApprox(method="sklearn", kernel="matern", kernel_params={"length_scale": 10.0, "nu": 0.5},
                            penalization=lam)
        t = time.perf_counter()
        clf.fit(X, Y)
        times.append(time.perf_counter() - t)
    return {"fit_s": float(np.median(times)), "what": "reference krr_approx.py KRRApprox(method='sklearn', kernel='matern', "
            f"nu=0.5 (= L2 Laplacian)) fit at C1 (n={n}, p={p}, d={d}), {os.cpu_count()} cores"}


def leg_c1(dev):
    """BASELINE config 1 on the GPU through the same API (n=2000, p=500, M=500, d=10): launch-latency bound."""
    from sobolev_alignment_b200 import KRRApprox, synthetic as syn

    n, p, M, d, kernel, kparams, lam = WORKLOADS["C1"]
    Xn = syn.log_counts_numpy(n, p, seed=0)
    X, Y = torch.from_numpy(Xn), torch.from_numpy(syn.targets_numpy(Xn, d, seed=1, linear=True))
    times = []
    for _ in range(4):
        clf = KRRApprox(method="falkon", kernel=kernel, kernel_params=kparams, M=M, penalization=lam, maxiter=MAXITER)
        torch.cuda.synchronize()
        t = time.perf_counter()
        clf.fit(X, Y)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t)
    return {"fit_s": float(np.median(times[1:])), "what": "KRRApprox(method='falkon') fit at C1, host inputs, wall clock"}


def leg_c4(dev, rank, world, timed, barrier):
    """BASELINE config 4: Matern nu=1.5, n=2M, M=100k, d=20 on 8 GPUs (row shards + per-iteration all-reduce)."""
    from sobolev_alignment_b200 import KRRApprox, _native as nat, synthetic as syn
    from sobolev_alignment_b200.falkon import Falkon, FalkonOptions

    n, p, M, d, kernel, kparams, lam = WORKLOADS["C4"]
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    n_local = int(bounds[rank + 1] - bounds[rank])
    X = syn.log_counts_torch(n_local, p, seed=4000 + rank, device=dev)
    Y = syn.targets_torch(X, d, seed=7)
    kobj = KRRApprox.falkon_kernel[kernel](**kparams)

    def fit():
        return Falkon(kernel=kobj, penalty=lam, M=M, maxiter=MAXITER,
                      options=FalkonOptions(row_sharded=world > 1, center_seed=11)).fit(X, Y)

    fit()                                               # warm-up
    nat.STATS.reset()
    nat.STATS.time_kmv = True
    ms, wall, est = timed(fit, 2)
    nat.STATS.time_kmv = False
    fl, ks, kn = nat.STATS.kmv_summary()
    tb, ts, tn = nat.STATS.trsm_summary()
    return {"fit_s": ms / 1e3, "steps": 2, "kmv_tflops_per_gpu": fl / ks / 1e12 if ks else None,
            "trsm_gbs": tb / ts / 1e9 if ts else None, "phases_s": est.solver_stats_["times"],
            "final_residual": est.solver_stats_["residuals"][-1], "workload": workload_config("C4", world)["workload"]}


def _guard(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as exc:  # noqa: BLE001 - a secondary leg must not cost the headline its line
        return {"error": f"{type(exc).__name__}: {exc}"[:400]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SAB_BENCH_WORKLOAD", "C3"), choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-sample", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (C1/C2/C4/C5, transform, memmap)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.workload

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(reference_arm(args, cfg_name)), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")

    import torch.distributed as dist

    from sobolev_alignment_b200 import KRRApprox, _native as nat, synthetic as syn
    from sobolev_alignment_b200.falkon import Falkon, FalkonOptions

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, p, M, d, kernel, kparams, lam = WORKLOADS[cfg_name]
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    n_local = int(bounds[rank + 1] - bounds[rank])

    # synthetic inputs generated on the device from the seed (12 GB at C3: never shipped as files)
    X = syn.log_counts_torch(n_local, p, seed=1000 + rank, device=dev)
    Y = syn.targets_torch(X, d, seed=7)
    kernel_obj = KRRApprox.falkon_kernel[kernel](**kparams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_fit(Xin, Yin, through_krr_approx=False, **extra):
        opts = dict(row_sharded=world > 1, center_seed=11, **extra)
        if through_krr_approx:
            clf = KRRApprox(method="falkon", kernel=kernel, kernel_params=kparams, M=M, penalization=lam,
                            maxiter=MAXITER, falkon_options=opts)
            clf.fit(Xin, Yin)
            return clf.ridge_clf_
        est = Falkon(kernel=kernel_obj, penalty=lam, M=M, maxiter=MAXITER, options=FalkonOptions(**opts))
        est.fit(Xin, Yin)
        return est

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        last = None
        for _ in range(steps):
            last = fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps, float(t[1]) / steps, last

    # ---------------- device-resident arm
    for _ in range(args.warmup):
        one_fit(X, Y)
    nat.STATS.reset()
    nat.STATS.time_kmv = True
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, wall_ms, est = timed(lambda: one_fit(X, Y), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    nat.STATS.time_kmv = False
    launches = nat.STATS.launches // max(1, args.steps)
    kmv_flops, kmv_s, kmv_n = nat.STATS.kmv_summary()
    trsm_bytes, trsm_s, trsm_n = nat.STATS.trsm_summary()
    stats = est.solver_stats_

    # ---------------- secondary: the same fit with stored kernel blocks (design B; K formed once per product pair)
    stored = None
    if not args.no_extras and cfg_name == "C3":
        def _stored():
            one_fit(X, Y, kernel_blocks="store")
            nat.STATS.reset()
            nat.STATS.time_kmv = True
            ms_b, _, est_b = timed(lambda: one_fit(X, Y, kernel_blocks="store"), 2)
            nat.STATS.time_kmv = False
            fl, ks, kn = nat.STATS.kmv_summary()
            by, bs, bn = nat.STATS.ktv_summary()
            return {"fit_s": ms_b / 1e3, "steps": 2, "kmv_tflops": fl / ks / 1e12 if ks else None, "kmv_launches_timed": kn,
                    "ktv_gbs": by / bs / 1e9 if bs else None, "ktv_seconds_per_fit": bs / 2, "kmv_seconds_per_fit": ks / 2,
                    "phases_s": est_b.solver_stats_["times"], "final_residual": est_b.solver_stats_["residuals"][-1],
                    "what": "FalkonOptions(kernel_blocks='store'): K formed once per CG product pair, row blocks left behind as "
                            "packed fp16 hi/lo tiles (transient scratch <= 24 GB), K^T t streamed back (sa_ktv)"}
        stored = _guard(_stored)

    # ---------------- end-to-end arm: pinned host inputs through KRRApprox.fit
    e2e = None
    extras_host = {}
    if not args.no_e2e:
        Xh = torch.empty(n_local, p, dtype=torch.float32).pin_memory()
        Yh = torch.empty(n_local, d, dtype=torch.float32).pin_memory()
        Xh.copy_(X)
        Yh.copy_(Y)
        X = Y = None
        torch.cuda.empty_cache()
        one_fit(Xh, Yh, through_krr_approx=True)   # warm-up (pinned staging buffers, allocator)
        e2e_ms, e2e_wall, est2 = timed(lambda: one_fit(Xh, Yh, through_krr_approx=True), max(1, min(args.steps, 2)))
        h2d = n_local * p * 4 + n_local * d * 4 + est2.ny_points_.numel() * 4
        d2h = est2.alpha_.numel() * 4
        e2e = {"value": e2e_ms / 1e3, "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "KRRApprox(method='falkon').fit(X_host_pinned, Y_host_pinned)",
               "stage_inputs_s": est2.solver_stats_["times"].get("stage_inputs"),
               "phases_s": est2.solver_stats_["times"], "host_overhead_s": e2e_wall / 1e3 - est2.solver_stats_["times"]["fit"]}
        extras_host = {}
        if not args.no_extras and world == 1 and cfg_name == "C3":
            extras_host["transform"] = _guard(leg_transform, est2, Xh, "pinned host tensor")
            extras_host["e2e_memmap"] = _guard(leg_memmap_e2e, one_fit, Xh, Yh, est2)
        del Xh, Yh
        torch.cuda.empty_cache()

    c4 = None
    if not args.no_extras and world == 8 and cfg_name == "C3":
        X = Y = est = None
        torch.cuda.empty_cache()
        c4 = _guard(leg_c4, dev, rank, world, timed, barrier)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = measured_peaks()
    kmv_tflops = kmv_flops / kmv_s / 1e12 if kmv_s > 0 else 0.0
    line = {
        "metric": f"krr_fit_seconds_{cfg_name}", "value": ms_step / 1e3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
        "config": workload_config(cfg_name, world),
        "roofline": {"bound": "tensor", "kernel": "kmv_kernel (fused tcgen05 K(A,B)@V)", "achieved": kmv_tflops,
                     "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": kmv_tflops / peaks["bf16_sustained"], "traffic": kmv_traffic(cfg_name, world)[0],
                     "traffic_source": kmv_traffic(cfg_name, world)[1],
                     "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                     "launches_timed": kmv_n, "kmv_seconds_per_fit": kmv_s / max(1, args.steps),
                     "algorithmic_flops_per_fit": kmv_flops / max(1, args.steps)},
        "trsm": {"bound": "hbm", "achieved": trsm_bytes / trsm_s / 1e9 if trsm_s > 0 else None,
                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "seconds_per_fit": trsm_s / max(1, args.steps),
                 "solves_timed": trsm_n},
        "phases_s": stats["times"], "cg_iterations": stats["n_iter"], "kmv_launches_per_fit": stats["n_kmv"],
        "final_residual": stats["residuals"][-1] if stats["residuals"] else None,
        "wall_ms_per_step": wall_ms, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
    }

    if e2e is not None and "e2e_memmap" in extras_host:
        e2e["memmap"] = extras_host["e2e_memmap"]
    if "transform" in extras_host:
        line["transform"] = extras_host["transform"]
    if c4 is not None:
        line["c4"] = c4
    if stored is not None:
        line["stored_kernel_blocks"] = stored
    if not args.no_extras and world == 1 and cfg_name == "C3":
        torch.cuda.empty_cache()
        line["c2"] = _guard(leg_c2, dev, args)
        line["grid"] = _guard(leg_grid, dev, line["c2"].get("fit_s"))
        line["c5"] = _guard(leg_c5, dev)
        line["c1"] = _guard(leg_c1, dev)
        if not args.no_cpu_baseline:
            line["c1"]["reference_sklearn"] = _guard(leg_baseline_b)
    if not args.no_cpu_baseline:
        ref = reference_arm(argparse.Namespace(gpus=world, steps=1, warmup=1, force_sample=args.force_sample), cfg_name)
        line["cpu_baseline"] = ref["cpu_baseline"]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
