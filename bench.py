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
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernel (ncu --set full), mean of the
# forward (77.3 GB) and transposed (87.9 GB) C3 launches; the tensor-bound kernel re-reads its fp16 operands
# (6.3 GB) ~13 times through L2-sized windows, which is 5 % of the HBM bandwidth over the launch
KMV_TRAFFIC = {("C3", 1): 82.6e9}
KMV_TRAFFIC_SOURCE = "profiles/r01_kmv_v9_c3_ncu.txt + r01_kmv_v8_c3_*_ncu_full.txt (bytes per launch)"


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

    def one_fit(Xin, Yin, through_krr_approx=False):
        opts = dict(row_sharded=world > 1, center_seed=11)
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

    # ---------------- end-to-end arm: pinned host inputs through KRRApprox.fit
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty(n_local, p, dtype=torch.float32).pin_memory()
        Yh = torch.empty(n_local, d, dtype=torch.float32).pin_memory()
        Xh.copy_(X)
        Yh.copy_(Y)
        del X, Y
        torch.cuda.empty_cache()
        one_fit(Xh, Yh, through_krr_approx=True)   # warm-up (pinned staging buffers, allocator)
        e2e_ms, e2e_wall, est2 = timed(lambda: one_fit(Xh, Yh, through_krr_approx=True), max(1, min(args.steps, 2)))
        h2d = n_local * p * 4 + n_local * d * 4 + est2.ny_points_.numel() * 4
        d2h = est2.alpha_.numel() * 4
        e2e = {"value": e2e_ms / 1e3, "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "KRRApprox(method='falkon').fit(X_host_pinned, Y_host_pinned)",
               "stage_inputs_s": est2.solver_stats_["times"].get("stage_inputs"),
               "phases_s": est2.solver_stats_["times"], "host_overhead_s": e2e_wall / 1e3 - est2.solver_stats_["times"]["fit"]}

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
                     "frac": kmv_tflops / peaks["bf16_sustained"], "traffic": KMV_TRAFFIC.get((cfg_name, world)),
                     "traffic_source": KMV_TRAFFIC_SOURCE if (cfg_name, world) in KMV_TRAFFIC else None,
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

    if not args.no_cpu_baseline:
        ref = reference_arm(argparse.Namespace(gpus=world, steps=1, warmup=1, force_sample=args.force_sample), cfg_name)
        line["cpu_baseline"] = ref["cpu_baseline"]
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
