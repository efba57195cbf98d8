"""Sustained-throughput sweep of the fused kernel over its launch-time knobs (library options).

python tools/sweep_kmv.py [seconds_per_config]
Each configuration runs back to back for ~seconds_per_config so the power-capped clock settles; the SM
clock is sampled with nvidia-smi in the middle of the run.
"""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sobolev_alignment_b200 import _native as nat  # noqa: E402

dev = torch.device("cuda:0")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.5
rows = 148 * 128 * 2
p, d = 3000, 20


def smi():
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        return out
    except Exception as e:  # noqa: BLE001
        return f"smi failed: {e}"


def run(mode, env):
    for k in ("SAB_STAGES", "SAB_CTA_GROUP", "SAB_DEBUG", "SAB_NSPLIT", "SAB_L2_BUDGET_MB", "SAB_VBUFS"):
        nat.set_option(k, None)
    for k, v in env.items():
        nat.set_option(k, v)        # (environment variables are read once per process: set the options explicitly)
    a, b = (rows, 50000) if mode == "fwd" else (50000, rows)
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(a, p, device=dev, generator=g) * 0.2
    B = torch.randn(b, p, device=dev, generator=g) * 0.2
    PA, PB = nat.prep(A, None, None, 1.0), nat.prep(B, None, None, 1.0)
    del A, B
    V = torch.randn(b, d, device=dev, generator=g)
    out = torch.zeros(a, d, device=dev)
    nat.kmv(1, 0.1, PA, PB, V, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    nat.kmv(1, 0.1, PA, PB, V, out)
    e1.record()
    torch.cuda.synchronize()
    one = e0.elapsed_time(e1)
    reps = max(3, int(secs * 1e3 / one))
    e0.record()
    for i in range(reps):
        nat.kmv(1, 0.1, PA, PB, V, out)
    e1.record()
    time.sleep(min(secs * 0.6, one * reps * 0.6e-3))
    clk = smi()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * a * b * (p + d) / ms / 1e9
    print(f"{mode} {env}: {ms:.2f} ms/launch  {tf:.0f} TFLOP/s  (burst {2.0 * a * b * (p + d) / one / 1e9:.0f})  clk,power={clk}", flush=True)


if __name__ == "__main__":
    configs = [
        {},
        {"SAB_DEBUG": 1},
        {"SAB_DEBUG": 2},
        {"SAB_DEBUG": 4},
        {"SAB_DEBUG": 6},
        {"SAB_DEBUG": 7},
        {"SAB_CTA_GROUP": 1, "SAB_DEBUG": 1},
        {"SAB_CTA_GROUP": 1, "SAB_DEBUG": 6},
        {"SAB_CTA_GROUP": 1, "SAB_DEBUG": 7},
    ]
    for mode in sys.argv[2:] or ("fwd", "tr"):
        for env in configs:
            run(mode, env)
