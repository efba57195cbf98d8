"""
Seeded synthetic workloads for tests and bench.

The reference ships no data; its own fixtures draw Poisson counts with per-gene rates
`randint(1, 25)` (/root/reference/tests/test_sobolev_alignment.py:24-28) and log-transform them
with `log10(x + 1)` (/root/reference/sobolev_alignment/sobolev_alignment.py:852).  Targets follow
the reference's KRR test (`Y = X W`, /root/reference/tests/test_krr_approx.py:38-39) at test scale
and a bounded non-linear map `tanh((X - mu) W / sqrt(p))` at benchmark scale.
"""

from __future__ import annotations

import numpy as np
import torch


def log_counts_numpy(n: int, p: int, seed: int = 0, rates=None) -> np.ndarray:
    """n x p float32 matrix of log10(Poisson + 1) artificial expression profiles (CPU, numpy)."""
    rng = np.random.default_rng(seed)
    if rates is None:
        rates = np.random.default_rng(0).integers(1, 25, size=p).astype(np.float64)
    lib = rng.lognormal(0.0, 0.25, size=(n, 1))
    counts = rng.poisson(rates[None, :] * lib)
    return np.log10(counts + 1.0).astype(np.float32)


def targets_numpy(X: np.ndarray, d: int, seed: int = 1, linear: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((X.shape[1], d))
    Xc = X.astype(np.float64)
    if linear:
        return (Xc @ W).astype(np.float32)
    Xc = Xc - Xc.mean(0, keepdims=True)
    return np.tanh(Xc @ W / np.sqrt(X.shape[1])).astype(np.float32)


def centre_indices(n: int, M: int, seed: int = 2) -> np.ndarray:
    """Uniform draw without replacement, sorted (Falkon's uniform centre selector)."""
    return np.sort(np.random.default_rng(seed).choice(n, size=M, replace=False))


def log_counts_torch(n: int, p: int, seed: int, device, rows_per_chunk: int = 65536,
                     out: torch.Tensor | None = None) -> torch.Tensor:
    """
    Same distribution generated directly on `device` (bench scale: 12-24 GB, never shipped as files).
    Not bit-identical to `log_counts_numpy`; parity tests use the numpy generator.
    """
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rates = torch.from_numpy(
        np.random.default_rng(0).integers(1, 25, size=p).astype(np.float32)).to(device)
    X = out if out is not None else torch.empty(n, p, dtype=torch.float32, device=device)
    for s in range(0, n, rows_per_chunk):
        e = min(n, s + rows_per_chunk)
        lib = torch.exp(0.25 * torch.randn(e - s, 1, device=device, generator=g))
        X[s:e] = torch.log10(torch.poisson(rates[None, :] * lib, generator=g) + 1.0)
    return X


def targets_torch(X: torch.Tensor, d: int, seed: int, rows_per_chunk: int = 65536) -> torch.Tensor:
    g = torch.Generator(device=X.device)
    g.manual_seed(seed)
    p = X.shape[1]
    W = torch.randn(p, d, device=X.device, generator=g) / float(np.sqrt(p))
    mu = torch.zeros(p, device=X.device)
    for s in range(0, X.shape[0], rows_per_chunk):
        mu += X[s : s + rows_per_chunk].sum(0)
    mu /= X.shape[0]
    Y = torch.empty(X.shape[0], d, dtype=torch.float32, device=X.device)
    for s in range(0, X.shape[0], rows_per_chunk):
        Y[s : s + rows_per_chunk] = torch.tanh((X[s : s + rows_per_chunk] - mu) @ W)
    return Y
