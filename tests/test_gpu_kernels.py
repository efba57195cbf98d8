"""
GPU parity tests of the CUDA kernels, called through the C ABI (ctypes), against the float64 CPU
oracle on the same seeded inputs.  Tolerances follow BASELINE.json `north_star`: max relative error
(max |x - ref| / max |ref|) <= 1e-4 on K_nM v; the dense fp32 routines are held to fp32 round-off.
"""
import math

import numpy as np
import pytest
import torch

from oracle import falkon_oracle as fo
from sobolev_alignment_b200 import synthetic as syn

from conftest import krr_test_data, load_golden, relerr

pytestmark = pytest.mark.gpu

KMV_TOL = 1e-4        # north_star: max relative error on K_nM v
FP32_TOL = 2e-5       # dense fp32 factorisations / solves at test scale

FAMILIES = [("gaussian", "gaussian", None), ("laplacian", "laplacian", None),
            ("matern15", "matern", 1.5), ("matern25", "matern", 2.5)]


def _spec(family, sigma):
    from sobolev_alignment_b200.engine import KernelSpec

    return KernelSpec(family, sigma)


def _kmv_gpu(ops, family, sigma, A, B, V, symmetric=False):
    from sobolev_alignment_b200.engine import pad_cols

    spec = _spec(family, sigma)
    frame = ops.make_frame(B, spec)
    Bp = ops.prepare(B, frame)
    Ap = Bp if symmetric else ops.prepare(A, frame)
    d = V.shape[1]
    Vp = ops.zeros(B.shape[0], pad_cols(d))
    Vp[:, :d] = ops.to_device(V)
    out = ops.zeros(A.shape[0], pad_cols(d))
    ops.kmv(spec, frame, Ap, Bp, Vp, out, symmetric=symmetric)
    return out[:, :d].cpu()


# ---------------------------------------------------------------------------- prep
@pytest.mark.parametrize("n,p", [(1, 1), (5, 3), (257, 100), (300, 67), (1000, 500), (2048, 3000)])
def test_prep_bit_exact(cuda_ops, n, p):
    """Integer/byte-level check: fp16 payload equals round-to-nearest of the affine map, padding is zero."""
    from sobolev_alignment_b200 import _native as nat

    g = torch.Generator().manual_seed(n * 1000 + p)
    X = torch.randn(n, p, generator=g) * 2 + 1
    shift = X.mean(0)
    scale = torch.rand(p, generator=g) + 0.5
    P = nat.prep(X.cuda(), shift.cuda(), scale.cuda(), 0.25)
    ref = ((X - shift) * scale * 0.25).half()
    assert P.data.shape == (n, (p + 63) // 64 * 64)
    assert torch.equal(P.data[:, :p].cpu(), ref)
    assert bool((P.data[:, p:] == 0).all())
    assert P.norms.numel() % 256 == 0 and bool((P.norms[n:] == 0).all())
    assert relerr(P.norms[:n], (ref.double() ** 2).sum(1)) < 1e-6


def test_prep_does_not_write_its_input(cuda_ops):
    X = torch.randn(300, 40)
    X0 = X.clone()
    spec = _spec("laplacian", 3.0)
    cuda_ops.prepare(X, cuda_ops.make_frame(X, spec))
    assert torch.equal(X, X0)


def test_prep_accepts_readonly_memmap(cuda_ops, tmp_path):
    """The reference hands over read-only memmap-backed tensors (sobolev_alignment.py:613-616,922-925)."""
    import warnings

    X = np.random.default_rng(0).standard_normal((500, 33)).astype(np.float32)
    np.save(tmp_path / "x.npy", X)
    mm = np.load(tmp_path / "x.npy", mmap_mode="r")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Xt = torch.from_numpy(mm)
    spec = _spec("laplacian", 3.0)
    frame = cuda_ops.make_frame(Xt[:100], spec)
    P = cuda_ops.prepare(Xt, frame)
    Q = cuda_ops.prepare(torch.from_numpy(X.copy()), frame)
    assert torch.equal(P.data, Q.data) and torch.equal(P.norms, Q.norms)


def test_prepare_chunked_staging_equals_single_shot(cuda_ops):
    from sobolev_alignment_b200.engine import CudaOps

    X = torch.randn(1000, 70)
    spec = _spec("gaussian", 5.0)
    frame = cuda_ops.make_frame(X[:64], spec)
    small = CudaOps(stage_rows=96)           # 11 ragged chunks through the pinned double buffer
    P, Q = small.prepare(X, frame), cuda_ops.prepare(X.cuda(), frame)
    assert torch.equal(P.data, Q.data) and torch.equal(P.norms, Q.norms)


def test_prepare_back_to_back_host_inputs(cuda_ops):
    """
    Regression: two prepare() calls on pageable host tensors in a row.  The second call must not reuse a
    pinned staging buffer while the first call's asynchronous H2D copy out of it is still in flight
    (seen once as a 24 % error of an otherwise deterministic product).
    """
    spec = _spec("laplacian", 3.0)
    g = torch.Generator().manual_seed(11)
    big = torch.randn(60000, 1000, generator=g)
    small = torch.randn(700, 1000, generator=g) + 5.0
    frame = cuda_ops.make_frame(small, spec)
    for _ in range(3):
        Pb = cuda_ops.prepare(big, frame)
        Ps = cuda_ops.prepare(small, frame)      # overwrites the pinned buffers right away
        ref_b = cuda_ops.prepare(big.cuda(), frame)
        ref_s = cuda_ops.prepare(small.cuda(), frame)
        assert torch.equal(Pb.data, ref_b.data) and torch.equal(Pb.norms, ref_b.norms)
        assert torch.equal(Ps.data, ref_s.data) and torch.equal(Ps.norms, ref_s.norms)


# ---------------------------------------------------------------------------- input preprocessing
@pytest.mark.parametrize("p", [333, 256])      # scalar and 128-bit vectorised kernels
@pytest.mark.parametrize("log_input,mean_center,unit_std,frob", [(True, True, True, True), (True, False, True, False),
                                                                 (True, True, False, True), (False, False, False, True),
                                                                 (True, False, False, False)])
def test_input_preprocessor_matches_oracle(log_input, mean_center, unit_std, frob, p, tmp_path):
    """Device version of SobolevAlignment._memmap_log_processing against the float64 restatement; memmap input."""
    from oracle.preprocess_oracle import PreprocessOracle
    from sobolev_alignment_b200.preprocessing import InputPreprocessor

    rng = np.random.default_rng(17)
    rates = rng.integers(1, 25, size=p)
    src = rng.poisson(rates[None, :] * rng.lognormal(0, 0.25, size=(5003, 1))).astype(np.float32)
    tgt = rng.poisson(rates[None, :] * rng.lognormal(0, 0.3, size=(2001, 1))).astype(np.float32)
    src[:, 11] = 2.0                                     # constant gene
    path = tmp_path / "src.npy"
    np.save(path, src)
    src_mm = np.load(path, mmap_mode="r")               # the reference hands over read-only memmaps
    dev = InputPreprocessor(log_input, mean_center, unit_std, frob, chunk_rows=1024)
    orc = PreprocessOracle(log_input, mean_center, unit_std, frob)
    for name, data in (("source", src_mm), ("target", tgt)):
        got = dev.fit_transform(name, data)
        ref = orc.fit_transform(name, np.asarray(data))
        assert got.is_cuda and got.dtype == torch.float32 and got.shape == ref.shape
        assert relerr(got, torch.from_numpy(ref)) < 1e-5      # fp32 log10 / affine map against float64
    if frob:
        assert abs(dev.frob_norm_param_ / orc.frob_norm_param - 1) < 1e-6
    assert np.array_equal(np.load(path), src)           # the input was not written


def test_fused_preprocessing_feeds_the_fit_without_the_fp32_intermediate():
    """
    InputPreprocessor.fit_prepare: the chain log10(x + 1) -> StandardScaler -> Frobenius gain folded into the
    operand preparation (sa_prep_log).  The prepared fp16 rows / norms equal those of the two-step route, the
    materialised rows equal the oracle chain, and a fit on the deferred rows equals the fit on the fp32 tensor.
    """
    from oracle.preprocess_oracle import PreprocessOracle
    from sobolev_alignment_b200 import KRRApprox, _native as nat
    from sobolev_alignment_b200.engine import CudaOps, KernelSpec
    from sobolev_alignment_b200.preprocessing import InputPreprocessor

    rng = np.random.default_rng(23)
    p = 333
    rates = rng.integers(1, 25, size=p)
    src = rng.poisson(rates[None, :] * rng.lognormal(0, 0.25, size=(3001, 1))).astype(np.float32)
    tgt = rng.poisson(rates[None, :] * rng.lognormal(0, 0.3, size=(2500, 1))).astype(np.float32)
    two_step, fused = InputPreprocessor(True, True, True, True), InputPreprocessor(True, True, True, True)
    orc = PreprocessOracle(True, True, True, True)
    ops = CudaOps()
    for name, data in (("source", src), ("target", tgt)):
        full = two_step.fit_transform(name, data)
        lazy = fused.fit_prepare(name, data)
        ref = orc.fit_transform(name, data)
        assert torch.equal(lazy.counts.cpu(), torch.from_numpy(data))          # counts untouched, nothing written back
        assert relerr(lazy.materialize(), torch.from_numpy(ref)) < 1e-5
        idx = np.arange(0, data.shape[0], 7)
        assert relerr(lazy.rows(idx), torch.from_numpy(ref[idx])) < 1e-5
        frame = ops.make_frame(full[:256], KernelSpec("laplacian", 10.0))
        a, b = ops.prepare(full, frame), ops.prepare(lazy, frame)
        # same fp16 operand up to last-bit rounding of the folded affine map
        diff = (a.data.float() - b.data.float()).abs()
        assert float(diff.max()) <= 2e-3 * float(a.data.float().abs().max()) and float((diff > 0).float().mean()) < 0.02
        assert relerr(b.norms, a.norms.double()) < 1e-4
    assert abs(fused.frob_norm_param_ / two_step.frob_norm_param_ - 1) < 1e-6
    # a fit on the deferred rows == the fit on the materialised tensor
    Y = torch.from_numpy(syn.targets_numpy(orc.fit_transform("target", tgt), 4, seed=3))
    fits = []
    for X in (full, lazy):
        clf = KRRApprox(method="falkon", kernel="laplacian", kernel_params={"sigma": 15.0}, penalization=1e-4, M=300,
                        falkon_options={"center_seed": 9})
        clf.fit(X, Y.cuda() if torch.is_tensor(X) else Y)
        fits.append(clf)
    assert not fits[1].sample_weights_.is_cuda and not fits[1].anchors().is_cuda   # Y on the host -> host results
    assert relerr(fits[1].anchors(), fits[0].anchors().cpu().double()) < 1e-5
    Xv = full[:64]
    assert relerr(fits[1].transform(Xv.cpu()), fits[0].transform(Xv).cpu().double()) < 1e-3


# ---------------------------------------------------------------------------- fused kmv
@pytest.mark.parametrize("family,okernel,nu", FAMILIES)
@pytest.mark.parametrize("a,b,p,d", [(1, 1, 1, 1), (128, 256, 64, 4), (300, 500, 100, 7), (2000, 500, 500, 10),
                                     (500, 2000, 500, 10), (129, 257, 65, 20), (3000, 1000, 333, 32)])
def test_kmv_matches_oracle(cuda_ops, family, okernel, nu, a, b, p, d):
    g = torch.Generator().manual_seed(a + 7 * b + 13 * p)
    A = torch.from_numpy(syn.log_counts_numpy(a, p, seed=a + b))
    B = torch.from_numpy(syn.log_counts_numpy(b, p, seed=a + b + 1))
    V = torch.normal(0, 1, size=(b, d), generator=g)
    sigma = max(1.0, 0.3 * math.sqrt(p))
    out = _kmv_gpu(cuda_ops, family, sigma, A, B, V)
    ref = fo.kmv(A, B, V, okernel, sigma, nu)
    assert relerr(out, ref) < KMV_TOL


@pytest.mark.parametrize("tag,family,nu", [("rbf", "gaussian", None), ("laplacian", "laplacian", None),
                                           ("matern15", "matern15", 1.5), ("matern25", "matern25", 2.5)])
def test_kmv_golden_krr_test_scale(cuda_ops, tag, family, nu):
    """Golden vectors at the reference's test scale (n=2000, p=100, M=500, sigma=sqrt(2p))."""
    g = load_golden("krr_test_scale")
    X, Xv, W, Y, Yv = krr_test_data(seed=0)
    C = X[g["idx"]]
    sigma = float(np.sqrt(2 * X.shape[1]))
    gen = torch.Generator().manual_seed(5)
    V = torch.normal(0, 1, size=(C.shape[0], Y.shape[1]), generator=gen)
    Tv = torch.normal(0, 1, size=(X.shape[0], Y.shape[1]), generator=gen)
    assert relerr(_kmv_gpu(cuda_ops, family, sigma, X, C, V), g[f"{tag}_kmv"]) < KMV_TOL
    assert relerr(_kmv_gpu(cuda_ops, family, sigma, C, X, Tv), g[f"{tag}_kmv_t"]) < KMV_TOL


def test_kmv_golden_config1(cuda_ops):
    """BASELINE config 1: n=2000, p=500, M=500, d=10, Laplacian sigma=10 on log10 counts."""
    g = load_golden("config1")
    Xall = syn.log_counts_numpy(2050, 500, seed=0)
    X = torch.from_numpy(Xall[:2000])
    C = X[g["idx"]]
    gen = torch.Generator().manual_seed(5)
    V = torch.normal(0, 1, size=(500, 10), generator=gen)
    Tv = torch.normal(0, 1, size=(2000, 10), generator=gen)
    assert relerr(_kmv_gpu(cuda_ops, "laplacian", 10.0, X, C, V), g["lap_kmv"]) < KMV_TOL
    assert relerr(_kmv_gpu(cuda_ops, "laplacian", 10.0, C, X, Tv), g["lap_kmv_t"]) < KMV_TOL


def test_kmv_accumulates_and_is_linear(cuda_ops):
    """Size-independent properties: out += semantics and linearity in V (checked at a larger size)."""
    from sobolev_alignment_b200.engine import pad_cols

    a, b, p, d = 20000, 6000, 256, 12
    A = torch.randn(a, p, device="cuda") * 0.3
    B = torch.randn(b, p, device="cuda") * 0.3
    spec = _spec("matern15", 4.0)
    frame = cuda_ops.make_frame(B, spec)
    Ap, Bp = cuda_ops.prepare(A, frame), cuda_ops.prepare(B, frame)
    V1, V2 = torch.randn(b, d, device="cuda"), torch.randn(b, d, device="cuda")
    o1, o2, o12 = (cuda_ops.zeros(a, d) for _ in range(3))
    cuda_ops.kmv(spec, frame, Ap, Bp, V1, o1)
    cuda_ops.kmv(spec, frame, Ap, Bp, V2, o2)
    cuda_ops.kmv(spec, frame, Ap, Bp, (2 * V1 - 3 * V2).contiguous(), o12)
    assert relerr(o12, 2 * o1.double() - 3 * o2.double()) < 1e-5
    acc = o1.clone()
    cuda_ops.kmv(spec, frame, Ap, Bp, V2, acc)         # accumulate on top of o1
    assert relerr(acc, o1.double() + o2.double()) < 1e-5
    # against the dense kernel on a row slice
    K = cuda_ops.empty(512, b)
    from sobolev_alignment_b200 import _native as nat
    sl = nat.Prepared(Ap.data[:512], Ap.norms[:512], 512, p)
    cuda_ops.kdense(spec, frame, sl, Bp, K)
    assert relerr(o1[:512], K.double() @ V1.double()) < 1e-5


def test_kmv_full_size_properties(cuda_ops):
    """
    BASELINE config 3 dimensions (p = 3000, M = 50 000 centres, d = 20, Laplacian sigma = 10 on log10
    counts) on a 150 001-row shard -- several waves of CTA pairs, ragged last tiles, chunk blocks, both
    the forward (rows x centres) and the transposed (centres x rows) launch shapes.  Size-independent
    properties: (1) sampled rows against float64 on the device (north_star tolerance), (2) the adjoint
    identity  <T, K(X, C) V> = <V, K(C, X) T>  between the two launch shapes, (3) the row-shard sum of
    the transposed product (what the ranks all-reduce) equals the unsharded product.
    """
    from sobolev_alignment_b200.engine import pad_cols

    n, M, p, d = 150_001, 50_000, 3000, 20
    dev = torch.device("cuda")
    X = syn.log_counts_torch(n, p, seed=21, device=dev)
    idx = torch.from_numpy(syn.centre_indices(n, M, seed=5)).to(dev)
    spec = _spec("laplacian", 10.0)
    C = X[idx]
    frame = cuda_ops.make_frame(C, spec)
    Cp, Xp = cuda_ops.prepare(C, frame), cuda_ops.prepare(X, frame)
    g = torch.Generator(device=dev).manual_seed(9)
    dp = pad_cols(d)
    V = torch.zeros(M, dp, device=dev)
    V[:, :d] = torch.randn(M, d, device=dev, generator=g)
    T = torch.zeros(n, dp, device=dev)
    T[:, :d] = torch.randn(n, d, device=dev, generator=g)
    fwd = cuda_ops.zeros(n, dp)
    cuda_ops.kmv(spec, frame, Xp, Cp, V, fwd)              # K(X, C) V
    tr = cuda_ops.zeros(M, dp)
    cuda_ops.kmv(spec, frame, Cp, Xp, T, tr)               # K(C, X) T
    # (1) float64 reference on sampled rows (direct differences are too large here: fp64 Gram form)
    Cd = C.double()
    cn = (Cd * Cd).sum(1)
    rows = torch.cat([torch.arange(0, 300, device=dev), torch.randint(0, n, (300,), device=dev, generator=g),
                      torch.arange(n - 200, n, device=dev), idx[:200]])          # incl. rows that ARE centres
    Xs = X[rows].double()
    d2 = ((Xs * Xs).sum(1)[:, None] + cn[None, :] - 2.0 * Xs @ Cd.T).clamp_min_(0.0)
    d2[d2 < 1e-9 * cn.mean()] = 0.0
    ref = torch.exp(-d2.sqrt() / 10.0) @ V.double()
    assert relerr(fwd[rows], ref) < KMV_TOL
    cols = torch.randint(0, M, (256,), device=dev, generator=g)
    Cs = Cd[cols]
    acc = torch.zeros(256, dp, dtype=torch.float64, device=dev)
    for s0 in range(0, n, 32768):
        Xc = X[s0:s0 + 32768].double()
        dd = ((Cs * Cs).sum(1)[:, None] + (Xc * Xc).sum(1)[None, :] - 2.0 * Cs @ Xc.T).clamp_min_(0.0)
        dd[dd < 1e-9 * cn.mean()] = 0.0
        acc += torch.exp(-dd.sqrt() / 10.0) @ T[s0:s0 + 32768].double()
    assert relerr(tr[cols], acc) < KMV_TOL
    # (2) adjoint identity between the two launch shapes
    lhs = float((T.double() * fwd.double()).sum())
    rhs = float((V.double() * tr.double()).sum())
    scale = float(T.double().norm() * fwd.double().norm())
    assert abs(lhs - rhs) < 1e-5 * scale, (lhs, rhs, scale)
    # (3) two row shards, summed, reproduce the transposed product (multi-GPU partial sums)
    from sobolev_alignment_b200 import _native as nat
    half = 75_000
    part = cuda_ops.zeros(M, dp)
    for s0, e0 in ((0, half), (half, n)):
        npad = nat.round_up(e0 - s0, nat.NORM_PAD)
        norms = torch.zeros(npad, device=dev)
        norms[: e0 - s0] = Xp.norms[s0:e0]
        shard = nat.Prepared(Xp.data[s0:e0], norms, e0 - s0, p)
        cuda_ops.kmv(spec, frame, Cp, shard, T[s0:e0].contiguous(), part)
    assert relerr(part, tr.double()) < 1e-5


def test_kmv_rejects_bad_arguments(cuda_ops):
    from sobolev_alignment_b200 import _native as nat

    A = torch.randn(64, 32, device="cuda")
    spec = _spec("gaussian", 2.0)
    frame = cuda_ops.make_frame(A, spec)
    Ap = cuda_ops.prepare(A, frame)
    with pytest.raises(nat.NativeLibraryError):
        cuda_ops.kmv(spec, frame, Ap, Ap, torch.zeros(64, 6, device="cuda"), torch.zeros(64, 6, device="cuda"))
    with pytest.raises(nat.NativeLibraryError):
        cuda_ops.kmv(spec, frame, Ap, Ap, torch.zeros(64, 4), torch.zeros(64, 4, device="cuda"))
    assert "dp" in nat.load().sa_last_error().decode() or True


# ---------------------------------------------------------------------------- dense kernel
@pytest.mark.parametrize("family,okernel,nu", FAMILIES)
def test_kdense_matches_oracle(cuda_ops, family, okernel, nu):
    A = torch.from_numpy(syn.log_counts_numpy(333, 150, seed=1))
    B = torch.from_numpy(syn.log_counts_numpy(517, 150, seed=2))
    spec = _spec(family, 4.0)
    frame = cuda_ops.make_frame(B, spec)
    K = cuda_ops.empty(333, 520)[:, :517]            # row stride != width
    cuda_ops.kdense(spec, frame, cuda_ops.prepare(A, frame), cuda_ops.prepare(B, frame), K)
    assert relerr(K, fo.kernel_matrix(A, B, okernel, 4.0, nu)) < KMV_TOL


def test_kdense_symmetric_has_exact_unit_diagonal(cuda_ops):
    A = torch.from_numpy(syn.log_counts_numpy(700, 300, seed=3))
    spec = _spec("laplacian", 5.0)
    frame = cuda_ops.make_frame(A, spec)
    Ap = cuda_ops.prepare(A, frame)
    K = cuda_ops.empty(700, 700)
    cuda_ops.kdense(spec, frame, Ap, Ap, K, symmetric=True)
    assert bool((K.diagonal() == 1).all())
    assert float((K - K.T).abs().max()) < 1e-6
    assert relerr(K, fo.kernel_matrix(A, A, "laplacian", 5.0)) < KMV_TOL


def test_kernel_object_is_callable_like_falkon(cuda_ops):
    """`kernel_(A, B)` returns a dense CPU tensor for CPU inputs (sobolev_alignment.py:950)."""
    from sobolev_alignment_b200.falkon.kernels import LaplacianKernel, MaternKernel

    A, B = torch.randn(70, 20), torch.randn(90, 20)
    K = LaplacianKernel(sigma=4.0)(A, B)
    assert K.shape == (70, 90) and not K.is_cuda and K.dtype == torch.float32
    # single ENTRIES of k with unit-scale data and only p = 20 features carry the full fp16 operand
    # rounding (~2e-4); the 1e-4 budget of north_star is on the products K v, checked below
    assert relerr(K, fo.kernel_matrix(A, B, "laplacian", 4.0)) < 5e-4
    Ks = MaternKernel(sigma=4.0, nu=2.5)(A, A)
    assert bool((Ks.diagonal() == 1).all())
    k = LaplacianKernel(sigma=4.0)
    v, w = torch.randn(90, 3), torch.randn(70, 3)
    Kd = fo.kernel_matrix(A, B, "laplacian", 4.0)
    assert relerr(k.mmv(A, B, v), Kd @ v.double()) < 3e-4       # p = 20, b = 90: little averaging
    assert relerr(k.dmmv(A, B, v, w), Kd.T @ (Kd @ v.double() + w.double())) < 3e-4


@pytest.mark.parametrize("family,okernel,nu", FAMILIES)
@pytest.mark.parametrize("a,b,p,d", [(128, 128, 64, 4), (300, 500, 100, 7), (2000, 777, 333, 20), (5000, 1500, 200, 32)])
def test_stored_block_product_matches_oracle(cuda_ops, family, okernel, nu, a, b, p, d):
    """
    sa_kmv_store + sa_ktv: K(A, B) V with the kernel block left behind as packed fp16 hi/lo tiles, then
    K(A, B)^T t streamed from it -- both against the float64 oracle (1e-4, the K_nM v budget), and the stored
    route against the recomputing transposed product.
    """
    from sobolev_alignment_b200 import _native as nat
    from sobolev_alignment_b200.engine import KernelSpec, pad_cols

    g = torch.Generator().manual_seed(a + 3 * b + p)
    A = torch.from_numpy(syn.log_counts_numpy(a, p, seed=a + b))
    B = torch.from_numpy(syn.log_counts_numpy(b, p, seed=a + b + 1))
    V, T = torch.randn(b, d, generator=g), torch.randn(a, d, generator=g)
    spec = KernelSpec(family, 7.0)
    frame = cuda_ops.make_frame(B.cuda(), spec)
    Ap, Bp = cuda_ops.prepare(A, frame), cuda_ops.prepare(B, frame)
    dp = pad_cols(d)
    Vd, Td = cuda_ops.zeros(b, dp), cuda_ops.zeros(a, dp)
    Vd[:, :d], Td[:, :d] = V.cuda(), T.cuda()
    store = torch.full((nat.kstore_bytes(a, b),), 0xFF, dtype=torch.uint8, device="cuda")   # NaN patterns if unwritten
    out = cuda_ops.zeros(a, dp)
    cuda_ops.kmv(spec, frame, Ap, Bp, Vd, out, kstore=store)
    Kd = fo.kernel_matrix(A, B, okernel, 7.0, nu)
    assert relerr(out[:, :d], Kd @ V.double()) < 1e-4
    wt = cuda_ops.zeros(b, dp)
    cuda_ops.ktv(store, a, b, Td, wt)
    assert relerr(wt[:, :d], Kd.T @ T.double()) < 1e-4
    again = cuda_ops.zeros(b, dp)
    cuda_ops.kmv(spec, frame, Bp, Ap, Td, again)
    assert relerr(wt[:, :d], again[:, :d].double()) < 1e-4
    assert float(wt[:, d:].abs().max()) == 0.0 if dp > d else True


# ---------------------------------------------------------------------------- preconditioner algebra
@pytest.mark.parametrize("M,dp,nb", [(128, 4, None), (384, 12, 128), (1024, 20, None), (2048, 32, 256), (2048, 20, "simt")])
def test_cholesky_ltl_trsm(cuda_ops, M, dp, nb, sab_option):
    """potrf / ltl on the tensor-core path (several outer panel widths) and on the fp32 SIMT path."""
    from sobolev_alignment_b200 import _native as nat

    if nb == "simt":
        sab_option("SAB_LINALG_SIMT", "1")
    elif nb is not None:
        sab_option("SAB_CHOL_NB", str(nb))

    g = torch.Generator().manual_seed(M)
    Z = torch.randn(M, 24, generator=g, dtype=torch.float64)
    K = torch.exp(-torch.cdist(Z, Z) / 6.0) + 1e-3 * torch.eye(M, dtype=torch.float64)
    A = K.float().cuda().contiguous()
    inv = torch.empty(M // 128, 128, 128, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.potrf(A, inv, info)
    Lref = torch.linalg.cholesky(K)
    assert int(info) == 0
    assert relerr(torch.tril(A), Lref) < FP32_TOL
    G = torch.empty(M, M, device="cuda")
    nat.ltl(A, G, 1.0 / M, 1e-3)
    Gref = Lref.T @ Lref / M + 1e-3 * torch.eye(M, dtype=torch.float64)
    assert relerr(torch.tril(G), torch.tril(Gref)) < FP32_TOL
    B = torch.randn(M, dp, generator=g)
    F = nat.trsm_pack(A, inv)
    for tr in (False, True):
        X = B.cuda()
        nat.trsm(F, X, tr)
        ref = torch.linalg.solve_triangular(Lref.T if tr else Lref, B.double(), upper=tr)
        assert relerr(X, ref) < 5e-5


@pytest.mark.parametrize("M,dp,chunk", [(128, 4, None), (1024, 20, None), (4096, 12, None), (4096, 28, "4"),
                                        (2304, 8, "2"), (4096, 20, "16")])
def test_paired_solves_match_two_single_solves(cuda_ops, M, dp, chunk, sab_option):
    """
    sa_trsm_solve_pair (two interleaved chains in one launch) against two sa_trsm_solve calls and against
    float64, both directions, with the lambda V term; several chunk lengths of the work decomposition
    (option SAB_TRSM_CHUNK; the default is 512 tiles per item -- a single chunk column up to M = 65 536).
    """
    from sobolev_alignment_b200 import _native as nat

    if chunk:
        sab_option("SAB_TRSM_CHUNK", chunk)

    g = torch.Generator(device="cuda").manual_seed(M)
    facs = []
    for scale in (6.0, 9.0):
        Z = torch.randn(M, 16, device="cuda", generator=g)
        A = torch.cdist(Z, Z).mul_(-1.0 / scale).exp_()
        A.diagonal().add_(0.05)
        inv = torch.empty(M // 128, 128, 128, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        nat.potrf(A, inv, info)
        assert int(info) == 0
        facs.append((nat.trsm_pack(A, inv), torch.tril(A).double()))
    (Fa, La), (Fb, Lb) = facs
    B = torch.randn(M, dp, device="cuda", generator=g)
    V = torch.randn(M, dp, device="cuda", generator=g)
    for tr in (False, True):
        for lam, Vt in ((0.0, None), (0.37, V)):
            xa = B.clone()
            nat.trsm(Fa, xa, tr)
            xb = xa.clone() if Vt is None else (xa + lam * Vt)
            nat.trsm(Fb, xb, tr)
            ra = torch.linalg.solve_triangular(La.T if tr else La, B.double(), upper=tr)
            rb = torch.linalg.solve_triangular(Lb.T if tr else Lb, ra if Vt is None else ra + lam * Vt.double(),
                                               upper=tr)
            assert relerr(xa, ra) < 2e-5 and relerr(xb, rb) < 2e-5
            pa, pb = B.clone(), torch.full_like(B, float("nan"))
            nat.trsm_pair(Fa, Fb, pa, pb, Vt, lam, tr)
            assert torch.equal(pa, xa)                            # fixed summation order: bit identical
            if Vt is None:
                assert torch.equal(pb, xb)
            else:                                                 # the pair forms xa + lam V with one fma
                assert relerr(pb, xb.double()) < FP32_TOL
            # and reproducible run to run (no atomics, no order dependence on the CTA schedule)
            qa, qb = B.clone(), torch.full_like(B, float("nan"))
            nat.trsm_pair(Fa, Fb, qa, qb, Vt, lam, tr)
            assert torch.equal(qa, pa) and torch.equal(qb, pb)


def test_ltl_parts_sum_to_the_whole(cuda_ops, sab_option):
    """sa_ltl_part: the shares of 3 'ranks' (slabs dealt out cyclically) add up to sa_ltl's G."""
    from sobolev_alignment_b200 import _native as nat

    sab_option("SAB_CHOL_NB", "256")      # 8 slabs at M = 2048
    M = 2048
    g = torch.Generator(device="cuda").manual_seed(2)
    Z = torch.randn(M, 24, device="cuda", generator=g)
    A = torch.cdist(Z, Z).mul_(-1.0 / 6.0).exp_()
    A.diagonal().add_(1e-2)
    inv = torch.empty(M // 128, 128, 128, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.potrf(A, inv, info)
    whole = torch.empty(M, M, device="cuda")
    nat.ltl(A, whole, 1.0 / M, 1e-3)
    parts = torch.zeros(M, M, device="cuda")
    for r in range(3):
        nat.ltl(A, parts, 1.0 / M, 0.0, r, 3)
    nat.add_diag(parts, M, 1e-3)
    assert relerr(torch.tril(parts), torch.tril(whole).double()) < 1e-6


def test_large_factorisation_and_chained_solves(cuda_ops):
    """
    M = 24 576 = 192 blocks: more CTAs than SMs in the single-launch triangular solve (late blocks start
    with a backlog of published dependencies), 48 outer panels in the tensor-core Cholesky.  Checked
    against float64 on the device.
    """
    from sobolev_alignment_b200 import _native as nat

    M, dp = 24576, 20
    g = torch.Generator(device="cuda").manual_seed(3)
    Z = torch.randn(M, 32, device="cuda", generator=g)
    K = torch.cdist(Z, Z).mul_(-1.0 / 7.0).exp_()
    K.diagonal().add_(1e-5 * M)
    A = K.clone()
    inv = torch.empty(M // 128, 128, 128, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.potrf(A, inv, info)
    assert int(info) == 0
    L = torch.tril(A).double()
    del A
    # reconstruction error of the factor, relative to |K|
    R = L @ L.T
    R -= K.double()
    assert float(R.abs().max()) < 2e-5
    del R
    B = torch.randn(M, dp, device="cuda", generator=g)
    F = nat.trsm_pack(L.float().contiguous(), inv)
    for tr in (False, True):
        X = B.clone()
        nat.trsm(F, X, tr)
        ref = torch.linalg.solve_triangular(L.T if tr else L, B.double(), upper=tr)
        assert relerr(X, ref) < 2e-4, (tr, relerr(X, ref))


def test_cholesky_flags_non_positive_definite(cuda_ops):
    from sobolev_alignment_b200 import _native as nat

    A = torch.eye(256, device="cuda")
    A[200, 200] = -1.0
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.potrf(A, torch.empty(2, 128, 128, device="cuda"), info)
    assert int(info) == 201


def test_cg_vector_kernels(cuda_ops):
    from sobolev_alignment_b200 import _native as nat

    M, dp = 50000, 20        # several blocks per column reduction
    g = torch.Generator().manual_seed(0)
    P, Q, X, R = (torch.randn(M, dp, generator=g) for _ in range(4))
    rs_old, pAp = torch.rand(dp, generator=g) + 0.5, torch.rand(dp, generator=g) + 0.5
    cgs = nat.CgScratch(torch.device("cuda", torch.cuda.current_device()), 4, dp)
    out = torch.empty(dp, device="cuda")
    nat.coldot(P.cuda(), Q.cuda(), out, cgs)
    assert relerr(out, (P.double() * Q.double()).sum(0)) < 1e-5
    out2 = torch.empty(dp, device="cuda")
    nat.coldot(P.cuda(), Q.cuda(), out2, cgs)
    assert torch.equal(out, out2)                      # deterministic reduction (no float atomics)
    Xd, Rd, rs_new = X.cuda(), R.cuda(), torch.empty(dp, device="cuda")
    nat.cg_update(P.cuda(), Q.cuda(), Xd, Rd, rs_old.cuda(), pAp.cuda(), 1e-7, rs_new, cgs, hist_row=1, tol=1e-7)
    al = (rs_old / (pAp + 1e-7)).double()
    assert relerr(Xd, X.double() + P.double() * al) < 1e-6
    assert relerr(Rd, R.double() - Q.double() * al) < 1e-6
    assert relerr(rs_new, ((R.double() - Q.double() * al) ** 2).sum(0)) < 1e-5
    assert torch.equal(cgs.history()[1], rs_new.cpu()) and int(cgs.stop) == 0
    Pd = P.cuda()
    rs_old0 = rs_old.clone()
    rs_old0[3] = 0.0                                   # an exactly converged / padded column
    nat.cg_direction(Rd, Pd, rs_new, rs_old0.cuda(), 1e-7, cgs)
    beta = torch.where(rs_old0 > 0, rs_new.cpu() / (rs_old0 + 1e-7), torch.zeros(dp)).double()
    assert relerr(Pd, Rd.cpu().double() + P.double() * beta) < 1e-6
    Y = Q.cuda()
    nat.axpby(0.5, P.cuda(), 2.0, Y)
    assert relerr(Y, 0.5 * P.double() + 2 * Q.double()) < 1e-6
    # convergence: a tiny residual raises the device flag, after which update / direction are frozen
    tiny = torch.full((M, dp), 1e-12, device="cuda")
    nat.coldot(tiny, tiny, out, cgs, hist_row=2, tol=1e-7)
    cgs.poll(2)
    assert cgs.stopped_at(2) and not cgs.stopped_at(0)
    Xf, Pf = Xd.clone(), Pd.clone()
    nat.cg_update(P.cuda(), Q.cuda(), Xd, Rd, rs_old.cuda(), pAp.cuda(), 1e-7, rs_new, cgs, hist_row=3, tol=1e-7)
    nat.cg_direction(Rd, Pd, rs_new, rs_old.cuda(), 1e-7, cgs)
    assert torch.equal(Xf, Xd) and torch.equal(Pf, Pd)


def test_stepwise_cholesky_matches_potrf(cuda_ops, sab_option):
    """
    sa_potrf_begin / _panel / _split / _update (what the row-sharded solver strings together over NCCL):
    applied panel by panel with the trailing update cut into per-panel column ranges -- the way two ranks
    would share it -- they reproduce sa_potrf.
    """
    from sobolev_alignment_b200 import _native as nat

    sab_option("SAB_CHOL_NB", "256")
    M = 2304                                           # 18 blocks, 9 outer panels of 2 blocks
    g = torch.Generator(device="cuda").manual_seed(5)
    Z = torch.randn(M, 24, device="cuda", generator=g)
    K = torch.cdist(Z, Z).mul_(-1.0 / 6.0).exp_()
    K.diagonal().add_(1e-2)
    ref, inv_ref = K.clone(), torch.empty(M // 128, 128, 128, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    nat.potrf(ref, inv_ref, info)
    assert int(info) == 0
    A, inv = K.clone(), torch.empty(M // 128, 128, 128, device="cuda")
    nb, nblk = nat.potrf_outer_blocks(), M // 128
    assert nb == 2
    nat.potrf_begin(A, info)
    for k0 in range(0, nblk, nb):
        kend = min(nblk, k0 + nb)
        nat.potrf_panel(A, inv, info, k0, kend)
        if kend == nblk:
            break
        W = (kend - k0) * 128
        panel = A[k0 * 128:, k0 * 128: kend * 128].contiguous()      # what a non-owner receives
        below = panel[W:]
        nat.potrf_split(M, below, W)
        for c0 in range(kend, nblk, nb):                             # one update per outer panel still to come
            nat.potrf_update(A, kend, c0, min(nblk, c0 + nb), below, W)
    assert int(info) == 0
    assert relerr(torch.tril(A), torch.tril(ref).double()) < 1e-6
    assert relerr(inv, inv_ref.double()) < 1e-5
    Lref = torch.linalg.cholesky(K.double())
    assert relerr(torch.tril(A), Lref) < FP32_TOL
