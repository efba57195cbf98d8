"""
CPU: host-side logic (FALKON driver, API plumbing, row sharding over gloo) with the test-only
stand-in engine of tests/cpu_ops.py.  Numerics of the CUDA kernels are NOT tested here (see -m gpu).
"""
import math
import os
import pickle
import socket
import warnings

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import falkon_oracle as fo
from sobolev_alignment_b200 import KRRApprox, MultiKRRApprox, mat_inv_sqrt
from sobolev_alignment_b200.engine import KernelSpec, pad_cols
from sobolev_alignment_b200.falkon import Falkon, FalkonOptions
from sobolev_alignment_b200.falkon import models as falkon_models
from sobolev_alignment_b200.falkon.kernels import GaussianKernel, LaplacianKernel, MaternKernel
from sobolev_alignment_b200.solver import Comm, FalkonSolver

from conftest import krr_test_data, relerr
from cpu_ops import CpuOps


# ---------------------------------------------------------------------------- constructors / API
def test_create_null_instance():
    KRRApprox()


def test_all_kernels_construct_under_both_methods():
    for kernel in KRRApprox.sklearn_kernel:
        KRRApprox(kernel=kernel, method="sklearn")
    for kernel in KRRApprox.falkon_kernel:
        clf = KRRApprox(kernel=kernel, method="falkon")
        assert callable(clf.kernel_)


def test_unknown_method_and_kernel_errors():
    with pytest.raises(NotImplementedError):
        KRRApprox(method="keops")
    with pytest.raises(KeyError):
        KRRApprox(method="falkon", kernel="polynomial", kernel_params={"sigma": 1})


def test_matern_nu_mapping_and_validation():
    assert MaternKernel(sigma=2.0, nu=0.5).spec().family == "laplacian"
    assert MaternKernel(sigma=2.0, nu=1.5).spec().family == "matern15"
    assert MaternKernel(sigma=2.0, nu=2.5).spec().family == "matern25"
    assert MaternKernel(sigma=2.0, nu=np.inf).spec().family == "gaussian"
    with pytest.raises(ValueError):
        MaternKernel(sigma=2.0, nu=3.5)
    assert GaussianKernel(sigma=2.0).spec().kscale(4.0) == pytest.approx(1 / (2 * 4 * 16))
    assert LaplacianKernel(sigma=2.0).spec().kscale(4.0) == pytest.approx(1 / 8)


def test_kernels_pickle_roundtrip():
    for k in (GaussianKernel(sigma=3.0), LaplacianKernel(sigma=1.5), MaternKernel(sigma=2.0, nu=2.5)):
        k2 = pickle.loads(pickle.dumps(k))
        assert repr(k2) == repr(k)


def test_falkon_options_accepts_model_selection_keys():
    """krr_model_selection.py:55-64 passes these through KRRApprox(falkon_options=...)."""
    opts = dict(never_store_kernel=True, min_cuda_iter_size_32=np.iinfo(np.int64).max,
                min_cuda_iter_size_64=np.iinfo(np.int64).max, num_fmm_streams=6, max_cpu_mem=1e9, max_gpu_mem=1e9)
    clf = KRRApprox(method="falkon", kernel="matern", kernel_params={"sigma": 1.0, "nu": np.inf},
                    falkon_options=opts)
    clf._setup_clf()
    assert clf.ridge_clf_.options.num_fmm_streams == 6
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        FalkonOptions(no_such_option=1)
        assert any("unknown option" in str(x.message) for x in w)


def test_pad_cols():
    assert [pad_cols(d) for d in (1, 4, 5, 10, 20, 32)] == [4, 4, 8, 12, 20, 32]


def test_mat_inv_sqrt_matches_reference_definition():
    rng = np.random.default_rng(1)
    B = rng.standard_normal((7, 7))
    M = B @ B.T
    S = mat_inv_sqrt(M)
    assert np.allclose(S, fo.mat_inv_sqrt(M), atol=1e-10)
    assert np.allclose(S @ M @ S, np.eye(7), atol=1e-8)
    # rank-deficient: small singular values are dropped, not inverted
    M2 = B[:, :3] @ B[:, :3].T
    assert np.linalg.matrix_rank(mat_inv_sqrt(M2)) == 3
    assert np.allclose(mat_inv_sqrt(torch.tensor(M)), S)


# ---------------------------------------------------------------------------- solver host logic
def _fit_cpu(X, Y, C, family, sigma, lam, maxiter=20, comm=None, **kw):
    solver = FalkonSolver(CpuOps(), KernelSpec(family, sigma), lam, maxiter, comm=comm, **kw)
    return solver.fit(X, Y, C, n_total=None), solver


@pytest.mark.parametrize("family,okernel,nu", [("laplacian", "laplacian", None), ("matern15", "matern", 1.5),
                                               ("gaussian", "rbf", None)])
def test_solver_matches_oracle_single_process(family, okernel, nu):
    X, Xv, W, Y, Yv = krr_test_data(seed=4, n=700, p=30, d=5)
    C = X[:130]          # M not a multiple of 128: exercises the identity padding of the preconditioner
    sigma = float(np.sqrt(60))
    alpha, solver = _fit_cpu(X, Y, C, family, sigma, 1e-4)
    ref = fo.falkon_fit(X, Y, C, okernel, sigma, 1e-4, nu=nu)
    assert alpha.shape == (130, 5)
    assert relerr(alpha, ref) < 1e-6
    # rhs + two launches per iteration + two per full-gradient recomputation (skipped on the last iteration)
    full_grads = sum(1 for it in range(1, solver.n_iter_ + 1) if it % 10 == 0 and it < solver.maxiter)
    assert solver.n_kmv_ == 1 + 2 * (solver.n_iter_ + full_grads)


def test_solver_splits_more_than_32_columns():
    """d = 40 latent columns: two CG runs (32 + 8) sharing the preconditioner; same answer as the oracle."""
    X, Xv, W, Y, Yv = krr_test_data(seed=6, n=500, p=20, d=40)
    C = X[:100]
    alpha, solver = _fit_cpu(X, Y, C, "gaussian", 6.0, 1e-4)
    ref = fo.falkon_fit(X, Y, C, "rbf", 6.0, 1e-4)
    assert alpha.shape == (100, 40)
    assert relerr(alpha, ref) < 1e-6
    assert solver.n_kmv_ == 2 * (1 + 2 * (solver.n_iter_ + 1))


def test_solver_stops_on_tolerance():
    X, Xv, W, Y, Yv = krr_test_data(seed=5, n=200, p=10, d=2)
    # M = n and a negligible jitter: the preconditioner is exact, CG converges at once
    alpha, solver = _fit_cpu(X, Y, X, "gaussian", 5.0, 1e-2, maxiter=50, pc_eps=1e-13, cg_eps=1e-15)
    assert solver.n_iter_ < 10
    assert solver.residuals[-1] < 1e-7


def test_falkon_estimator_on_cpu_standin(monkeypatch, tmp_path):
    monkeypatch.setattr(falkon_models, "CudaOps", CpuOps)
    X, Xv, W, Y, Yv = krr_test_data(seed=6, n=500, p=20, d=3)
    idx = np.sort(np.random.default_rng(0).choice(500, 100, replace=False))
    clf = KRRApprox(method="falkon", kernel="matern", kernel_params={"sigma": 6.0, "nu": 1.5}, M=100,
                    penalization=1e-4)
    clf._setup_clf()
    # same centres for oracle and solver: Falkon draws them from an unseeded host RNG
    orig = Falkon._select_centres
    monkeypatch.setattr(Falkon, "_select_centres", lambda self, n: idx)
    clf.fit(X, Y)
    monkeypatch.setattr(Falkon, "_select_centres", orig)
    ref = fo.falkon_fit(X, Y, X[idx], "matern", 6.0, 1e-4, nu=1.5)
    assert clf.sample_weights_.dtype == torch.float32 and not clf.sample_weights_.is_cuda
    assert clf.sample_weights_.shape == (100, 3)
    assert torch.equal(clf.anchors(), X[idx])
    assert relerr(clf.sample_weights_, ref) < 1e-4
    pred = clf.transform(Xv)
    assert pred.shape == (50, 3) and pred.dtype == torch.float32
    assert relerr(pred, fo.falkon_predict(Xv, X[idx], ref, "matern", 6.0, 1.5)) < 1e-4
    # save / load: the estimator must predict from ASSIGNED ny_points_/alpha_ (krr_approx.py:402-404)
    folder = str(tmp_path / "krr")
    clf.save(folder)
    for f in ("params.pkl", "sample_anchors.pt", "sample_weights.pt", "sample_anchors.csv", "sample_weights.csv"):
        assert os.path.exists(os.path.join(folder, f))
    clf2 = KRRApprox.load(folder)
    assert clf2.method == "falkon" and clf2.M == 100
    assert relerr(clf2.transform(Xv), pred) < 1e-6
    assert np.allclose(np.loadtxt(os.path.join(folder, "sample_weights.csv")), clf.sample_weights_.numpy(), atol=1e-6)
    # MultiKRRApprox is a pure consumer of the same surface
    multi = MultiKRRApprox()
    multi.add_clf(clf)
    multi.add_clf(clf2)
    multi.process_clfs()
    assert multi.anchors().shape == (200, 20) and multi.sample_weights_.shape == (200, 3)
    assert relerr(multi.transform(Xv), pred) < 1e-6


def test_centre_selection_rules():
    f = Falkon(kernel=GaussianKernel(1.0), penalty=1e-3, M=10, seed=3)
    idx = f._select_centres(100)
    assert len(idx) == 10 and len(set(idx.tolist())) == 10 and (np.diff(idx) > 0).all()
    assert np.array_equal(idx, Falkon(kernel=GaussianKernel(1.0), penalty=1e-3, M=10, seed=3)._select_centres(100))
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        assert np.array_equal(Falkon(kernel=GaussianKernel(1.0), penalty=1e-3, M=500)._select_centres(20), np.arange(20))
    explicit = Falkon(kernel=GaussianKernel(1.0), penalty=1e-3, M=3, center_selection=torch.tensor([5, 1, 9]))
    assert explicit._select_centres(100).tolist() == [5, 1, 9]


def test_sklearn_method_still_works():
    """The reference's sklearn branch (tests/test_krr_approx.py:178-205,235-249) on the same logic."""
    X, Xv, W, Y, Yv = krr_test_data(seed=7, n=400, p=30, d=4)
    clf = KRRApprox(method="sklearn", kernel="matern", kernel_params={"length_scale": 30.0, "nu": 1.5},
                    penalization=1e-4).fit(X, Y)
    pred = clf.transform(Xv)
    assert np.corrcoef(np.asarray(pred).ravel(), Yv.numpy().ravel())[0, 1] > 0.99
    rec = clf.kernel_(Xv, X[clf.ridge_samples_idx_, :]).dot(clf.sample_weights_)
    np.testing.assert_array_almost_equal(rec, pred, decimal=3)


def test_reference_krr_approx_runs_on_the_lookalike(monkeypatch):
    """
    INTEGRATION.md option B: the reference's OWN krr_approx.py, loaded by file path, with `falkon`
    aliased to this package's look-alike (CPU stand-in engine, since there is no GPU in this test tier).
    """
    import importlib.util
    import sys

    ref = "/root/reference/sobolev_alignment/krr_approx.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present on this machine")
    import sobolev_alignment_b200.falkon as b200_falkon

    monkeypatch.setitem(sys.modules, "falkon", b200_falkon)
    monkeypatch.setitem(sys.modules, "falkon.kernels", b200_falkon.kernels)
    monkeypatch.setitem(sys.modules, "falkon.options", b200_falkon.options)
    monkeypatch.setattr(falkon_models, "CudaOps", CpuOps)
    spec = importlib.util.spec_from_file_location("ref_krr_approx_aliased", ref)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.FALKON_IMPORTED
    X, Xv, W, Y, Yv = krr_test_data(seed=9, n=600, p=20, d=3)
    clf = mod.KRRApprox(method="falkon", kernel="matern", kernel_params={"sigma": 6.0, "nu": 1.5}, M=150,
                        penalization=1e-4, falkon_options={"never_store_kernel": True, "num_fmm_streams": 6,
                                                           "center_seed": 4})
    clf.fit(X, Y)
    pred = clf.transform(Xv)
    idx = np.sort(np.random.default_rng(4).choice(600, size=150, replace=False))
    ref_alpha = fo.falkon_fit(X, Y, X[idx], "matern", 6.0, 1e-4, nu=1.5)
    assert torch.equal(clf.anchors(), X[idx])
    assert relerr(pred, fo.falkon_predict(Xv, X[idx], ref_alpha, "matern", 6.0, 1.5)) < 1e-4
    assert np.corrcoef(pred.numpy().ravel(), Yv.numpy().ravel())[0, 1] > 0.99


# ---------------------------------------------------------------------------- row sharding over gloo
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        X, Xv, W, Y, Yv = krr_test_data(seed=8, n=601, p=24, d=3)     # 601 rows: ragged shards
        bounds = np.linspace(0, X.shape[0], world + 1).astype(int)
        Xl, Yl = X[bounds[rank]:bounds[rank + 1]], Y[bounds[rank]:bounds[rank + 1]]
        falkon_models.CudaOps = CpuOps     # test-only engine swap inside this worker process
        clf = Falkon(kernel=LaplacianKernel(sigma=7.0), penalty=1e-4, M=90, seed=5,
                     options=FalkonOptions(row_sharded=True))
        clf.fit(Xl, Yl)
        pred = clf.predict(Xv)
        torch.save({"alpha": clf.alpha_, "centres": clf.ny_points_, "pred": pred,
                    "n_total": clf.solver_stats_["n_total"]}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_row_sharded_fit_world2_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["n_total"] == 601
    # replicated state stays bit-identical across ranks
    assert torch.equal(r0["alpha"], r1["alpha"]) and torch.equal(r0["centres"], r1["centres"])
    X, Xv, W, Y, Yv = krr_test_data(seed=8, n=601, p=24, d=3)
    idx = np.sort(np.random.default_rng(5).choice(601, size=90, replace=False))
    assert relerr(r0["centres"], X[idx]) < 1e-7
    ref = fo.falkon_fit(X, Y, X[idx], "laplacian", 7.0, 1e-4)
    assert relerr(r0["alpha"], ref) < 1e-4
    assert relerr(r0["pred"], fo.falkon_predict(Xv, X[idx], ref, "laplacian", 7.0)) < 1e-4


# ---------------------------------------------------------------------------- distributed Cholesky over gloo
def _potrf_rank_main(rank, world, port, out_dir):
    import torch.distributed as dist

    from sobolev_alignment_b200 import distributed

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        M = 7 * 128                                   # 7 outer panels (one block each on the stand-in engine)
        g = torch.Generator().manual_seed(3)
        Z = torch.randn(M, 12, generator=g, dtype=torch.float64)
        K = torch.exp(-torch.cdist(Z, Z) / 4.0) + 1e-2 * torch.eye(M, dtype=torch.float64)
        ops = CpuOps()
        A = K.clone()
        inv = ops.empty(M // 128, 128, 128)
        info = torch.zeros(1, dtype=torch.int32)
        distributed.potrf(ops, A, inv, info, Comm(None, enabled=True))
        torch.save({"L": torch.tril(A), "K": K, "info": info}, os.path.join(out_dir, f"potrf{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_distributed_potrf_gloo(tmp_path, world):
    """Outer panels dealt out over the ranks, panels broadcast with one step of look-ahead: every rank ends with the full factor."""
    port = _free_port()
    mp.spawn(_potrf_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"potrf{r}.pt")) for r in range(world)]
    ref = torch.linalg.cholesky(outs[0]["K"])
    for o in outs:
        assert int(o["info"]) == 0
        assert torch.equal(o["L"], outs[0]["L"])            # bit-identical on all ranks
        assert relerr(o["L"], ref) < 1e-12


def _sharded_large_m_rank_main(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        X, Xv, W, Y, Yv = krr_test_data(seed=8, n=1501, p=24, d=3)
        bounds = np.linspace(0, X.shape[0], world + 1).astype(int)
        Xl, Yl = X[bounds[rank]:bounds[rank + 1]], Y[bounds[rank]:bounds[rank + 1]]
        falkon_models.CudaOps = CpuOps
        clf = Falkon(kernel=LaplacianKernel(sigma=7.0), penalty=1e-4, M=600, seed=5,
                     options=FalkonOptions(row_sharded=True))      # 5 outer panels per factorisation
        clf.fit(Xl, Yl)
        torch.save({"alpha": clf.alpha_, "pred": clf.predict(Xv)}, os.path.join(out_dir, f"big{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_row_sharded_fit_with_distributed_preconditioner_gloo(tmp_path):
    """The row-sharded fit with M large enough that both Cholesky factorisations run panel-distributed."""
    world = 2
    port = _free_port()
    mp.spawn(_sharded_large_m_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "big0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "big1.pt"))
    assert torch.equal(r0["alpha"], r1["alpha"])
    X, Xv, W, Y, Yv = krr_test_data(seed=8, n=1501, p=24, d=3)
    idx = np.sort(np.random.default_rng(5).choice(1501, size=600, replace=False))
    ref = fo.falkon_fit(X, Y, X[idx], "laplacian", 7.0, 1e-4)
    assert relerr(r0["alpha"], ref) < 1e-4
    assert relerr(r0["pred"], fo.falkon_predict(Xv, X[idx], ref, "laplacian", 7.0)) < 1e-4


# ---------------------------------------------------------------------------- consensus features (host part)
def test_interpolation_functions_match_reference_vectors():
    """
    sobolev_alignment_b200.consensus.compute_optimal_tau / project_on_interpolate_PV (batched torch KS statistics)
    against the fixture generated by the reference's interpolated_features.py (oracle/make_golden.py).
    """
    from conftest import load_golden
    from oracle import consensus_oracle as co
    from sobolev_alignment_b200 import consensus

    g = load_golden("consensus")
    proj = {m: {k: g[f"proj_{m}_{k}"] for k in ("source", "target")} for m in ("source", "target")}
    angles = np.arccos(g["cosines"])
    for pv in range(len(angles)):
        ks = consensus.ks_statistics(pv, proj, angles, device=torch.device("cpu"))
        _, ks_ref = co.ks_curve(pv, proj, angles)
        assert np.abs(ks - ks_ref).max() < 1e-12
        tau = consensus.compute_optimal_tau(pv, proj, angles)
        assert abs(tau - g["taus"][pv]) < 1e-12
        s, t = consensus.project_on_interpolate_PV(angles[pv], pv, tau, proj)
        assert np.abs(s - g[f"interp_source_{pv}"]).max() < 1e-12
        assert np.abs(t - g[f"interp_target_{pv}"]).max() < 1e-12


class _FakeKrr:
    """The three things the alignment-side helpers read from a fitted KRRApprox."""

    def __init__(self, kernel, anchors, weights):
        self.kernel_, self._anchors, self.sample_weights_ = kernel, anchors, weights

    def anchors(self):
        return self._anchors


def test_null_model_host_logic_on_stand_in_engine():
    """Batched Gram-Schmidt through K c products == the literal dense restatement (same seed); CPU stand-in engine."""
    from oracle import null_oracle
    from sobolev_alignment_b200 import synthetic as syn
    from sobolev_alignment_b200.null_model import null_model_similarity

    A = torch.from_numpy(syn.log_counts_numpy(70, 9, seed=1))
    B = torch.from_numpy(syn.log_counts_numpy(55, 9, seed=2))
    k = MaternKernel(sigma=4.0, nu=1.5)
    src, tgt = _FakeKrr(k, A, None), _FakeKrr(k, B, None)
    M_X, M_Y = np.diag([2.0, 1.0, 0.5]), np.diag([1.5, 1.0, 0.7])
    for nf in (3, None):
        torch.manual_seed(5)
        got = null_model_similarity(src, tgt, M_X, M_Y, n_iter=40, return_all=True, n_factors=nf, ops=CpuOps())
        torch.manual_seed(5)
        ref = null_oracle.null_model(A, B, M_X, M_Y, "matern", 4.0, 1.5, n_iter=40, n_factors=nf)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-9


def test_consensus_projection_host_logic_on_stand_in_engine():
    """pv_projections (shared frame, folded rotation, chunked rows) == the float64 restatement; CPU stand-in engine."""
    from oracle import consensus_oracle as co
    from sobolev_alignment_b200 import consensus, synthetic as syn

    g = torch.Generator().manual_seed(3)
    Cs, Ct = torch.from_numpy(syn.log_counts_numpy(40, 12, seed=1)), torch.from_numpy(syn.log_counts_numpy(50, 12, seed=2))
    k = LaplacianKernel(sigma=5.0)
    krr = {"source": _FakeKrr(k, Cs, torch.randn(40, 3, generator=g)),
           "target": _FakeKrr(k, Ct, torch.randn(50, 3, generator=g))}
    coef = {"source": torch.randn(3, 40, generator=g).numpy(), "target": torch.randn(3, 50, generator=g).numpy()}
    X = {"source": torch.from_numpy(syn.log_counts_numpy(130, 12, seed=4)),
         "target": torch.from_numpy(syn.log_counts_numpy(90, 12, seed=5))}
    proj = consensus.pv_projections(krr, X, coef, row_chunk=50, ops=CpuOps())
    ref = co.pv_projections(None, {m: krr[m].anchors() for m in krr}, X, coef, "laplacian", 5.0)
    for m in proj:
        for d in proj[m]:
            assert relerr(proj[m][d], ref[m][d]) < 1e-5      # coefficients travel as float32


def test_grid_shares_preconditioner_and_cg_launches_stand_in_engine():
    """
    KRRGrid: per kernel one T, per penalty one A, CG of several penalties in lock-step through shared K_nM
    products -- every grid point equals the independent float64 oracle fit with the same centres.
    """
    from sobolev_alignment_b200.model_selection import KRRGrid

    X, Xv, W, Y, Yv = krr_test_data(seed=12, n=420, p=10, d=3, n_valid=20)
    ops = CpuOps()
    calls = {"kmv": 0}
    orig = ops.kmv

    def counting_kmv(*a, **k):
        calls["kmv"] += 1
        return orig(*a, **k)

    ops.kmv = counting_kmv
    grid = KRRGrid(X, Y, M=64, sigma=4.0, kernel="matern", center_seed=3, ops=ops)
    lams = [1e-6, 1e-4, 1e-2, 1.0, 10.0]
    out = grid.fit(nus=(0.5, 2.5), penalizations=lams)
    assert len(out) == 10
    # 2 kernels x (1 rhs product + 21 product pairs of ONE group of 5 penalties): not 10 x 43
    assert calls["kmv"] == 2 * (1 + 2 * 21)
    C = X[grid.centre_idx]
    for (nu, lam), clf in out.items():
        ref = fo.falkon_fit(X, Y, C, "matern", 4.0, lam, nu=nu)
        assert relerr(clf.sample_weights_, ref) < 1e-6, (nu, lam)
        assert clf.penalization == lam and clf.kernel_params["nu"] == nu
        assert torch.equal(clf.anchors(), C)


def test_stored_kernel_blocks_solver_path_stand_in_engine():
    """kernel_blocks='store' (K formed once per product pair, row block by row block) == the recompute path."""
    X, Xv, W, Y, Yv = krr_test_data(seed=15, n=700, p=12, d=3, n_valid=10)
    idx = np.sort(np.random.default_rng(1).choice(700, size=80, replace=False))
    outs = {}
    for mode in ("recompute", "store"):
        ops = CpuOps()
        ops.kstore_bytes = lambda a, b: ((a + 127) // 128) * 1000       # tiny budget below: 3 row blocks of 256 rows
        ops.wave_rows = 256
        solver = FalkonSolver(ops, KernelSpec("laplacian", 6.0), 1e-4, kernel_blocks=mode, kstore_budget=2000)
        outs[mode] = solver.fit(X, Y, X[idx])
        if mode == "store":
            assert len(solver._blocks) > 1
            assert solver.n_kmv_ == 1 + 21        # rhs + one K formation per product pair (43 launches when recomputed)
    assert relerr(outs["store"], outs["recompute"]) < 1e-10
