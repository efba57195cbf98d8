"""CPU: the oracle against its golden vectors and against the reference's runnable (sklearn) path."""
import math

import numpy as np
import pytest
import torch

from oracle import falkon_oracle as fo
from sobolev_alignment_b200 import synthetic as syn

from conftest import krr_test_data, load_golden, relerr


def test_sklearn_pin_recorded():
    """The generating script asserted <1e-6; the fixture records what it measured (~2e-11)."""
    g = load_golden("sklearn_pin")
    for name in ("laplacian", "matern15", "matern25", "rbf"):
        assert float(g[f"{name}_err"]) < 1e-9
        assert relerr(g[f"{name}_oracle_pred"], g[f"{name}_ref_pred"]) < 1e-9


def test_oracle_reproduces_sklearn_pin():
    """Recompute the M = n FALKON solution and compare with the reference sklearn predictions."""
    g = load_golden("sklearn_pin")
    n, p, d = 400, 60, 5
    X, Xv, W, Y, Yv = krr_test_data(seed=11, n=n, p=p, d=d, n_valid=40)
    sigma = float(np.sqrt(2 * p))
    for name, kern, nu in (("laplacian", "laplacian", None), ("matern15", "matern", 1.5),
                           ("matern25", "matern", 2.5), ("rbf", "gaussian", None)):
        alpha = fo.falkon_fit(X, Y, X, kern, sigma, 1e-4, nu=nu, pc_eps=1e-13, cg_eps=1e-15)
        pred = fo.falkon_predict(Xv, X, alpha, kern, sigma, nu)
        assert relerr(pred, g[f"{name}_ref_pred"]) < 1e-8, name


def test_matern_matches_sklearn_kernel():
    from sklearn.gaussian_process.kernels import Matern
    from sklearn.metrics.pairwise import rbf_kernel

    A = torch.randn(40, 9, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    B = torch.randn(30, 9, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    for nu in (0.5, 1.5, 2.5):
        ref = Matern(length_scale=2.0, nu=nu)(A.numpy(), B.numpy())
        assert relerr(fo.kernel_matrix(A, B, "matern", 2.0, nu), ref) < 1e-12
    assert relerr(fo.kernel_matrix(A, B, "rbf", 2.0), rbf_kernel(A.numpy(), B.numpy(), gamma=1 / 8.0)) < 1e-12
    assert relerr(fo.kernel_matrix(A, B, "matern", 2.0, math.inf), fo.kernel_matrix(A, B, "gaussian", 2.0)) == 0


def test_direct_and_gemm_distances_agree():
    A = torch.randn(50, 20, dtype=torch.float64)
    B = torch.randn(60, 20, dtype=torch.float64)
    assert relerr(fo.sq_dists(A, B, direct=False), fo.sq_dists(A, B, direct=True)) < 1e-12


@pytest.mark.parametrize("tag,kernel,nu", [("laplacian", "laplacian", None), ("matern15", "matern", 1.5)])
def test_golden_krr_test_scale(tag, kernel, nu):
    g = load_golden("krr_test_scale")
    X, Xv, W, Y, Yv = krr_test_data(seed=0)
    assert abs(float(X.double().sum()) - float(g["x_sum"])) < 1e-6
    C = X[g["idx"]]
    sigma = float(np.sqrt(2 * X.shape[1]))
    V = torch.normal(0, 1, size=(C.shape[0], Y.shape[1]), generator=torch.Generator().manual_seed(5))
    assert relerr(fo.kmv(X, C, V, kernel, sigma, nu), g[f"{tag}_kmv"]) < 1e-12
    alpha = fo.falkon_fit(X, Y, C, kernel, sigma, 1e-4, nu=nu)
    assert relerr(alpha, g[f"{tag}_alpha"]) < 1e-7
    pred = fo.falkon_predict(Xv, C, alpha, kernel, sigma, nu)
    assert relerr(pred, g[f"{tag}_pred"]) < 1e-8
    # the reference's quality bar (tests/test_krr_approx.py:207-233)
    assert np.corrcoef(pred.numpy().ravel(), Yv.numpy().ravel())[0, 1] > 0.99


def test_golden_config1_inputs_stable():
    g = load_golden("config1")
    Xall = syn.log_counts_numpy(2050, 500, seed=0)
    assert abs(float(Xall[:2000].astype(np.float64).sum()) - float(g["x_sum"])) < 1e-6
    assert np.array_equal(syn.centre_indices(2000, 500, seed=2), g["idx"])


def test_f32_port_close_to_f64_oracle():
    X, Xv, W, Y, Yv = krr_test_data(seed=3, n=600, p=40, d=3)
    C = X[:150]
    sigma = float(np.sqrt(80))
    a64 = fo.falkon_fit(X, Y, C, "laplacian", sigma, 1e-4)
    a32 = fo.falkon_fit_f32(X, Y, C, "laplacian", sigma, 1e-4)
    p64 = fo.falkon_predict(Xv, C, a64, "laplacian", sigma)
    p32 = fo.falkon_predict(Xv, C, a32.double(), "laplacian", sigma)
    assert relerr(p32, p64) < 5e-3
    # with V = alpha (large entries of both signs) K V cancels heavily: even the float32 GEMM-based
    # reference structure is ~4e-4 off the float64 answer here; a random V still gives ~1e-4 (fp32 cancellation, no fp16 involved)
    assert relerr(fo.kmv_f32(X, C, a32, "laplacian", sigma), fo.kmv(X, C, a32, "laplacian", sigma)) < 2e-3
    V = torch.randn(150, 3, generator=torch.Generator().manual_seed(0))
    assert relerr(fo.kmv_f32(X, C, V, "laplacian", sigma), fo.kmv(X, C, V, "laplacian", sigma)) < 5e-4  # fp32 |x|^2+|y|^2-2xy cancellation


def test_mat_inv_sqrt_identity():
    rng = np.random.default_rng(0)
    B = rng.standard_normal((6, 6))
    M = B @ B.T + 0.1 * np.eye(6)
    S = fo.mat_inv_sqrt(M)
    assert np.allclose(S @ M @ S, np.eye(6), atol=1e-10)


# ---------------------------------------------------------------------------- preprocessing chain
def test_preprocess_oracle_matches_sklearn_and_the_reference_lines():
    """
    The restatement of sobolev_alignment.py:851-902 against the calls the reference itself makes:
    np.log10(x + 1), sklearn StandardScaler(with_mean, with_std).fit_transform, mean row norm scaling.
    """
    from sklearn.preprocessing import StandardScaler

    from oracle.preprocess_oracle import PreprocessOracle

    rng = np.random.default_rng(3)
    src = rng.poisson(rng.integers(1, 25, size=40)[None, :] * rng.lognormal(0, 0.25, size=(300, 1))).astype(np.float64)
    tgt = rng.poisson(rng.integers(1, 25, size=40)[None, :] * rng.lognormal(0, 0.25, size=(200, 1))).astype(np.float64)
    src[:, 7] = 3.0                          # constant genes: StandardScaler leaves their scale at 1
    src[:, 9] = 2.0                          # (log10(3) does not average exactly: sklearn's round-off rule)
    tgt[:, 5] = 0.0
    for mc in (False, True):
        for us in (False, True):
            orc = PreprocessOracle(log_input=True, mean_center=mc, unit_std=us, frob_norm_source=True)
            got_s, got_t = orc.fit_transform("source", src), orc.fit_transform("target", tgt)
            ref_s = StandardScaler(with_mean=mc, with_std=us).fit_transform(np.log10(src + 1))
            ref_t = StandardScaler(with_mean=mc, with_std=us).fit_transform(np.log10(tgt + 1))
            param = np.mean(np.linalg.norm(ref_s, axis=1))
            ref_t = ref_t * param / np.mean(np.linalg.norm(ref_t, axis=1))
            assert np.allclose(got_s, ref_s, rtol=0, atol=1e-12) and np.allclose(got_t, ref_t, rtol=0, atol=1e-12)
            assert abs(orc.frob_norm_param - param) < 1e-12
    raw = PreprocessOracle(log_input=False, frob_norm_source=False).fit_transform("source", src)
    assert np.array_equal(raw, src)


def _golden_projections():
    g = load_golden("consensus")
    proj = {m: {k: g[f"proj_{m}_{k}"] for k in ("source", "target")} for m in ("source", "target")}
    return g, proj


def test_consensus_oracle_matches_reference_interpolation():
    """oracle/consensus_oracle.py against vectors produced by the reference's own interpolated_features.py."""
    from oracle import consensus_oracle as co

    g, proj = _golden_projections()
    angles = np.arccos(g["cosines"])
    for pv in range(len(angles)):
        tau = co.optimal_tau(pv, proj, angles)
        assert abs(tau - g["taus"][pv]) < 1e-12
        s, t = co.interpolate(angles[pv], pv, tau, proj)
        assert np.abs(s - g[f"interp_source_{pv}"]).max() < 1e-12
        assert np.abs(t - g[f"interp_target_{pv}"]).max() < 1e-12
