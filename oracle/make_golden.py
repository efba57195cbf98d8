"""
Generates the golden fixtures under tests/golden/ (TEST INFRASTRUCTURE ONLY).

Run in the authoring container (needs /root/reference for the sklearn pin):

    python oracle/make_golden.py

It (1) pins the oracle against the reference's own runnable code -- `KRRApprox(method="sklearn")`
loaded by file path from /root/reference/sobolev_alignment/krr_approx.py -- in the M = n regime where
FALKON and exact KRR coincide, and records both outputs; (2) writes seeded golden vectors for the
Falkon path at the reference's test scale (tests/test_krr_approx.py:8-14) and at BASELINE config 1.
Inputs are regenerated from seeds by `sobolev_alignment_b200.synthetic`; the fixtures hold checksums of
the inputs plus the oracle outputs.
"""

from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import falkon_oracle as fo  # noqa: E402
from sobolev_alignment_b200 import synthetic as syn  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/sobolev_alignment/krr_approx.py"


def load_reference_krr():
    spec = importlib.util.spec_from_file_location("ref_krr_approx", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.KRRApprox


def krr_test_data(seed=0, n=2000, p=100, d=7, n_valid=50):
    """The reference's KRR test distribution (tests/test_krr_approx.py:22-44), seeded."""
    g = torch.Generator().manual_seed(seed)
    X = torch.normal(0, 1, size=(n, p), generator=g)
    Xv = torch.normal(0, 1, size=(n_valid, p), generator=g)
    W = torch.normal(0, 1, size=(p, d), generator=g)
    return X, Xv, W, X.matmul(W), Xv.matmul(W)


def sklearn_pin():
    """M = n: FALKON == exact KRR with sklearn alpha = lambda * n."""
    KRRApprox = load_reference_krr()
    n, p, d = 400, 60, 5
    X, Xv, W, Y, Yv = krr_test_data(seed=11, n=n, p=p, d=d, n_valid=40)
    out = {}
    lam = 1e-4
    sigma = float(np.sqrt(2 * p))
    for name, ref_kwargs, (kern, nu) in [
        # Matern nu = 0.5 IS the L2 Laplacian exp(-|a-b|_2 / sigma) -- the kernel of BASELINE configs 1-3 --
        # through the reference's own sklearn path (krr_approx.py:59-64,176-178,251)
        ("laplacian", dict(kernel="matern", kernel_params={"length_scale": sigma, "nu": 0.5}), ("laplacian", None)),
        ("matern15", dict(kernel="matern", kernel_params={"length_scale": sigma, "nu": 1.5}), ("matern", 1.5)),
        ("matern25", dict(kernel="matern", kernel_params={"length_scale": sigma, "nu": 2.5}), ("matern", 2.5)),
        ("rbf", dict(kernel="rbf", kernel_params={"gamma": 1.0 / (2 * sigma * sigma)}), ("gaussian", None)),
    ]:
        ref = KRRApprox(method="sklearn", penalization=lam * n, **ref_kwargs)
        ref.fit(X.double(), Y.double())
        ref_pred = np.asarray(ref.transform(Xv.double()))
        alpha = fo.falkon_fit(X, Y, X, kern, sigma, lam, nu=nu, maxiter=20, pc_eps=1e-13, cg_eps=1e-15)
        pred = fo.falkon_predict(Xv, X, alpha, kern, sigma, nu).numpy()
        err = np.abs(pred - ref_pred).max() / np.abs(ref_pred).max()
        print(f"[sklearn pin] {name}: max rel err oracle vs reference sklearn path = {err:.3e}")
        assert err < 1e-6, err
        out[f"{name}_ref_pred"] = ref_pred
        out[f"{name}_oracle_pred"] = pred
        out[f"{name}_err"] = np.array(err)
    np.savez_compressed(os.path.join(GOLD, "sklearn_pin.npz"), **out)


def falkon_case(tag, X, Y, Xv, idx, kernel, sigma, lam, nu=None, seed_v=5):
    C = X[idx]
    g = torch.Generator().manual_seed(seed_v)
    V = torch.normal(0, 1, size=(C.shape[0], Y.shape[1]), generator=g)
    Tv = torch.normal(0, 1, size=(X.shape[0], Y.shape[1]), generator=g)
    trace = []
    alpha = fo.falkon_fit(X, Y, C, kernel, sigma, lam, nu=nu, maxiter=20, trace=trace)
    out = {
        f"{tag}_kmv": fo.kmv(X, C, V, kernel, sigma, nu).numpy(),
        f"{tag}_kmv_t": fo.kmv(C, X, Tv, kernel, sigma, nu).numpy(),
        f"{tag}_alpha": alpha.numpy(),
        f"{tag}_pred": fo.falkon_predict(Xv, C, alpha, kernel, sigma, nu).numpy(),
        f"{tag}_kvalid": fo.kernel_matrix(Xv, C, kernel, sigma, nu).numpy()[:, :64],
        f"{tag}_trace": np.array(trace),
    }
    print(f"[{tag}] |alpha|max={alpha.abs().max():.3e} cg residual trace: "
          f"{trace[0]:.2e} -> {trace[-1]:.2e} ({len(trace)} it)")
    return out, alpha


def krr_test_goldens():
    X, Xv, W, Y, Yv = krr_test_data(seed=0)
    idx = syn.centre_indices(X.shape[0], 500, seed=2)
    sigma = float(np.sqrt(2 * X.shape[1]))
    out = {"idx": idx, "x_sum": np.array(float(X.double().sum())), "y_sum": np.array(float(Y.double().sum()))}
    for tag, kernel, nu in [("rbf", "rbf", None), ("matern15", "matern", 1.5),
                            ("matern25", "matern", 2.5), ("laplacian", "laplacian", None)]:
        o, alpha = falkon_case(tag, X, Y, Xv, idx, kernel, sigma, 1e-4, nu)
        pear = np.corrcoef(o[f"{tag}_pred"].ravel(), Yv.numpy().ravel())[0, 1]
        print(f"   pearson(pred, truth) = {pear:.5f}")
        out.update(o)
    np.savez_compressed(os.path.join(GOLD, "krr_test_scale.npz"), **out)


def config1_goldens():
    """BASELINE config 1: n=2000, p=500, d=10, Laplacian sigma=10, penalization 1e-3, M=500."""
    n, p, d, M = 2000, 500, 10, 500
    Xall = syn.log_counts_numpy(n + 50, p, seed=0)
    Yall = syn.targets_numpy(Xall, d, seed=1, linear=True)
    X, Xv = torch.from_numpy(Xall[:n]), torch.from_numpy(Xall[n:])
    Y = torch.from_numpy(Yall[:n])
    idx = syn.centre_indices(n, M, seed=2)
    out = {"idx": idx, "x_sum": np.array(float(X.double().sum())), "y_sum": np.array(float(Y.double().sum()))}
    o, alpha = falkon_case("lap", X, Y, Xv, idx, "laplacian", 10.0, 1e-3)
    out.update(o)
    np.savez_compressed(os.path.join(GOLD, "config1.npz"), **out)


def alignment_goldens():
    """Config 5 at test scale: two fits + cross-kernels + principal vectors."""
    n, p, d, M = 1500, 200, 8, 300
    Xs = syn.log_counts_numpy(n, p, seed=21)
    Xt = syn.log_counts_numpy(n, p, seed=22)
    Ys = torch.from_numpy(syn.targets_numpy(Xs, d, seed=23))
    Yt = torch.from_numpy(syn.targets_numpy(Xt, d, seed=23))  # same encoder on both sides
    Xs, Xt = torch.from_numpy(Xs), torch.from_numpy(Xt)
    idx_s = syn.centre_indices(n, M, seed=25)
    idx_t = syn.centre_indices(n, M, seed=26)
    out = {"idx_s": idx_s, "idx_t": idx_t}
    for tag, kernel, sigma, nu in [("lap", "laplacian", 8.0, None), ("m15", "matern", 8.0, 1.5)]:
        a_s = fo.falkon_fit(Xs, Ys, Xs[idx_s], kernel, sigma, 1e-6, nu=nu)
        a_t = fo.falkon_fit(Xt, Yt, Xt[idx_t], kernel, sigma, 1e-6, nu=nu)
        al = fo.alignment(a_s, Xs[idx_s], a_t, Xt[idx_t], kernel, sigma, nu)
        print(f"[align {tag}] principal cosines: {np.round(al['principal_angles'], 4)}")
        out[f"{tag}_alpha_s"] = a_s.numpy()
        out[f"{tag}_alpha_t"] = a_t.numpy()
        for k, v in al.items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, "alignment.npz"), **out)


def consensus_goldens():
    """
    Pins oracle/consensus_oracle.py against the reference's own interpolated_features.py (imported by file
    path): optimal interpolation times and interpolated projections on seeded principal-vector projections.
    """
    from oracle import consensus_oracle as co

    spec = importlib.util.spec_from_file_location("ref_interp", "/root/reference/sobolev_alignment/interpolated_features.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(7)
    ns, nt, d = 700, 900, 4
    shift = rng.normal(0, 1.0, size=d)
    proj = {"source": {"source": rng.normal(0, 1, (ns, d)), "target": rng.normal(0, 1, (nt, d)) + 0.5 * shift},
            "target": {"source": rng.normal(0, 1, (ns, d)) - 0.3 * shift, "target": rng.normal(0, 1.5, (nt, d))}}
    for m in proj:
        for k in proj[m]:
            proj[m][k] = proj[m][k] - proj[m][k].mean(0, keepdims=True)
    cosines = np.array([0.95, 0.8, 0.55, 0.2])
    angles = np.arccos(cosines)
    out = {"cosines": cosines}
    for m in proj:
        for k in proj[m]:
            out[f"proj_{m}_{k}"] = proj[m][k]
    taus = []
    for pv in range(d):
        t_ref = ref.compute_optimal_tau(pv, proj, angles, n_interpolation=100)
        t_or = co.optimal_tau(pv, proj, angles)
        s_ref, g_ref = ref.project_on_interpolate_PV(angles[pv], pv, t_ref, proj)
        s_or, g_or = co.interpolate(angles[pv], pv, t_or, proj)
        print(f"[consensus pin] PV {pv}: tau reference {t_ref:.2f} oracle {t_or:.2f}; "
              f"max |interp diff| {max(np.abs(s_ref - s_or).max(), np.abs(g_ref - g_or).max()):.1e}")
        assert abs(t_ref - t_or) < 1e-12 and np.abs(s_ref - s_or).max() < 1e-12 and np.abs(g_ref - g_or).max() < 1e-12
        taus.append(t_ref)
        out[f"interp_source_{pv}"], out[f"interp_target_{pv}"] = s_ref, g_ref
    out["taus"] = np.array(taus)
    np.savez_compressed(os.path.join(GOLD, "consensus.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    sklearn_pin()
    krr_test_goldens()
    config1_goldens()
    alignment_goldens()
    consensus_goldens()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
