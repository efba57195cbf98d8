"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "sobolev_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sa_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("sa_prep", "sa_prep_log", "sa_colstats", "sa_transform_rows", "sa_kmv", "sa_kmv_workspace_bytes", "sa_kstore_bytes", "sa_kmv_store",
              "sa_ktv_workspace_bytes", "sa_ktv", "sa_kdense", "sa_linalg_workspace_bytes", "sa_potrf", "sa_potrf_outer_blocks", "sa_potrf_begin", "sa_potrf_panel", "sa_potrf_split",
              "sa_potrf_update", "sa_potrf_update_cyclic", "sa_ltl", "sa_ltl_part", "sa_trsm_workspace_bytes", "sa_trsm_packed_bytes", "sa_trsm_pack_workspace_bytes", "sa_trsm_pack",
              "sa_trsm_solve", "sa_trsm_solve_pair", "sa_set_option", "sa_cg_workspace_bytes", "sa_coldot", "sa_cg_update",
              "sa_cg_direction", "sa_axpby", "sa_add_diag", "sa_last_error", "sa_version", "sa_device_info"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    from sobolev_alignment_b200 import _native as nat

    if not os.path.exists(nat.library_path()):
        import __graft_entry__ as ge

        ge.build()
    lib = ctypes.CDLL(nat.library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(nat.EXPORTED_SYMBOLS) == declared_symbols()
    nat.load()
    assert nat.load().sa_version() >= 100
    assert nat.load().sa_last_error() is not None


def test_header_cites_the_reference_interfaces():
    text = open(HEADER).read()
    assert "krr_approx.py:227" in text and ":314" in text and "sobolev_alignment.py:950" in text


def test_product_path_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sobolev_alignment_b200 import KRRApprox
    from sobolev_alignment_b200._native import NativeLibraryError

    clf = KRRApprox(method="falkon", kernel="laplacian", kernel_params={"sigma": 3.0}, M=8)
    with pytest.raises(NativeLibraryError):
        clf.fit(torch.zeros(16, 4), torch.zeros(16, 2))
    with pytest.raises(NativeLibraryError):
        clf.kernel_(torch.zeros(4, 4), torch.zeros(4, 4))
    from sobolev_alignment_b200.preprocessing import InputPreprocessor

    with pytest.raises(NativeLibraryError):
        InputPreprocessor(log_input=True)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sobolev_alignment_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "cpu_ops" not in src, f
