"""
sobolev_alignment_b200 -- B200-native kernel-ridge-regression hot path of Sobolev Alignment.

Public surface mirrors what the reference exports for this path
(/root/reference/sobolev_alignment/__init__.py:1-3 and krr_approx.py / multi_krr_approx.py /
kernel_operations.py): `KRRApprox`, `MultiKRRApprox`, `mat_inv_sqrt`, plus the `falkon` look-alike
sub-package and the encoder comparison of sobolev_alignment.py:929-988.
"""

from .kernel_operations import mat_inv_sqrt  # noqa: F401
from .krr_approx import KRRApprox  # noqa: F401
from .multi_krr_approx import MultiKRRApprox  # noqa: F401

__version__ = "0.1.0"
__all__ = ["KRRApprox", "MultiKRRApprox", "mat_inv_sqrt"]
