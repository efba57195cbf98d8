"""
Drop-in for the subset of the third-party `falkon` package that
/root/reference/sobolev_alignment/krr_approx.py:29-34 imports:

    from falkon import Falkon
    from falkon.kernels import GaussianKernel, LaplacianKernel, MaternKernel
    from falkon.options import FalkonOptions

backed by the B200-native CUDA library (no CPU path).
"""

from . import kernels, options  # noqa: F401
from .models import Falkon  # noqa: F401
from .options import FalkonOptions  # noqa: F401

__all__ = ["Falkon", "FalkonOptions", "kernels", "options"]
