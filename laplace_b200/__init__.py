"""laplace_b200 -- B200-native curvature backend for the Laplace library (``laplace-torch``).

Public surface (mirrors ``laplace.curvature``): ``B200GGN``, ``B200EF`` plus the Kronecker containers
they return.  See DESIGN.md / INTEGRATION.md.
"""
from .backend import B200EF, B200GGN
from .matrix import B200Kron, B200KronDecomposed, adopt

__all__ = ["B200GGN", "B200EF", "B200Kron", "B200KronDecomposed", "adopt"]
__version__ = "0.1.0"
