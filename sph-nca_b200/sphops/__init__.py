"""Drop-in `sphops` package: same public names as the reference (code/sphops/__init__.py:1-2),
implemented by hand-written sm_100a CUDA kernels behind the C ABI in include/sphnca.h."""
from .preprocess import initialize_hashgrid
from .ops import volume, gradient, divergence, count, blur

__all__ = ["initialize_hashgrid", "volume", "gradient", "divergence", "count", "blur"]


def fused_available():
    """True when libsphnca exports the fused NCA step (sphnca_step_fwd / sphnca_step_bwd)."""
    from . import _lib
    return "sphnca_step_fwd" not in _lib.MISSING_SYMBOLS


from . import fused  # noqa: E402
