"""Module path kept for code that reaches `sphops.hashgrid.single_impl.Hashgrid` like the reference."""
from . import Hashgrid, INTTYPE  # noqa: F401
