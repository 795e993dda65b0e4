"""ribasim_b200 — B200-native implicit-integration inner loop for Ribasim networks.

Only the hot path of the reference (RHS `water_balance!`, analytic Jacobian, batched
sparse Newton solve, step limiter) lives on the GPU, behind the C ABI declared in
`include/ribasim_b200.h` (`librbs.so`).  The Python side mirrors the host role the
Julia core plays in the reference: it reads TOML + GeoPackage tables, compiles them to
the flat network descriptor and drives the time loop.
"""

__version__ = "0.1.0"

from .tables import ModelTables  # noqa: F401
