"""eagle-mpc_b200 — B200-native SbFDDP hot path of eagle-mpc (host mirror + CUDA kernels behind a C ABI).

The directory name contains a hyphen (it mirrors the reference's name); import it with
    import importlib; empc = importlib.import_module("eagle-mpc_b200")
"""
from . import abi  # noqa: F401
