"""quisquis-rust_b200 -- B200 (sm_100a) engine for quisquis-rust's batched Ristretto255 hot path.

The directory name carries a hyphen (it is the reference's crate name), so import it through
`__graft_entry__.load_package()` which registers it as module `quisquis_rust_b200`.
The product is the C-ABI shared library `libqq_b200.so` (include/qq_b200.h); this Python layer is the ctypes
binding plus a mirror of the reference's operator interface used by tests/ and bench.py.
"""
from .binding import Engine, MultiEngine, QQError, lib_path, load_library  # noqa: F401
from .api import Account, ElGamalCommitment, RistrettoPublicKey, Verifier  # noqa: F401

__all__ = ["Engine", "MultiEngine", "QQError", "lib_path", "load_library", "Account", "ElGamalCommitment", "RistrettoPublicKey",
           "Verifier"]
