"""lighthouse_b200 — B200-native (sm_100a) batch BLS12-381 verification and SSZ SHA-256 merkleization
behind Lighthouse's `bls::SignatureSet` / `tree_hash::TreeHash` surfaces.  See DESIGN.md.

Importing the package loads liblhb200.so (fails loudly if it has not been built); nothing here falls back
to CPU arithmetic.
"""
import os as _os

# more hardware work queues than the default 8: concurrent verify calls each drive their own streams (see lhb200_init);
# must be in the environment before anything in the process creates the CUDA context (e.g. torch)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _ffi  # noqa: E402,F401  (raises ImportError if the CUDA library is missing)
from ._ffi import init, Lhb200Error  # noqa: E402,F401

__all__ = ["init", "Lhb200Error", "tree_hash", "merkle_proof", "bls", "shuffle"]
