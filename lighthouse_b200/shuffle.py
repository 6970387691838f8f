"""Mirror of consensus/swap_or_not_shuffle::shuffle_list (src/shuffle_list.rs:79-160) over the CUDA library."""
import ctypes as C

import numpy as np

from ._ffi import lib, buf, EINVAL

SHUFFLE_ROUND_COUNT = 90  # mainnet (consensus/types/presets/mainnet/phase0.yaml)


def shuffle_list(input_list, rounds: int, seed: bytes, forwards: bool):
    """-> list of ints, or None where the reference returns None (empty list, > 2**24 elements, zero rounds)."""
    a = np.ascontiguousarray(input_list, dtype=np.uint64)
    n = a.shape[0]
    if n == 0 or n > (1 << 24) or rounds == 0 or len(seed) != 32:
        return None
    out = np.empty_like(a)
    sp, keep = buf(seed)
    rc = lib.lhb200_shuffle_list(C.c_void_p(a.ctypes.data), n, rounds, sp, 1 if forwards else 0, C.c_void_p(out.ctypes.data))
    if rc == EINVAL:
        return None
    if rc != 0:
        from ._ffi import Lhb200Error
        raise Lhb200Error(rc, "lhb200_shuffle_list")
    return out.tolist()
