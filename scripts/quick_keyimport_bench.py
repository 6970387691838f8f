"""Start-up key import (ValidatorPubkeyCache::import): n compressed 48-byte keys through
lhb200_g1_decompress_validate (decompress + infinity + subgroup check), wall time with host buffers, keys/s.

usage: python scripts/quick_keyimport_bench.py [n ...]   -> one JSON line per n
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lighthouse_b200 as lhb
from lighthouse_b200 import bls


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16384, 500000]
    lhb.init(0)
    rng = np.random.default_rng(7)
    for n in sizes:
        sks = rng.integers(1, 2 ** 62, size=n, dtype=np.int64)
        sk_bytes = b"".join(int(s).to_bytes(32, "big") for s in sks)
        pk48, pk96 = bls.sk_to_pk(sk_bytes)
        unc, st = bls.decompress_validate_pubkeys(pk48)          # warm-up + check
        assert not st.any() and unc == pk96
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            bls.decompress_validate_pubkeys(pk48)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps({"workload": f"{n} compressed pubkeys -> decompress + key_validate (host buffers, wall)",
                          "n": n, "ms": best * 1e3, "keys_per_s": n / best}))


if __name__ == "__main__":
    main()
