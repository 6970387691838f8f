"""Times lhb200_shuffle_list (host buffers in/out) against the CPU oracle at mainnet-like sizes."""
import hashlib, json, time
import numpy as np
import lighthouse_b200
from lighthouse_b200.shuffle import shuffle_list
from lighthouse_b200._ffi import lib
import ctypes as C
from tests import oracle_lib as O

lighthouse_b200.init(0)
seed = hashlib.sha256(b"bench").digest()
for n in (500_000, 1_000_000, 1 << 24):
    a = np.arange(n, dtype=np.uint64)
    out = np.empty_like(a)
    sp = (C.c_uint8 * 32).from_buffer_copy(seed)
    for _ in range(3):
        lib.lhb200_shuffle_list(C.c_void_p(a.ctypes.data), n, 90, C.cast(sp, C.c_void_p), 0, C.c_void_p(out.ctypes.data))
    t = time.perf_counter()
    for _ in range(10):
        lib.lhb200_shuffle_list(C.c_void_p(a.ctypes.data), n, 90, C.cast(sp, C.c_void_p), 0, C.c_void_p(out.ctypes.data))
    gpu_ms = (time.perf_counter() - t) * 100
    t = time.perf_counter()
    ref = O.shuffle_list(a, 90, seed, False)
    cpu_ms = (time.perf_counter() - t) * 1000
    print(json.dumps({"n": n, "rounds": 90, "gpu_e2e_ms": round(gpu_ms, 3), "cpu_oracle_ms": round(cpu_ms, 2),
                      "match": bool((np.asarray(ref, dtype=np.uint64) == out).all())}))
