"""Prints verdict, statuses and GT bytes of a fixed small batch (valid and with two swapped messages).  The kernel-selection
switches are read once per process, so tests/test_bls_gpu.py::test_latency_and_lane_modes_agree runs this script twice —
default (latency-mode kernels) and with LHB_G2_WARP=0 LHB_MILLER_WARP=0 LHB_FINAL_WARP=0 — and compares the output."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import lighthouse_b200
from lighthouse_b200 import bls
from lighthouse_b200.synthetic import attestation_batch

lighthouse_b200.init(0)
n, k = 37, 5
ab = attestation_batch(n, keys_per_set=k, n_validators=512, seed=0xD1FF)
rands = (np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) | np.uint64(1)
b = bls.Batch(n, n * k)
msgs = bytearray(ab.msgs)
msgs[0:32], msgs[32:64] = ab.msgs[32:64], ab.msgs[0:32]
sigs_bad = bytearray(ab.sigs)
sigs_bad[96 * 3] ^= 0x20                                   # flips the sign bit of signature 3: another curve point
sigs_inf = bytearray(ab.sigs)
sigs_inf[96 * 5:96 * 6] = bytes([0xC0]) + bytes(95)        # the infinity signature
for name, s, m in (("valid", ab.sigs, ab.msgs), ("swapped", ab.sigs, bytes(msgs)), ("negated", bytes(sigs_bad), ab.msgs),
                   ("infinity", bytes(sigs_inf), ab.msgs)):
    b.upload(s, m, ab.pks, ab.offsets, rands)
    b.enqueue()
    ok, st = b.result(want_status=True)
    print(name, ok, bytes(st).hex(), b.gt_bytes().hex() if not st.any() else "-")
print("launches", b.launches)
