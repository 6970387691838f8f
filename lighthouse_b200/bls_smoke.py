"""smoke(): one tiny batch BLS verification on cuda:0, checked against the big-integer oracle."""
import hashlib

import numpy as np


def run():
    from . import bls
    from oracle import bls_ref as B  # checker only
    sks = [7, 11, 13]
    msg = hashlib.sha256(b"smoke").digest()
    pk96 = b"".join(B.g1_uncompressed(B.sk_to_pk(s)) for s in sks)
    sig = B.g2_compress(B.sign(sum(sks) % B.R, msg))
    ok = bls.verify_signature_sets_raw(sig, msg, pk96, np.array([0, 3], dtype=np.uint32))
    bad = bls.verify_signature_sets_raw(sig, hashlib.sha256(b"other").digest(), pk96, np.array([0, 3], dtype=np.uint32))
    assert ok and not bad, (ok, bad)
    assert bls.sign((7).to_bytes(32, "big"), msg) == B.g2_compress(B.sign(7, msg))
    print("smoke: verify_signature_sets(1 set x 3 keys) True / tampered False == oracle; device sign == oracle")
