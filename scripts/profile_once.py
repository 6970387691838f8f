"""One pass of each hot path for ncu (never a timing source)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lighthouse_b200
from lighthouse_b200 import bls, tree_hash as T
from lighthouse_b200.synthetic import attestation_batch, beacon_state_deneb_ssz
lighthouse_b200.init(0)
what = sys.argv[1] if len(sys.argv) > 1 else "bls"
if what == "bls":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    ab = attestation_batch(n, keys_per_set=128, n_validators=16384)
    b = bls.Batch(n, n * 128)
    b.upload(ab.sigs, ab.msgs, ab.pks, ab.offsets)
    for _ in range(2):
        b.enqueue(); print("verdict", b.result())
else:
    ssz = beacon_state_deneb_ssz(500_000, seed=42)
    st = T.ResidentState(ssz)
    for _ in range(2):
        print(st.root().hex())
