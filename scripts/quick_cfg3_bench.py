"""cfg3 shape on one GPU: the SignatureSets of 32 full Deneb blocks (1/1/256 x128/512/1 x32 keys per block), resident.
usage: python scripts/quick_cfg3_bench.py [n_blocks] [n_validators]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import lighthouse_b200
from lighthouse_b200 import bls
from lighthouse_b200 import synthetic as S

lighthouse_b200.init(0)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
kc = S.block_signature_key_counts(nb, nv)
t = time.time()
tab = S.interop_pubkey_table(nv)
w = S.sets_workload(kc, nv, seed=3, first_index=0)
a = S.materialize_sets(w, tab, bls.sign)
print(f"gen {len(kc)} sets / {int(w['offsets'][-1])} keys: {time.time() - t:.1f}s")
b = bls.Batch(len(kc), int(w["offsets"][-1]))
b.upload(a.sigs, a.msgs, a.pks, a.offsets)
ts = torch.cuda.Stream()
s = ts.cuda_stream
b.enqueue(s)
print("verdict", b.result(s))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(ts):
    e0.record(ts)
    b.enqueue(s)
    e1.record(ts)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"resident verify: {ms:.2f} ms for {len(kc)} sets -> {len(kc) / ms * 1e3:.0f} sets/s; launches {b.launches}")
