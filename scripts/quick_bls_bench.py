import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lighthouse_b200
from lighthouse_b200 import bls
from lighthouse_b200.synthetic import attestation_batch
lighthouse_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = int(sys.argv[2]) if len(sys.argv) > 2 else 128
t = time.time(); ab = attestation_batch(n, keys_per_set=k, n_validators=16384); print(f"gen {n} sets: {time.time()-t:.2f}s")
b = bls.Batch(n, n * k)
b.upload(ab.sigs, ab.msgs, ab.pks, ab.offsets)
ts = torch.cuda.Stream(); s = ts.cuda_stream
b.enqueue(s); print("verdict", b.result(s))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(ts):
    e0.record(ts); b.enqueue(s); e1.record(ts)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"resident verify: {ms:.2f} ms for {n} sets -> {n/ms*1e3:.0f} sets/s; launches {b.launches}")
