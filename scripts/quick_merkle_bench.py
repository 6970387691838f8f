import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lighthouse_b200
from lighthouse_b200 import tree_hash as T, _ffi
from lighthouse_b200.synthetic import beacon_state_deneb_ssz
lighthouse_b200.init(0)
V = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
ssz = beacon_state_deneb_ssz(V, seed=42)
st = T.ResidentState(ssz)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts); s = ts.cuda_stream
for _ in range(3): st.enqueue(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K): st.enqueue(s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"resident root: {ms:.3f} ms/root, units={st.hash_units}, {st.hash_units/ms/1e6:.2f} G hash32_concat/s")
t = time.time(); r = T.beacon_state_root_deneb(ssz); print("e2e pageable ms", (time.time() - t) * 1e3, r.hex())
from tests import oracle_lib as O
O.set_threads(O.hw_threads()); t = time.time(); w, _ = O.beacon_state_root_deneb(ssz); print("oracle mt ms", (time.time()-t)*1e3, O.hw_threads(), w == r)
O.set_threads(1); t = time.time(); w, _ = O.beacon_state_root_deneb(ssz); print("oracle 1t ms", (time.time()-t)*1e3)
