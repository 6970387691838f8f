"""Warm (dirty-path) vs cold BeaconState root on a resident 500k-validator Deneb state."""
import json, struct, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lighthouse_b200
from lighthouse_b200 import tree_hash as T
from lighthouse_b200.synthetic import beacon_state_deneb_ssz
from tests import oracle_lib as O
lighthouse_b200.init(0)
n = 500_000
ssz = bytearray(beacon_state_deneb_ssz(n, seed=42))
st = T.ResidentState(bytes(ssz))
ts = torch.cuda.Stream(); s = ts.cuda_stream
def timed(f, reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): f()
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        e0.record(ts)
        for _ in range(reps): f()
        e1.record(ts)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cold = timed(lambda: st.enqueue(s))
st.enable_incremental(); st.root()
o_val, o_bal = struct.unpack_from("<II", ssz, 524552)
rng = np.random.default_rng(1)
res = {"n_validators": n, "cold_resident_ms": round(cold, 4), "cold_hash_units": int(st.hash_units)}
for n_dirty in (0, 64, 2048, 16384):
    def slot():
        edits = []
        for vi in rng.choice(n, size=n_dirty, replace=False) if n_dirty else []:
            off = o_val + 121 * int(vi) + 80; d = struct.pack("<Q", int(rng.integers(1, 1 << 40))); ssz[off:off+8] = d; edits.append((off, d))
            off = o_bal + 8 * int(vi); ssz[off:off+8] = d; edits.append((off, d))
        st.patch_batch(edits)
    # device time of the warm root alone (patches applied beforehand)
    ms = []
    for _ in range(5):
        slot()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(ts):
            e0.record(ts); st.enqueue(s); e1.record(ts)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t0 = time.perf_counter(); slot(); r = st.root(); e2e = (time.perf_counter() - t0) * 1e3
    ok = r == O.beacon_state_root_deneb(bytes(ssz))[0]
    res[f"warm_{n_dirty}_validators+balances"] = {"device_ms": round(sorted(ms)[2], 4), "hashes": int(st.last_root_hashes),
                                                  "patch+root_host_ms": round(e2e, 3), "matches_oracle": ok}
print(json.dumps(res))
