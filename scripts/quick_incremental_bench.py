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
import ctypes as C
from lighthouse_b200 import _ffi
mv = np.frombuffer(ssz, dtype=np.uint8)
for n_dirty in (0, 64, 2048, 16384):
    def make_slot():
        """A slot's worth of mutations as the C ABI takes them (offsets, lengths, one blob), built OUTSIDE every timed region."""
        vi = rng.choice(n, size=n_dirty, replace=False).astype(np.uint64) if n_dirty else np.zeros(0, dtype=np.uint64)
        vals = rng.integers(1, 1 << 40, size=n_dirty, dtype=np.uint64)
        offs = np.empty(2 * n_dirty, dtype=np.uint64)
        offs[0::2] = o_val + 121 * vi + 80
        offs[1::2] = o_bal + 8 * vi
        blob = np.repeat(vals, 2).view(np.uint8).copy()
        for k in range(2 * n_dirty):                 # mirror into the host copy for the oracle check
            o = int(offs[k]); mv[o:o + 8] = blob[8 * k:8 * k + 8]
        return offs, np.full(2 * n_dirty, 8, dtype=np.uint32), blob
    def apply(sl):
        offs, lens, blob = sl
        if len(offs):
            _ffi.check(_ffi.lib.lhb200_state_patch_batch(st._h, offs.ctypes.data, lens.ctypes.data, blob.ctypes.data, len(offs)), "patch_batch")
    def slot():
        apply(make_slot())
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
    sl = make_slot()
    t0 = time.perf_counter(); apply(sl); t1 = time.perf_counter(); r = st.root(); e2e = (time.perf_counter() - t0) * 1e3
    patch_ms = (t1 - t0) * 1e3
    ok = r == O.beacon_state_root_deneb(bytes(ssz))[0]
    res[f"warm_{n_dirty}_validators+balances"] = {"device_ms": round(sorted(ms)[2], 4), "hashes": int(st.last_root_hashes),
                                                  "patch_batch_wall_ms": round(patch_ms, 3), "patch+root_wall_ms": round(e2e, 3), "matches_oracle": ok,
                                                  "note": "wall clock of lhb200_state_patch_batch + lhb200_state_root on pre-built arrays (round 1 timed the Python loop that built the edits)"}
print(json.dumps(res))
