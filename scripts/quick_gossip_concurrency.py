"""Steady state of the reference's gossip path: T blocking workers (beacon_processor/src/lib.rs:256,1396), each verifying
its own 64-set batch through the plugin call lhb200_verify_signature_sets (host buffers, pooled handles), back to back.
Prints one JSON line per T: batches/s, sets/s, mean latency.  usage: python scripts/quick_gossip_concurrency.py [keys_per_set]"""
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lighthouse_b200
from lighthouse_b200 import bls
from lighthouse_b200.synthetic import attestation_batch

lighthouse_b200.init(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
N_SETS, SECONDS = 64, 1.5
batches = [attestation_batch(N_SETS, keys_per_set=k, n_validators=2048, seed=900 + t) for t in range(32)]
for ab in batches[:2]:
    assert bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, ab.offsets)
for T in (1, 2, 4, 8, 16, 32):
    counts = [0] * T
    stop = time.perf_counter() + SECONDS
    def work(t):
        ab = batches[t]
        while time.perf_counter() < stop:
            assert bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, ab.offsets)
            counts[t] += 1
    # one untimed concurrent pass creates the pooled handles
    warm = [threading.Thread(target=lambda t=t: bls.verify_signature_sets_raw(batches[t].sigs, batches[t].msgs, batches[t].pks, batches[t].offsets)) for t in range(T)]
    [w.start() for w in warm]; [w.join() for w in warm]
    stop = time.perf_counter() + SECONDS
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.perf_counter() - t0
    n = sum(counts)
    print(json.dumps({"workers": T, "keys_per_set": k, "batches_per_s": n / dt, "sets_per_s": n * N_SETS / dt,
                      "mean_latency_ms": dt * T / max(n, 1) * 1e3}))
