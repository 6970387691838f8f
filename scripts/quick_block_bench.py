"""Times BeaconBlockDeneb roots through the C ABI (host buffers in/out) against the CPU oracle."""
import json, time
import lighthouse_b200
from lighthouse_b200 import synthetic, tree_hash
from tests import oracle_lib as O

lighthouse_b200.init(0)
blocks = [synthetic.beacon_block_deneb(seed=100 + i, n_transactions=150)[1] for i in range(32)]
for label, batch in (("1 block", blocks[:1]), ("32 blocks", blocks)):
    for _ in range(3):
        got = tree_hash.beacon_block_roots_deneb(batch)
    t = time.perf_counter()
    for _ in range(10):
        got = tree_hash.beacon_block_roots_deneb(batch)
    gpu_ms = (time.perf_counter() - t) * 100
    t = time.perf_counter()
    want = [O.beacon_block_root_deneb(b)[0] for b in batch]
    cpu_ms = (time.perf_counter() - t) * 1000
    print(json.dumps({"workload": label, "bytes": sum(len(b) for b in batch), "gpu_e2e_ms": round(gpu_ms, 3),
                      "cpu_oracle_1thread_ms": round(cpu_ms, 3), "match": got == want}))
