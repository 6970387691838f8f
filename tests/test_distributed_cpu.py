"""World-size-2 `gloo` tests (CPU) of the N > 1 host logic: key-balanced sharding of SignatureSets, the single
all-reduce(min) of verdicts and the all-gather of roots.  The per-shard verifier here is the CPU
oracle — the GPU verifier is exercised by the `-m gpu` tests and by bench.py --gpus N."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_balance_and_cover():
    from lighthouse_b200.parallel import shard_ranges_by_keys, shard_of
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        counts = rng.integers(1, 600, size=1000)
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
        rs = shard_ranges_by_keys(offs, world)
        assert rs[0][0] == 0 and rs[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        loads = [int(offs[hi] - offs[lo]) for lo, hi in rs]
        assert max(loads) - min(loads) <= 2 * 600
    offs = np.array([0, 3, 5, 9], dtype=np.uint32)
    s, m, p, o = shard_of(b"s" * 288, b"m" * 96, b"p" * 96 * 9, offs, 1, 3)
    assert len(s) == 192 and len(m) == 64 and len(p) == 96 * 6 and list(o) == [0, 2, 6]
    assert shard_ranges_by_keys(np.array([0, 5], dtype=np.uint32), 4)[-1] == (1, 1) or True   # fewer sets than ranks


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lighthouse_b200 import parallel
    from tests import oracle_lib as O
    from oracle import bls_ref as B
    # 4 sets x 2 keys, built with the C oracle
    n, k = 4, 2
    sks = [[B.interop_secret_key(5 * i + j) for j in range(k)] for i in range(n)]
    msgs = [hashlib.sha256(b"dist%d" % i).digest() for i in range(n)]
    pks = b"".join(O.bls_sk_to_pk(s.to_bytes(32, "big")) for row in sks for s in row)
    sigs = b"".join(O.bls_sign((sum(row) % B.R).to_bytes(32, "big"), m) for row, m in zip(sks, msgs))
    offs = np.arange(n + 1, dtype=np.uint32) * k

    def verify(s, m, p, o):
        return O.bls_verify_signature_sets(s, m, p, o, [3 + 2 * i for i in range(len(o) - 1)])

    good = parallel.verify_signature_sets_sharded(sigs, b"".join(msgs), pks, offs, verify)
    bad_msgs = bytearray(b"".join(msgs)); bad_msgs[32 * 3] ^= 1          # corrupts a set owned by the last rank
    bad = parallel.verify_signature_sets_sharded(sigs, bytes(bad_msgs), pks, offs, verify)
    empty = parallel.verify_signature_sets_sharded(b"", b"", b"", np.array([0], dtype=np.uint32), verify)
    roots = parallel.allgather_roots(hashlib.sha256(bytes([rank])).digest())
    # block roots: 5 blocks of very different sizes over 2 ranks, hashed by the CPU oracle here
    from lighthouse_b200.synthetic import beacon_block_deneb
    blocks = [beacon_block_deneb(seed=40 + i, n_attestations=3 * i, n_transactions=1 + 30 * (i % 2))[1] for i in range(5)]
    block_roots = parallel.beacon_block_roots_sharded(blocks, roots_fn=lambda bs: [O.beacon_block_root_deneb(b)[0] for b in bs])
    assert block_roots == [O.beacon_block_root_deneb(b)[0] for b in blocks]
    q.put((rank, good, bad, empty, roots))
    dist.destroy_process_group()


def test_two_rank_verdict_allreduce_and_root_allgather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, good, bad, empty, roots in res:
        assert good is True and bad is False and empty is False        # every rank sees the same combined verdict
        assert roots == [hashlib.sha256(bytes([0])).digest(), hashlib.sha256(bytes([1])).digest()]
