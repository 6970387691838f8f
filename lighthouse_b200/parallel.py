"""Multi-GPU plumbing for the hot paths (one process per GPU, torch.distributed).

BLS batches shard by independent SignatureSet: contiguous ranges balanced by the number of public keys
(the CSR offsets), every rank verifies its own range (its own random scalars and final exponentiation) and
one all-reduce(min) of a 1 x int32 verdict combines them.  State roots: whole states round-robin over ranks and
one all-gather of the 32-byte roots.  Backend: "nccl" on GPUs, "gloo" in the CPU tests.
"""
import numpy as np


def shard_ranges_by_keys(offsets, world):
    """Split sets [0, n) into `world` contiguous ranges with ~equal key counts.  Returns [(lo, hi)] * world."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        i = int(np.searchsorted(offsets, target, side="left"))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_of(sigs, msgs, pks, offsets, lo, hi):
    """Slice the SoA buffers of a flattened batch to sets [lo, hi) (offsets rebased to 0)."""
    offsets = np.asarray(offsets, dtype=np.uint32)
    k0, k1 = int(offsets[lo]), int(offsets[hi])
    return (sigs[96 * lo:96 * hi], msgs[32 * lo:32 * hi], pks[96 * k0:96 * k1],
            (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint32))


def allreduce_verdict(ok, device=None):
    """ncclAllReduce(min) of the per-shard verdicts; an empty shard contributes True (identity of min)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


def allgather_roots(root32, device=None):
    """all-gather of 32-byte roots -> list[bytes] in rank order."""
    import torch
    import torch.distributed as dist
    mine = torch.tensor(list(root32), dtype=torch.uint8, device=device)
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return [bytes(root32)]
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(x.cpu().tolist()) for x in out]


def verify_signature_sets_sharded(sigs, msgs, pks, offsets, verify_fn, device=None):
    """Each rank verifies its key-balanced shard with `verify_fn(sigs, msgs, pks, offsets) -> bool`, then one
    all-reduce(min).  n == 0 -> False on every rank (blst.rs:42-44)."""
    import torch.distributed as dist
    n = len(offsets) - 1
    if n == 0:
        return False
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_ranges_by_keys(offsets, world)[rank]
    ok = True if hi == lo else verify_fn(*shard_of(sigs, msgs, pks, offsets, lo, hi))
    return allreduce_verdict(ok, device)


def shard_ranges_by_bytes(sizes, world):
    """Split items [0, n) into `world` contiguous ranges with ~equal byte totals (blocks differ a lot in size)."""
    offs = np.concatenate([[0], np.cumsum(np.asarray(sizes, dtype=np.uint64))])
    return shard_ranges_by_keys(offs, world)


def beacon_block_roots_sharded(blocks, roots_fn=None, device=None):
    """canonical_root of every BeaconBlockDeneb in `blocks` (BASELINE configs[3]: the 32 blocks of a chain segment over
    the GPUs of one box): independent units, so each rank hashes a contiguous, byte-balanced slice in one pass and one
    all-gather of 32 B x len(blocks) returns all roots, in order, on every rank.  `roots_fn(list of ssz) -> list of 32-byte
    roots` defaults to the CUDA path (tree_hash.beacon_block_roots_deneb)."""
    import torch
    import torch.distributed as dist
    if roots_fn is None:
        from . import tree_hash as T
        roots_fn = T.beacon_block_roots_deneb
    n = len(blocks)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    ranges = shard_ranges_by_bytes([len(b) for b in blocks], world)
    lo, hi = ranges[rank]
    mine = roots_fn(blocks[lo:hi]) if hi > lo else []
    if world == 1:
        return [bytes(r) for r in mine]
    buf = torch.zeros(32 * n, dtype=torch.uint8, device=device)   # every rank fills its own slots; sum == gather
    if mine:
        buf[32 * lo:32 * hi] = torch.tensor(list(b"".join(mine)), dtype=torch.uint8, device=device)
    dist.all_reduce(buf.view(torch.int32) if buf.numel() % 4 == 0 else buf, op=dist.ReduceOp.SUM)
    raw = bytes(buf.cpu().tolist())
    return [raw[32 * i:32 * i + 32] for i in range(n)]


def beacon_state_root_sharded(ssz, device=None):
    """hash_tree_root(BeaconStateDeneb) with the big lists sharded over the ranks of the default process group:
    per-rank subtree roots -> one all-gather (32 bytes x lists x ranks) -> every rank folds the top.  world must be
    a power of two.  Every SHA-256 runs on the GPUs (lighthouse_b200.tree_hash.ShardedState)."""
    import torch
    import torch.distributed as dist
    from . import tree_hash as T
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    st = T.ShardedState(ssz, rank, world)
    mine = st.shard_roots()
    if world > 1:
        t = torch.tensor(list(mine), dtype=torch.uint8, device=device)
        out = torch.empty(world * len(mine), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(out, t)
        gathered = bytes(out.cpu().tolist())
    else:
        gathered = mine
    root = st.combine(gathered)
    st.release()
    return root
