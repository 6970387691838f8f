#!/usr/bin/env python3
"""Multi-GPU parity check of the library-owned NCCL path (run under torchrun, one rank per GPU):
   * lhb200_verify_signature_sets_collective: key-balanced shards, verdict all-reduced inside the library; a bad set in
     ONE shard must flip the verdict on EVERY rank;
   * lhb200_state_root_sharded: one BeaconState over all ranks == the single-GPU root == the CPU oracle.
torch.distributed is used only to broadcast the 128-byte NCCL id and to compare results across ranks."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # control plane only: the data path must not need torch's NCCL
    import lighthouse_b200
    from lighthouse_b200 import _ffi, bls, tree_hash as T, parallel as PAR, synthetic as S
    lighthouse_b200.init(local)
    lib = _ffi.lib
    ident = (C.c_uint8 * 128)()
    if rank == 0:
        _ffi.check(lib.lhb200_comm_unique_id(ident), "unique_id")
    obj = [bytes(ident)]
    dist.broadcast_object_list(obj, src=0)
    ident = (C.c_uint8 * 128).from_buffer_copy(obj[0])
    _ffi.check(lib.lhb200_comm_init(rank, world, ident), "comm_init")
    r, w = C.c_int32(), C.c_int32()
    lib.lhb200_comm_info(C.byref(r), C.byref(w))
    assert (r.value, w.value) == (rank, world)

    # ---- BLS: 600 ragged sets, sharded by keys
    kc = np.random.default_rng(5).integers(1, 40, size=600)
    tab = S.interop_pubkey_table(1024)
    work = S.sets_workload(kc, 1024, seed=77)
    ab = S.materialize_sets(work, tab, bls.sign)
    ranges = PAR.shard_ranges_by_keys(np.concatenate([[0], np.cumsum(kc)]), world)
    lo, hi = ranges[rank]
    ok = C.create_string_buffer(1)

    def collective(sigs):
        s, m, p, o = PAR.shard_of(sigs, ab.msgs, ab.pks, ab.offsets, lo, hi)
        o = np.ascontiguousarray(o, dtype=np.uint32)
        bs, k1 = _ffi.buf(s if len(s) else b"\0"); bm, k2 = _ffi.buf(m if len(m) else b"\0"); bp, k3 = _ffi.buf(p if len(p) else b"\0")
        _ffi.check(lib.lhb200_verify_signature_sets_collective(bs, bm, bp, o.ctypes.data, None, hi - lo, ok), "collective verify")
        return ok.raw[0] == 1

    assert collective(ab.sigs) is True
    bad = bytearray(ab.sigs)
    j = ranges[world - 1][0]                     # first set of the LAST rank's shard gets its neighbour's signature
    bad[96 * j:96 * j + 96] = ab.sigs[96 * (j + 1):96 * (j + 2)]
    assert collective(bytes(bad)) is False, "a bad set in one shard must fail the batch on every rank"
    assert collective(ab.sigs) is True

    # ---- one BeaconState over all ranks
    if world & (world - 1) == 0:
        ssz = S.beacon_state_deneb_ssz(50_000, seed=11)
        sh = T.ShardedState(ssz, rank, world)
        got = sh.root_collective()
        got2 = sh.root_collective()
        sh.release()
        want = T.beacon_state_root_deneb(ssz)
        assert got == want and got2 == want, (got.hex(), want.hex())
        if rank == 0:
            from tests import oracle_lib as O
            assert O.beacon_state_root_deneb(ssz)[0] == want
    roots = [None] * world
    dist.all_gather_object(roots, "ok")
    lib.lhb200_comm_destroy()
    if rank == 0:
        print(f"mgpu_check: world={world} collective verify (valid / one bad shard / valid) and sharded state root == single-GPU root == oracle: OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
