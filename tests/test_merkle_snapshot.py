"""GPU: deposit-tree snapshot API of consensus/merkle_proof (finalize_deposits, get_finalized_hashes,
from_finalized_snapshot, lib.rs:185-288) — the mirror in lighthouse_b200.merkle_proof (device hashing) against the
recursive restatement in tests/merkle_tree_spec.py, the way common/deposit_data_tree.rs:55-95 drives it."""
import hashlib

import pytest

from tests import merkle_tree_spec as SP

pytestmark = pytest.mark.gpu


def leaves_of(n, seed=b"dep"):
    return [hashlib.sha256(seed + i.to_bytes(4, "little")).digest() for i in range(n)]


@pytest.mark.parametrize("depth,n", [(3, 5), (5, 23), (10, 700), (32, 100)])
def test_finalize_snapshot_roundtrip(gpu, depth, n):
    from lighthouse_b200.merkle_proof import MerkleTree, MerkleTreeError
    lv = leaves_of(n)
    tree = MerkleTree.create(lv, depth)
    spec = SP.create(lv, depth)
    assert tree.hash() == SP.node_hash(spec)
    for fin in sorted({1, 2, 3, n // 2, n - 1, n} - {0}):
        if fin > n:
            continue
        tree.finalize_deposits(fin, depth)
        spec = SP.finalize(spec, fin, depth)
        assert tree.get_finalized_hashes() == SP.finalized_hashes(spec), (depth, n, fin)
        assert tree.hash() == SP.node_hash(spec)
        # proofs: finalized leaves are refused, later ones match the reference's branch
        with pytest.raises(MerkleTreeError):
            tree.generate_proof(fin - 1, depth)
        for idx in sorted({fin, (fin + n) // 2, n - 1, n}):
            if idx < fin or idx >= (1 << depth):
                continue
            assert tree.generate_proof(idx, depth) == SP.generate_proof(spec, idx, depth), (fin, idx)
        # snapshot -> tree -> same root; pushing further deposits keeps agreeing
        snap = MerkleTree.from_finalized_snapshot(tree.get_finalized_hashes(), fin, depth)
        sspec = SP.from_snapshot(SP.finalized_hashes(spec), fin, depth)
        assert snap.hash() == SP.node_hash(sspec)
        for extra in leaves_of(3, b"more"):
            if len(snap) < (1 << depth):
                snap.push_leaf(extra, depth)
                sspec = SP.push_leaf(sspec, extra, depth)
                assert snap.hash() == SP.node_hash(sspec)
        assert snap.generate_proof(fin, depth) == SP.generate_proof(sspec, fin, depth)


def test_finalize_and_snapshot_errors(gpu):
    from lighthouse_b200.merkle_proof import MerkleTree, MerkleTreeError
    # every (leaf count, finalize count) of a depth-4 tree: same Ok / ZeroNodeFinalized outcome and the same finalized
    # hashes as the reference's recursion (a partly populated subtree IS finalized whole when the count covers it)
    for m in range(0, 12):
        for n in range(0, 17):
            t = MerkleTree.create(leaves_of(m), 4)
            try:
                want = SP.finalized_hashes(SP.finalize(SP.create(leaves_of(m), 4), n, 4))
            except SP.SpecError:
                want = None
            if want is None:
                with pytest.raises(MerkleTreeError):
                    t.finalize_deposits(n, 4)
            else:
                t.finalize_deposits(n, 4)
                assert t.get_finalized_hashes() == want, (m, n)
                assert t.hash() == SP.node_hash(SP.create(leaves_of(m), 4))
    with pytest.raises(MerkleTreeError):                      # EmptyBranchWithNonZeroDeposits (lib.rs:243-248)
        MerkleTree.from_finalized_snapshot([], 3, 4)
    with pytest.raises(MerkleTreeError):                      # EndOfTree (lib.rs:258-260)
        MerkleTree.from_finalized_snapshot([bytes(32)], 0, 4)
    z = MerkleTree.from_finalized_snapshot([], 0, 4)
    assert z.hash() == SP.ZERO[4]
    full = MerkleTree.create(leaves_of(16), 4)
    full.finalize_deposits(16, 4)
    assert full.get_finalized_hashes() == [SP.node_hash(SP.create(leaves_of(16), 4))] and full.hash() == full.get_finalized_hashes()[0]
