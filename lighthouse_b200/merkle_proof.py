"""Mirror of consensus/merkle_proof (MerkleTree::create / generate_proof / verify_merkle_proof,
/root/reference/consensus/merkle_proof/src/lib.rs:68-99,290-324,357-389) over the CUDA library."""
import ctypes as C

from ._ffi import lib, check, buf

MAX_TREE_DEPTH = 32


class MerkleTreeError(Exception):
    pass


def _hash_pairs(pairs: bytes) -> bytes:
    """n x 64 bytes -> n x 32 bytes on the device (lhb200_hash_pairs)."""
    n = len(pairs) // 64
    if n == 0:
        return b""
    out = C.create_string_buffer(32 * n)
    p, keep = buf(pairs)
    check(lib.lhb200_hash_pairs(p, out, n), "lhb200_hash_pairs")
    return out.raw


def _zero_hash(level: int) -> bytes:
    out = C.create_string_buffer(32)
    check(lib.lhb200_zero_hash(level, out), "lhb200_zero_hash")
    return out.raw


class MerkleTree:
    """Right-sparse fixed-depth tree (consensus/merkle_proof/src/lib.rs:27-45) with a FINALIZED prefix:
    the first `finalized_count` leaves are represented only by the hashes of their maximal aligned subtrees
    (`MerkleTree::Finalized` nodes, one per set bit of the count, left to right), the rest are explicit leaves.
    All hashing runs on the device."""

    def __init__(self, leaves, depth, finalized=None, finalized_count=0):
        if depth > MAX_TREE_DEPTH or finalized_count + len(leaves) > (1 << depth):
            raise MerkleTreeError("DepthTooSmall" if finalized_count + len(leaves) > (1 << depth) else "Invalid")
        self.leaves = list(leaves)          # explicit leaves, indices finalized_count ...
        self.depth = depth
        self.finalized = list(finalized or [])
        self.finalized_count = finalized_count

    @classmethod
    def create(cls, leaves, depth):
        return cls(leaves, depth)

    def __len__(self):
        return self.finalized_count + len(self.leaves)

    def push_leaf(self, elem, depth=None):
        if self.depth == 0:
            raise MerkleTreeError("DepthTooSmall")
        if len(self) >= (1 << self.depth):
            raise MerkleTreeError("MerkleTreeFull")
        self.leaves.append(elem)

    # ---- level arrays: level l holds the nodes with index >= start_l (start_l even, or 0), the first one being the
    # finalized maximal subtree of that level when bit l of the finalized count is set
    def _levels(self):
        F, depth = self.finalized_count, self.depth
        fin = {}
        hashes = list(self.finalized)
        for l in range(depth, -1, -1):          # left to right = descending levels of the set bits of F
            if (F >> l) & 1:
                fin[l] = hashes.pop(0)
        levels = []
        nodes, start = list(self.leaves), F
        for l in range(depth + 1):
            if l in fin and l < depth:
                nodes, start = [fin[l]] + nodes, start - 1
            elif l in fin:                       # F == 2^depth: the root itself is finalized
                nodes, start = [fin[l]], 0
            levels.append((start, nodes))
            if l == depth:
                break
            if len(nodes) % 2:
                nodes = nodes + [_zero_hash(l)]
            out = _hash_pairs(b"".join(nodes))
            nodes, start = [out[32 * i:32 * i + 32] for i in range(len(nodes) // 2)], start // 2
        return levels

    def _proof(self, index):
        if not self.finalized_count:             # no finalized prefix: one fused device pass
            root = C.create_string_buffer(32)
            branch = C.create_string_buffer(max(32 * self.depth, 1))
            p, keep = buf(b"".join(self.leaves))
            check(lib.lhb200_merkle_tree_proof(p, len(self.leaves), self.depth, index, root, branch),
                  "lhb200_merkle_tree_proof")
            return root.raw, [branch.raw[32 * i: 32 * i + 32] for i in range(self.depth)]
        levels = self._levels()
        start, top = levels[self.depth]
        root = top[0] if top else _zero_hash(self.depth)
        branch = []
        idx = index
        for l in range(self.depth):
            start, nodes = levels[l]
            sib = (idx ^ 1) - start
            branch.append(nodes[sib] if 0 <= sib < len(nodes) else _zero_hash(l))
            idx >>= 1
        return root, branch

    def hash(self):
        if not self.finalized_count:
            return self._proof(0)[0]
        top = self._levels()[self.depth][1]
        return top[0] if top else _zero_hash(self.depth)

    def generate_proof(self, index, depth=None):
        """-> (leaf, branch bottom-up).  Leaf beyond the populated range is the zero chunk; a leaf inside the finalized
        prefix raises ProofEncounteredFinalizedNode (lib.rs:300-302)."""
        if index >= (1 << self.depth):
            raise MerkleTreeError("Invalid")
        if index < self.finalized_count:
            raise MerkleTreeError("ProofEncounteredFinalizedNode")
        _, branch = self._proof(index)
        k = index - self.finalized_count
        leaf = self.leaves[k] if k < len(self.leaves) else b"\0" * 32
        return leaf, branch

    def finalize_deposits(self, deposits_to_finalize, level=None):
        """lib.rs:185-219: every subtree the reference's recursion marks Finalized.  It descends from the root: a node
        of 2^l leaves with 2^l <= count is finalized whole (even when partly populated: its hash then covers zero
        leaves), otherwise its left child is visited and, when count exceeds half, its right child with the rest —
        which is the error ZeroNodeFinalized if that child holds no leaf.  A count of 0 still finalizes leaf 0 (the
        Leaf arm has no count check)."""
        n = max(int(deposits_to_finalize), 1)
        m = len(self)
        if m == 0:
            raise MerkleTreeError("ZeroNodeFinalized")
        pos, rem, lvl = 0, n, self.depth
        while True:                                   # the reference's walk, to find the Zero-node error cases
            if (1 << lvl) <= rem or lvl == 0:
                break
            half = 1 << (lvl - 1)
            if rem > half:
                if m <= pos + half:                   # right child is Zero(lvl - 1)
                    raise MerkleTreeError("ZeroNodeFinalized")
                pos, rem = pos + half, rem - half
            lvl -= 1
        n = min(n, 1 << self.depth)
        if n <= self.finalized_count:
            return
        levels = self._levels()
        hashes, pos = [], 0
        for l in range(self.depth, -1, -1):
            if (n >> l) & 1:
                start, nodes = levels[l]
                k = (pos >> l) - start
                hashes.append(nodes[k] if 0 <= k < len(nodes) else _zero_hash(l))
                pos += 1 << l
        self.leaves = self.leaves[n - self.finalized_count:]
        self.finalized, self.finalized_count = hashes, n

    def get_finalized_hashes(self):
        """lib.rs:232-236"""
        return list(self.finalized)

    @classmethod
    def from_finalized_snapshot(cls, finalized_branch, deposit_count, level):
        """lib.rs:238-288: rebuild a tree whose first `deposit_count` leaves are finalized."""
        branch = list(finalized_branch)

        def walk(br, count, lvl):                # returns the hashes the reference's recursion actually consumes
            if not br:
                if count == 0:
                    return []
                raise MerkleTreeError(f"InvalidSnapshot(EmptyBranchWithNonZeroDeposits({count}))")
            if count == (1 << lvl):
                return [br[0]]
            if lvl == 0:
                raise MerkleTreeError("InvalidSnapshot(EndOfTree)")
            half = 1 << (lvl - 1)
            if count >= half:
                return [br[0]] + walk(br[1:], count - half, lvl - 1)
            return walk(br, count, lvl - 1)

        used = walk(branch, deposit_count, level)
        return cls([], level, finalized=used, finalized_count=deposit_count if used else 0)


def verify_merkle_proofs(leaves, branches, depth, indices, roots):
    """Batch of verify_merkle_proof; returns list[bool]."""
    import numpy as np
    n = len(leaves)
    if n == 0:
        return []
    for b in branches:
        if len(b) != depth:  # merkle_proof/src/lib.rs:364
            raise ValueError("branch length != depth")
    lp, k1 = buf(b"".join(leaves))
    bp, k2 = buf(b"".join(b"".join(b) for b in branches) or b"\0")
    idx = np.asarray(indices, dtype=np.uint64)
    rp, k3 = buf(b"".join(roots))
    ok = C.create_string_buffer(n)
    check(lib.lhb200_verify_merkle_proofs(lp, bp, depth, C.c_void_p(idx.ctypes.data), rp, n, ok),
          "lhb200_verify_merkle_proofs")
    return [bool(x) for x in ok.raw]


def verify_merkle_proof(leaf, branch, depth, index, root):
    if len(branch) != depth:
        return False
    return verify_merkle_proofs([leaf], [branch], depth, [index], [root])[0]
