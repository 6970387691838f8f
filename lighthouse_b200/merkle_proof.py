"""Mirror of consensus/merkle_proof (MerkleTree::create / generate_proof / verify_merkle_proof,
/root/reference/consensus/merkle_proof/src/lib.rs:68-99,290-324,357-389) over the CUDA library."""
import ctypes as C

from ._ffi import lib, check, buf

MAX_TREE_DEPTH = 32


class MerkleTreeError(Exception):
    pass


class MerkleTree:
    """Right-sparse fixed-depth tree over `leaves` (list of 32-byte values)."""

    def __init__(self, leaves, depth):
        if depth > MAX_TREE_DEPTH or len(leaves) > (1 << depth):
            raise MerkleTreeError("DepthTooSmall" if len(leaves) > (1 << depth) else "Invalid")
        self.leaves = list(leaves)
        self.depth = depth

    @classmethod
    def create(cls, leaves, depth):
        return cls(leaves, depth)

    def push_leaf(self, elem, depth=None):
        if self.depth == 0:
            raise MerkleTreeError("DepthTooSmall")
        if len(self.leaves) >= (1 << self.depth):
            raise MerkleTreeError("MerkleTreeFull")
        self.leaves.append(elem)

    def _proof(self, index):
        root = C.create_string_buffer(32)
        branch = C.create_string_buffer(max(32 * self.depth, 1))
        p, keep = buf(b"".join(self.leaves))
        check(lib.lhb200_merkle_tree_proof(p, len(self.leaves), self.depth, index, root, branch),
              "lhb200_merkle_tree_proof")
        return root.raw, [branch.raw[32 * i: 32 * i + 32] for i in range(self.depth)]

    def hash(self):
        return self._proof(0)[0]

    def generate_proof(self, index, depth=None):
        """-> (leaf, branch bottom-up).  Leaf beyond the populated range is the zero chunk."""
        if index >= (1 << self.depth):
            raise MerkleTreeError("Invalid")
        _, branch = self._proof(index)
        leaf = self.leaves[index] if index < len(self.leaves) else b"\0" * 32
        return leaf, branch


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
