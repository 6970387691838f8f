"""TEST INFRASTRUCTURE: a direct restatement (hashlib, recursive enum-like nodes) of the reference's MerkleTree
(consensus/merkle_proof/src/lib.rs:27-345) — create, push_leaf, hash, finalize_deposits, get_finalized_hashes,
from_finalized_snapshot, generate_proof — used only to check lighthouse_b200.merkle_proof."""
import hashlib

ZERO = [bytes(32)]
for _ in range(64):
    ZERO.append(hashlib.sha256(ZERO[-1] + ZERO[-1]).digest())


class SpecError(Exception):
    pass


def h2(a, b):
    return hashlib.sha256(a + b).digest()


# nodes: ("Z", depth) | ("L", hash) | ("F", hash) | ("N", hash, left, right)
def create(leaves, depth):                                   # lib.rs:68-99
    if not leaves:
        return ("Z", depth)
    if depth == 0:
        return ("L", leaves[0])
    cap = 1 << (depth - 1)
    l, r = create(leaves[:cap], depth - 1), create(leaves[cap:], depth - 1)
    return ("N", h2(node_hash(l), node_hash(r)), l, r)


def node_hash(t):                                            # lib.rs:161-168
    return ZERO[t[1]] if t[0] == "Z" else t[1]


def push_leaf(t, elem, depth):                               # lib.rs:103-158
    if depth == 0:
        raise SpecError("DepthTooSmall")
    if t[0] in ("L", "F"):
        raise SpecError("LeafReached" if t[0] == "L" else "FinalizedNodePushed")
    if t[0] == "Z":
        return create([elem], depth)
    _, _, l, r = t
    if l[0] in ("L",) and r[0] == "L":
        raise SpecError("MerkleTreeFull")
    # left is a full subtree (or finalized) -> go right
    def full(x, d):
        if x[0] in ("L", "F"):
            return True
        if x[0] == "Z":
            return False
        return full(x[3], d - 1)
    if full(l, depth - 1):
        if full(r, depth - 1):
            raise SpecError("MerkleTreeFull")
        r = create([elem], depth - 1) if r[0] == "Z" else push_leaf(r, elem, depth - 1)
    else:
        l = create([elem], depth - 1) if l[0] == "Z" else push_leaf(l, elem, depth - 1)
    return ("N", h2(node_hash(l), node_hash(r)), l, r)


def finalize(t, n, level):                                   # lib.rs:185-219
    if t[0] == "F":
        return t
    if t[0] == "Z":
        raise SpecError("ZeroNodeFinalized")
    if t[0] == "L":
        if level != 0:
            raise SpecError("PleaseNotifyTheDevs")
        return ("F", t[1])
    _, hsh, l, r = t
    if level == 0:
        raise SpecError("PleaseNotifyTheDevs")
    deposits = 1 << level
    if deposits <= n:
        return ("F", hsh)
    l = finalize(l, n, level - 1)
    if n > deposits // 2:
        r = finalize(r, n - deposits // 2, level - 1)
    return ("N", hsh, l, r)


def finalized_hashes(t):                                     # lib.rs:221-236
    if t[0] in ("Z", "L"):
        return []
    if t[0] == "F":
        return [t[1]]
    return finalized_hashes(t[2]) + finalized_hashes(t[3])


def from_snapshot(branch, count, level):                     # lib.rs:238-288
    if not branch:
        if count == 0:
            return ("Z", level)
        raise SpecError("EmptyBranchWithNonZeroDeposits")
    if count == (1 << level):
        return ("F", branch[0])
    if level == 0:
        raise SpecError("EndOfTree")
    half = 1 << (level - 1)
    if count >= half:
        l, r = ("F", branch[0]), from_snapshot(branch[1:], count - half, level - 1)
    else:
        l, r = from_snapshot(branch, count, level - 1), ("Z", level - 1)
    return ("N", h2(node_hash(l), node_hash(r)), l, r)


def generate_proof(t, index, depth):                         # lib.rs:290-324
    proof, cur, d = [], t, depth
    while d > 0:
        bit = (index >> (d - 1)) & 1
        if cur[0] == "F":
            raise SpecError("ProofEncounteredFinalizedNode")
        if cur[0] == "Z":
            l = r = ("Z", cur[1] - 1)
        else:
            l, r = cur[2], cur[3]
        if bit:
            proof.append(node_hash(l)); cur = r
        else:
            proof.append(node_hash(r)); cur = l
        d -= 1
    return node_hash(cur), proof[::-1]
