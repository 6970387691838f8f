"""From-spec SSZ hash_tree_root over hashlib (consensus-specs ssz/simple-serialize.md "Merkleization"), generic over
the type descriptors of lighthouse_b200/ssz_schema.py.  TEST INFRASTRUCTURE: an independent second implementation
used to pin oracle/ssz_sha256.c's hand-unrolled BeaconBlock restatement (the reference's own pin for block roots is
the EF ssz_static suite, which is not on disk)."""
import hashlib

from lighthouse_b200.ssz_schema import pack_bits, serialize

ZERO = [b"\0" * 32]
for _ in range(64):
    ZERO.append(hashlib.sha256(ZERO[-1] + ZERO[-1]).digest())


def _h(a, b):
    return hashlib.sha256(a + b).digest()


def merkleize(chunks, limit=None):
    """chunks: list of 32-byte values; limit: chunk-count limit (None = next_pow_of_two(len))."""
    n = len(chunks)
    if limit is None:
        limit = max(n, 1)
    assert n <= limit
    depth = max(limit - 1, 0).bit_length()
    if n == 0:
        return ZERO[depth]
    layer = list(chunks)
    for d in range(depth):
        if len(layer) % 2:
            layer.append(ZERO[d])
        layer = [_h(layer[i], layer[i + 1]) for i in range(0, len(layer), 2)]
    return layer[0]


def pack(b):
    b = bytes(b) + b"\0" * (-len(b) % 32)
    return [b[i:i + 32] for i in range(0, len(b), 32)]


def mix_in_length(root, n):
    return _h(root, n.to_bytes(32, "little"))


def is_basic(t):
    return t[0] == "uint"


def hash_tree_root(t, v):
    k = t[0]
    if k == "uint":
        return merkleize(pack(serialize(t, v)))
    if k == "bytes":
        return merkleize(pack(v), (t[1] + 31) // 32)
    if k == "bytelist":
        return mix_in_length(merkleize(pack(v), (t[1] + 31) // 32), len(v))
    if k == "bitvector":
        return merkleize(pack(pack_bits(v, False)), (t[1] + 255) // 256)
    if k == "bitlist":
        return mix_in_length(merkleize(pack(pack_bits(v, False)), (t[1] + 255) // 256), len(v))
    if k == "vector":
        if is_basic(t[1]):
            return merkleize(pack(b"".join(serialize(t[1], e) for e in v)), (t[2] * t[1][1] + 31) // 32)
        return merkleize([hash_tree_root(t[1], e) for e in v], t[2])
    if k == "list":
        if is_basic(t[1]):
            root = merkleize(pack(b"".join(serialize(t[1], e) for e in v)), (t[2] * t[1][1] + 31) // 32)
        else:
            root = merkleize([hash_tree_root(t[1], e) for e in v], t[2])
        return mix_in_length(root, len(v))
    return merkleize([hash_tree_root(ft, v[name]) for name, ft in t[1]])
