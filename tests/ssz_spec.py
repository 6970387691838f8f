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


# ---- generic SSZ decoder (test side only): lets the spec merkleization run on SSZ blobs the generators emit as bytes
from lighthouse_b200.ssz_schema import fixed_size, is_fixed  # noqa: E402


def deserialize(t, b):
    b = bytes(b)
    k = t[0]
    if k == "uint":
        assert len(b) == t[1]
        return int.from_bytes(b, "little")
    if k == "bytes":
        assert len(b) == t[1]
        return b
    if k == "bytelist":
        assert len(b) <= t[1]
        return b
    if k == "bitvector":
        assert len(b) == (t[1] + 7) // 8
        return [bool((b[i // 8] >> (i % 8)) & 1) for i in range(t[1])]
    if k == "bitlist":
        assert b and b[-1]
        n = 8 * (len(b) - 1) + b[-1].bit_length() - 1
        return [bool((b[i // 8] >> (i % 8)) & 1) for i in range(n)]
    if k in ("vector", "list"):
        et = t[1]
        if is_fixed(et):
            sz = fixed_size(et)
            assert len(b) % sz == 0
            return [deserialize(et, b[i:i + sz]) for i in range(0, len(b), sz)]
        if not b:
            return []
        first = int.from_bytes(b[:4], "little")
        offs = [int.from_bytes(b[4 * i:4 * i + 4], "little") for i in range(first // 4)] + [len(b)]
        return [deserialize(et, b[offs[i]:offs[i + 1]]) for i in range(len(offs) - 1)]
    # container
    out, pos, var = {}, 0, []
    for name, ft in t[1]:
        if is_fixed(ft):
            sz = fixed_size(ft)
            out[name] = deserialize(ft, b[pos:pos + sz])
            pos += sz
        else:
            var.append((name, ft, int.from_bytes(b[pos:pos + 4], "little")))
            pos += 4
    for i, (name, ft, off) in enumerate(var):
        end = var[i + 1][2] if i + 1 < len(var) else len(b)
        out[name] = deserialize(ft, b[off:end])
    return out
