"""GPU stage-by-stage parity for the BLS pipeline (every SURVEY §2d kernel stage against the oracle),
through the lhb200_debug_bls test hook."""
import hashlib
import random

import pytest

from oracle import bls_ref as B

pytestmark = pytest.mark.gpu
P = B.P


def be(x): return x.to_bytes(48, "big")
def b2(a): return be(a[0]) + be(a[1])
def f2_from(r): return (int.from_bytes(r[:48], "big"), int.from_bytes(r[48:96], "big"))


@pytest.fixture(scope="module")
def dbg(gpu):
    from lighthouse_b200 import bls
    return bls.debug_stage


def test_fp2_ops(dbg):
    rnd = random.Random(2)
    vals = [(0, 0), (1, 0), (0, 1), (P - 1, P - 1), (5, 0), (0, 7)] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(10)]
    for a in vals:
        b = vals[rnd.randrange(len(vals))]
        assert f2_from(dbg(5, bytes([0]) + b2(a) + b2(b), 96)[1]) == B.f2_mul(a, b)
        assert f2_from(dbg(5, bytes([1]) + b2(a) + b2(b), 96)[1]) == B.f2_sqr(a)
        if a != (0, 0):
            assert f2_from(dbg(5, bytes([2]) + b2(a) + b2(b), 96)[1]) == B.f2_inv(a)
        ok, r = dbg(5, bytes([3]) + b2(a) + b2(b), 96)
        assert bool(ok) == (B.f2_sqrt(a) is not None)
        if ok:
            assert B.f2_sqr(f2_from(r)) == a
        assert dbg(5, bytes([4]) + b2(a) + b2(b), 96)[0] == B.f2_sgn0(a)


def test_expand_and_sswu_and_hash_to_g2(dbg):
    for msg in (bytes(range(32)), bytes(32), hashlib.sha256(b"x").digest()):
        assert dbg(0, msg, 256)[1] == B.expand_message_xmd(msg, B.DST, 256)
        for u in B.hash_to_field_fp2(msg):
            r = dbg(2, b2(u), 192)[1]
            assert (f2_from(r[:96]), f2_from(r[96:])) == B.map_to_curve_sswu(u)
        assert dbg(1, msg, 96)[1] == B.g2_compress(B.hash_to_g2(msg))
    for u in [(0, 0), (1, 0), (0, 1)]:
        r = dbg(2, b2(u), 192)[1]
        assert (f2_from(r[:96]), f2_from(r[96:])) == B.map_to_curve_sswu(u)


def test_g2_decompress_subgroup_mul(dbg):
    rnd = random.Random(5)
    Q = B.g2_mul(B.G2_GEN, rnd.randrange(B.R))
    qb = B.g2_compress(Q)
    rc, out = dbg(3, qb, 97)
    assert rc == 0 and out[0] == 1 and out[1:] == qb
    rc, out = dbg(3, B.g2_compress(None), 97)
    assert rc == 1
    x = (3, 1)
    while True:
        y = B.f2_sqrt(B.f2_add(B.f2_mul(B.f2_sqr(x), x), B.B2))
        if y:
            break
        x = (x[0] + 1, 1)
    rc, out = dbg(3, B.g2_compress((x, y)), 97)
    assert rc == 0 and out[0] == 0
    k = rnd.randrange(1 << 64)
    assert dbg(4, qb + k.to_bytes(8, "little"), 96)[1] == B.g2_compress(B.g2_mul(Q, k))


def test_g1_sum(dbg):
    rnd = random.Random(4)
    pts = [B.g1_mul(B.G1_GEN, rnd.randrange(B.R)) for _ in range(5)]
    pts += [pts[0]]
    s = None
    for p in pts:
        s = B.g1_add(s, p)
    assert dbg(7, bytes([len(pts)]) + b"".join(B.g1_uncompressed(p) for p in pts), 96)[1] == B.g1_uncompressed(s)


def test_pairing(dbg):
    Q = B.g2_mul(B.G2_GEN, 777)
    Pt = B.g1_mul(B.G1_GEN, 12345)
    rc, out = dbg(6, B.g1_uncompressed(Pt) + B.g2_compress(Q), 576)
    assert rc == 0
    ref = B.pairing(Pt, Q)
    cube = B.f12_mul(B.f12_sqr(ref), ref)
    (a, b, c), (d, e, g) = cube
    assert out == b"".join(b2(x) for x in (a, b, c, d, e, g))
