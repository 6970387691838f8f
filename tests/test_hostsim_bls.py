"""CPU-only: the device BLS math headers (lighthouse_b200/csrc/bls/*.cuh) compiled for the host with the PTX
carry chains emulated (tests/hostsim), checked limb-exactly against the big-integer oracle.  This validates the
exact algorithms that run on the GPU; the `-m gpu` tests then validate the GPU execution itself."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import pytest

from oracle import bls_ref as B
from tests import oracle_lib as O

HS = os.path.join(O.ROOT, "tests", "hostsim")
P = B.P


@pytest.fixture(scope="module")
def L():
    subprocess.check_call(["make", "-C", HS, "-s"])
    return C.CDLL(os.path.join(HS, "libhostsim.so"))


def be(x): return x.to_bytes(48, "big")
def b2(a): return be(a[0]) + be(a[1])
def f2_from(r): return (int.from_bytes(r[:48], "big"), int.from_bytes(r[48:96], "big"))


def test_fp_ops(L):
    rnd = random.Random(1)
    def op(o, a, b=0):
        out = C.create_string_buffer(48)
        ok = L.hs_fp_op(o, be(a), be(b), out)
        return ok, int.from_bytes(out.raw, "big")
    edge = [0, 1, 2, P - 1, P - 2, (P - 1) // 2, (1 << 380) % P, 2 ** 32 - 1, 2 ** 64 - 1, 2 ** 352]
    vals = edge + [rnd.randrange(P) for _ in range(120)]
    for a in vals:
        for b in vals[:12] + [rnd.randrange(P)]:
            assert op(0, a, b)[1] == a * b % P
            assert op(1, a, b)[1] == (a + b) % P
            assert op(2, a, b)[1] == (a - b) % P
        assert op(4, a)[1] == -a % P
    for a in vals:
        assert op(6, a)[1] == (pow(a, -1, P) if a else 0)     # binary-GCD inversion (final exponentiation)
    for a in vals[:40]:
        if a:
            assert op(3, a)[1] == pow(a, -1, P)
        ok, s = op(5, a)
        assert bool(ok) == (B.fp_sqrt(a) is not None)
        if ok:
            assert s * s % P == a


def test_split_multiplier_matches_montgomery_product(L):
    """fp_redc_inl(fp_mulw_inl(a, b)) (the lazy-reduction building blocks) == a b R^-1 mod p on raw limbs, including
    unreduced operands up to 2^384 - 1 on one side (what Karatsuba's operand sums can look like)."""
    rnd = random.Random(7)
    R = 1 << 384
    Rinv = pow(R, -1, P)

    def limbs(x):
        return (C.c_uint32 * 12)(*[(x >> (32 * i)) & 0xffffffff for i in range(12)])

    cases = [(0, 0), (1, 1), (P - 1, P - 1), (2 * P - 2, 2 * P - 2), (R - 1, 1), (P - 1, 2 * P - 1)]
    cases += [(rnd.randrange(2 * P), rnd.randrange(2 * P)) for _ in range(300)]
    for a, b in cases:
        if a * b >= P * R:
            continue
        out = (C.c_uint32 * 12)()
        L.hs_mont_mul_split(limbs(a), limbs(b), out)
        got = sum(int(out[i]) << (32 * i) for i in range(12))
        assert got == a * b * Rinv % P, (hex(a), hex(b))


def test_dedicated_squaring_matches_product(L):
    """fp_sqr_inl (66 doubled cross products + 12 squares, then the split reduction) == a a R^-1 mod p on raw limbs."""
    rnd = random.Random(11)
    R = 1 << 384
    Rinv = pow(R, -1, P)
    vals = [0, 1, 2, P - 1, P - 2, (1 << 383) % P, ((1 << 384) - 1) % P, 0xffffffff, (1 << 352) - 1]
    vals += [rnd.randrange(P) for _ in range(400)]
    vals += [int("f" * 95, 16) % P, sum(0xffffffff << (64 * i) for i in range(6)) % P]
    for a in vals:
        inp = (C.c_uint32 * 12)(*[(a >> (32 * i)) & 0xffffffff for i in range(12)])
        out = (C.c_uint32 * 12)()
        L.hs_mont_sqr(inp, out)
        got = sum(int(out[i]) << (32 * i) for i in range(12))
        assert got == a * a * Rinv % P, hex(a)


def test_fp2_ops(L):
    rnd = random.Random(2)
    def op(o, a, b=(0, 0)):
        out = C.create_string_buffer(96)
        ok = L.hs_fp2_op(o, b2(a), b2(b), out)
        return ok, f2_from(out.raw)
    vals = [(0, 0), (1, 0), (0, 1), (P - 1, P - 1), (5, 0), (0, 7)] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(40)]
    for a in vals:
        for b in vals[:8]:
            assert op(0, a, b)[1] == B.f2_mul(a, b)
        assert op(1, a)[1] == B.f2_sqr(a)
        if a != (0, 0):
            assert op(2, a)[1] == B.f2_inv(a)
        ok, s = op(3, a)
        assert bool(ok) == (B.f2_sqrt(a) is not None)
        if ok:
            assert B.f2_sqr(s) == a
        sq = B.f2_sqr(a)
        ok, s = op(3, sq)
        assert ok and B.f2_sqr(s) == sq
        assert op(4, a)[0] == B.f2_sgn0(a)


def f12_bytes(f):
    (a, b, c), (d, e, g) = f
    return b"".join(b2(x) for x in (a, b, c, d, e, g))


def f12_from(bs):
    v = [f2_from(bs[96 * i:96 * i + 96]) for i in range(6)]
    return ((v[0], v[1], v[2]), (v[3], v[4], v[5]))


def test_fp12_and_final_exp(L):
    rnd = random.Random(3)
    rf2 = lambda: (rnd.randrange(P), rnd.randrange(P))
    rf12 = lambda: ((rf2(), rf2(), rf2()), (rf2(), rf2(), rf2()))
    def op(o, a, b=None):
        out = C.create_string_buffer(576)
        L.hs_fp12_op(o, f12_bytes(a), f12_bytes(b or a), out)
        return f12_from(out.raw)
    a, b = rf12(), rf12()
    assert op(0, a, b) == B.f12_mul(a, b)
    assert op(1, a) == B.f12_sqr(a)
    assert op(2, a) == B.f12_inv(a)
    assert op(3, a) == B.f12_frob(a)
    assert op(4, a) == B.f12_frob(B.f12_frob(a))
    sp = ((b[0][0], b[0][1], (0, 0)), ((0, 0), b[0][2], (0, 0)))
    assert op(7, a, b) == B.f12_mul(a, sp)
    e = B.f12_mul(B.f12_conj(a), B.f12_inv(a))
    e = B.f12_mul(B.f12_frob(B.f12_frob(e)), e)
    assert op(5, e) == B.f12_sqr(e)                        # Granger-Scott cyclotomic squaring
    ref = B.final_exp(a)
    assert op(6, a) == B.f12_mul(B.f12_sqr(ref), ref)      # device final_exp returns the cube


def words(k, n): return (C.c_uint32 * n)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def test_g1(L):
    rnd = random.Random(4)
    g96 = B.g1_uncompressed(B.G1_GEN)
    for k in O.golden_json("interop_keypairs.json")[:4]:
        sk = int(k["privkey"], 16)
        o96, o48 = C.create_string_buffer(96), C.create_string_buffer(48)
        assert L.hs_g1_mul(g96, words(sk, 8), 255, o96, o48) == 0
        assert o48.raw.hex() == k["pubkey"][2:]
        d96 = C.create_string_buffer(96)
        assert L.hs_g1_decompress(o48.raw, d96) == 0 and d96.raw == o96.raw
    L.hs_g1_mul_u64.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
    Pt = B.g1_mul(B.G1_GEN, 99)
    for r in (1, 2, 3, 4, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000, 0xAAAAAAAAAAAAAAAA, 0x5555555555555555,
              rnd.randrange(1, 1 << 64)):
        o = C.create_string_buffer(96)
        assert L.hs_g1_mul_u64(B.g1_uncompressed(Pt), r, o) == 0
        assert o.raw == B.g1_uncompressed(B.g1_mul(Pt, 2 * r)), hex(r)
    pts = [B.g1_mul(B.G1_GEN, rnd.randrange(B.R)) for _ in range(9)]
    pts += [pts[0], pts[3]]
    s = None
    for p in pts:
        s = B.g1_add(s, p)
    o = C.create_string_buffer(96)
    assert L.hs_g1_sum(b"".join(B.g1_uncompressed(p) for p in pts), len(pts), o) == 0 and o.raw == B.g1_uncompressed(s)
    assert L.hs_g1_sum(B.g1_uncompressed(pts[0]) + B.g1_uncompressed(B.g1_neg(pts[0])), 2, o) == 0 and o.raw[0] == 0x40


def test_g1_subgroup_check_by_endomorphism(L):
    """g1_in_subgroup (phi(P) == -[x^2]P) against the definition [r]P == inf (oracle), on subgroup points, on curve points
    outside the subgroup, and on points of the cofactor's small prime orders (3, 11)"""
    rnd = random.Random(11)
    for _ in range(3):
        assert L.hs_g1_subgroup(B.g1_compress(B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R)))) == 1
    h = (B.X_ABS + 1) ** 2 // 3
    assert h % 3 == 0 and h % 11 == 0
    x, seen = 1, 0
    while seen < 6:
        y = B.fp_sqrt((x ** 3 + 4) % B.P)
        x += 1
        if y is None:
            continue
        pt = (x - 1, y)
        assert not B.g1_in_subgroup(pt)
        assert L.hs_g1_subgroup(B.g1_compress(pt)) == 0
        co = B.g1_mul(pt, B.R)                       # a point of the cofactor group
        if co is not None:
            assert L.hs_g1_subgroup(B.g1_compress(co)) == 0
            for q in (3, 11):
                t = B.g1_mul(co, h // q)             # order q (or infinity)
                if t is not None:
                    assert L.hs_g1_subgroup(B.g1_compress(t)) == 0
            mixed = B.g1_add(co, B.G1_GEN)           # subgroup component + cofactor component
            assert L.hs_g1_subgroup(B.g1_compress(mixed)) == 0
        seen += 1


def test_g2_and_hash_to_curve(L):
    rnd = random.Random(5)
    Q = B.g2_mul(B.G2_GEN, rnd.randrange(B.R))
    qb = B.g2_compress(Q)
    o = C.create_string_buffer(96)
    assert L.hs_g2_roundtrip(qb, o) == 0 and o.raw == qb
    assert L.hs_g2_roundtrip(B.g2_compress(None), o) == 1
    assert L.hs_g2_subgroup(qb) == 1
    x = (3, 1)
    while True:
        y = B.f2_sqrt(B.f2_add(B.f2_mul(B.f2_sqr(x), x), B.B2))
        if y:
            break
        x = (x[0] + 1, 1)
    assert not B.g2_in_subgroup((x, y)) and L.hs_g2_subgroup(B.g2_compress((x, y))) == 0
    k = rnd.randrange(1 << 64)
    assert L.hs_g2_mul(qb, words(k, 2), 64, o) == 0 and o.raw == B.g2_compress(B.g2_mul(Q, k))
    L.hs_g2_mul_r_and_x.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p]
    for r in (1, 2, 3, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000, 0xAAAAAAAAAAAAAAAA, rnd.randrange(1, 1 << 64)):
        o2 = C.create_string_buffer(96)
        assert L.hs_g2_mul_r_and_x(qb, r, o, o2) == 0
        assert o.raw == B.g2_compress(B.g2_mul(Q, r)), hex(r)
        assert o2.raw == B.g2_compress(B.g2_mul(Q, B.X_ABS))
    for msg in (bytes(range(32)), bytes(32), hashlib.sha256(b"x").digest()):
        o256 = C.create_string_buffer(256)
        L.hs_expand(msg, o256)
        assert o256.raw == B.expand_message_xmd(msg, B.DST, 256)
        for u in B.hash_to_field_fp2(msg):
            o192 = C.create_string_buffer(192)
            L.hs_sswu(b2(u), o192)
            assert (f2_from(o192.raw[:96]), f2_from(o192.raw[96:])) == B.map_to_curve_sswu(u)
        L.hs_hash_to_g2(msg, o)
        assert o.raw == B.g2_compress(B.hash_to_g2(msg))
    for u in [(0, 0), (1, 0), (0, 1)]:  # includes the exceptional branch inputs
        o192 = C.create_string_buffer(192)
        L.hs_sswu(b2(u), o192)
        assert (f2_from(o192.raw[:96]), f2_from(o192.raw[96:])) == B.map_to_curve_sswu(u)


def test_pairing_and_verify(L):
    Q = B.g2_mul(B.G2_GEN, 777)
    Pt = B.g1_mul(B.G1_GEN, 12345)
    for proj in (0, 1):
        out = C.create_string_buffer(576)
        assert L.hs_pairing(B.g1_uncompressed(Pt), B.g2_compress(Q), 1, proj, out) == 0
        ref = B.pairing(B.g1_add(Pt, Pt) if proj else Pt, Q)
        assert f12_from(out.raw) == B.f12_mul(B.f12_sqr(ref), ref)
    # multi-pairing Miller loop (shared squarings, Jacobian Q with Z != 1) == product of the single affine loops after
    # the final exponentiation, limb for limb
    for m in (1, 2, 3, 4):
        ps = b"".join(B.g1_uncompressed(B.g1_mul(B.G1_GEN, 1000 + 7 * j)) for j in range(m))
        qs = b"".join(B.g2_compress(B.g2_mul(B.G2_GEN, 55 + 3 * j)) for j in range(m))
        a, b = C.create_string_buffer(576), C.create_string_buffer(576)
        assert L.hs_miller_multi(ps, qs, m, a, b) == 0
        assert a.raw == b.raw
    d = O.golden_json("deposit_data.json")[0]
    fv = bytes.fromhex(d["fork_version"])
    fdr = hashlib.sha256(fv + bytes(28) + bytes(32)).digest()
    m = hashlib.sha256(bytes.fromhex(d["deposit_message_root"]) + bytes([3, 0, 0, 0]) + fdr[:28]).digest()
    pk = B.g1_uncompressed(B.g1_decompress(bytes.fromhex(d["pubkey"])))
    sig = bytes.fromhex(d["signature"])
    assert L.hs_verify(pk, m, sig) == 1
    assert L.hs_verify(pk, bytes(32), sig) == 0


def test_fused_sum_of_products_matches_big_integers(L):
    """bls/sop.cuh: fp_sop2<K> / fp_sop1<K> == (sum_q x_q y_q) R^-1 mod p on raw limbs, at the operand bounds the call
    sites use (x < X p with K X <= 8; y < p), streaming y from a strided array."""
    rnd = random.Random(21)
    R = 1 << 384
    Rinv = pow(R, -1, P)

    def arr(vals):
        return (C.c_uint32 * (12 * len(vals)))(*[(v >> (32 * i)) & 0xffffffff for v in vals for i in range(12)])

    for k, xmax in ((1, 2 * P), (2, P + 1), (2, 2 * P), (4, 2 * P), (6, P + 1), (8, P + 1)):
        for trial in range(40):
            if trial == 0:
                xa = [xmax - 1] * k; ya = [P - 1] * k; xb = [xmax - 1] * k; yb = [P - 1] * k
            elif trial == 1:
                xa = [0] * k; ya = [0] * k; xb = [1] * k; yb = [1] * k
            else:
                xa = [rnd.randrange(xmax) for _ in range(k)]; ya = [rnd.randrange(P) for _ in range(k)]
                xb = [rnd.randrange(xmax) for _ in range(k)]; yb = [rnd.randrange(P) for _ in range(k)]
            oa, ob = (C.c_uint32 * 12)(), (C.c_uint32 * 12)()
            assert L.hs_sop2(k, arr(xa), arr(ya), arr(xb), arr(yb), oa, ob) == 0
            ga = sum(int(oa[i]) << (32 * i) for i in range(12))
            gb = sum(int(ob[i]) << (32 * i) for i in range(12))
            assert ga == sum(x * y for x, y in zip(xa, ya)) * Rinv % P, (k, trial)
            assert gb == sum(x * y for x, y in zip(xb, yb)) * Rinv % P, (k, trial)


def test_cooperative_miller_program_matches_oracle(L):
    """bls/miller_coop.cuh run lane by lane (two groups of six per block): product of pairings, after the final
    exponentiation, == cube of the oracle's GT value — for ragged set counts, several rounds per lane, skipped sets, an
    infinite Q and the appended (-g1, Q_extra) pair."""
    def pts(n, seed):
        ps = [B.g1_mul(B.G1_GEN, 1000 + 7 * j + seed) for j in range(n)]
        qs = [B.g2_mul(B.G2_GEN, 55 + 3 * j + seed) for j in range(n)]
        return ps, qs

    def ref_cube(pairs):
        f = None
        for p, q in pairs:
            if q is None:
                continue
            m = B.miller_loop(p, q)
            f = m if f is None else B.f12_mul(f, m)
        g = B.final_exp(f)
        return B.f12_mul(B.f12_sqr(g), g)

    for n, spb, skip, with_extra, inf_at in ((1, 12, (), False, None), (7, 12, (), False, None), (13, 12, (3,), True, None),
                                             (29, 36, (0, 17), True, 5), (6, 12, (), True, None)):
        ps, qs = pts(n, n)
        if inf_at is not None:
            qs[inf_at] = None
        status = bytes(1 if j in skip else 0 for j in range(n))
        extra = B.g2_mul(B.G2_GEN, 424242) if with_extra else None
        out = C.create_string_buffer(576)
        rc = L.hs_miller_coop(b"".join(B.g1_uncompressed(p) for p in ps), b"".join(B.g2_compress(q) for q in qs), status, n,
                              B.g2_compress(extra) if with_extra else None, spb, out)
        assert rc == 0
        pairs = [(B.g1_add(p, p), q) for j, (p, q) in enumerate(zip(ps, qs)) if j not in skip]
        if with_extra:
            pairs.append((B.g1_neg(B.G1_GEN), extra))
        assert f12_from(out.raw) == ref_cube(pairs), (n, spb)


def test_warp_per_pairing_miller_program_matches_oracle(L):
    """bls/miller_warp.cuh (phase tables from scripts/gen_miller_warp.py) run lane by lane: product of pairings, after the
    final exponentiation, == cube of the oracle's GT value — with skipped sets, an infinite Q, the appended (-g1, Q_extra)
    pair and blocks of 1, 3 and 4 warps (the dense product section)."""
    def pts(n, seed):
        ps = [B.g1_mul(B.G1_GEN, 2000 + 11 * j + seed) for j in range(n)]
        qs = [B.g2_mul(B.G2_GEN, 77 + 5 * j + seed) for j in range(n)]
        return ps, qs

    def ref_cube(pairs):
        f = B.F12_ONE
        for p, q in pairs:
            if q is not None:
                f = B.f12_mul(f, B.miller_loop(p, q))
        g = B.final_exp(f)
        return B.f12_mul(B.f12_sqr(g), g)

    L.hs_miller_warp.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p]
    for n, wpb, skip, with_extra, inf_at in ((1, 1, (), False, None), (3, 4, (), True, None), (5, 3, (1,), True, 3)):
        ps, qs = pts(n, n)
        if inf_at is not None:
            qs[inf_at] = None
        status = bytes(1 if j in skip else 0 for j in range(n))
        extra = B.g2_mul(B.G2_GEN, 31337) if with_extra else None
        out = C.create_string_buffer(576)
        rc = L.hs_miller_warp(b"".join(B.g1_uncompressed(p) for p in ps), b"".join(B.g2_compress(q) for q in qs), status, n,
                              B.g2_compress(extra) if with_extra else None, wpb, out)
        assert rc == 0
        pairs = [(B.g1_add(p, p), q) for j, (p, q) in enumerate(zip(ps, qs)) if j not in skip]
        if with_extra:
            pairs.append((B.g1_neg(B.G1_GEN), extra))
        assert f12_from(out.raw) == ref_cube(pairs), (n, wpb)


def test_g2_warp_programs_match_oracle(L):
    """bls/g2_warp.cuh lane by lane: [r] sig and the psi-based subgroup test for a G2 point and for a curve point outside
    G2, and clear_cofactor == [h_eff] P (RFC 9380 G.3) on a point outside G2."""
    L.hs_sig_warp.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
    L.hs_clear_cofactor_warp.argtypes = [C.c_char_p, C.c_char_p]
    rnd = random.Random(99)
    o = C.create_string_buffer(96)
    for r in (1, 2, 3, (1 << 64) - 1, rnd.randrange(1, 1 << 64)):
        sig = B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R))
        assert L.hs_sig_warp(B.g2_compress(sig), r, o) == 1 and o.raw == B.g2_compress(B.g2_mul(sig, r)), r
    x = (7, 1)
    while True:
        y = B.f2_sqrt(B.f2_add(B.f2_mul(B.f2_sqr(x), x), B.B2))
        if y and not B.g2_in_subgroup((x, y)):
            break
        x = (x[0] + 1, 1)
    assert L.hs_sig_warp(B.g2_compress((x, y)), 12345, o) == 0
    # points of order 13 and 23 (the cofactor of G2 is 13^2 23^2 2713 ...): the additions of both ladders degenerate
    # (T == +-Q), Z becomes 0 and stays 0, and the test must reject them — for every scalar
    xx = -B.X_ABS
    h2 = (xx ** 8 - 4 * xx ** 7 + 5 * xx ** 6 - 4 * xx ** 4 + 6 * xx ** 3 - 4 * xx ** 2 - 4 * xx + 13) // 9
    assert h2 % (13 * 13 * 23 * 23) == 0
    co = B.g2_mul((x, y), B.R)
    for q in (13, 23):
        t = B.g2_mul(co, h2 // (q * q))
        if t is not None and B.g2_mul(t, q) is not None:
            t = B.g2_mul(t, q)
        assert t is not None and B.g2_mul(t, q) is None
        for r in (1, 5, q, q + 1, (1 << 64) - 1):
            assert L.hs_sig_warp(B.g2_compress(t), r, o) == 0, (q, r)
    assert L.hs_clear_cofactor_warp(B.g2_compress((x, y)), o) == 0 and o.raw == B.g2_compress(B.g2_mul((x, y), B.H_EFF))


def test_final_exponentiation_warp_program_matches_single_thread(L):
    """bls/fe_warp.cuh lane by lane: product of two Miller values, then f^(3 (p^12 - 1) / r) — equal to the single-thread
    final_exp of pairing.cuh (hs_pairing) and to the oracle's GT value cubed."""
    L.hs_final_warp.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
    p1, q1 = B.g1_mul(B.G1_GEN, 4242), B.g2_mul(B.G2_GEN, 777)
    p2, q2 = B.g1_mul(B.G1_GEN, 99), B.g2_mul(B.G2_GEN, 31)
    m1, m2, o = C.create_string_buffer(576), C.create_string_buffer(576), C.create_string_buffer(576)
    assert L.hs_pairing(B.g1_uncompressed(p1), B.g2_compress(q1), 0, 0, m1) == 0      # Miller values, no final exp
    assert L.hs_pairing(B.g1_uncompressed(p2), B.g2_compress(q2), 0, 0, m2) == 0
    assert L.hs_final_warp(m1.raw, m2.raw, o) == 0
    g = B.final_exp(B.f12_mul(B.miller_loop(p1, q1), B.miller_loop(p2, q2)))
    assert f12_from(o.raw) == B.f12_mul(B.f12_sqr(g), g)


def test_g2_sum_warp_program_matches_oracle(L):
    """The point-sum step of k_g2_sum_warp (sum r_i sig_i tree in latency mode) lane by lane: sums with infinity entries
    skipped == the oracle's g2_add chain; nothing but infinities gives the infinity encoding."""
    L.hs_g2_sum_warp.argtypes = [C.c_char_p, C.c_int, C.c_char_p]
    rnd = random.Random(123)
    o = C.create_string_buffer(96)
    pts = [B.g2_mul(B.G2_GEN, rnd.randrange(1, B.R)) for _ in range(5)]
    for pick in ([0], [0, 1], [None, 0, 1, 2, None, 3, 4], [None, None]):
        enc = b"".join(B.g2_compress(None if k is None else pts[k]) for k in pick)
        assert L.hs_g2_sum_warp(enc, len(pick), o) == 0
        acc = None
        for k in pick:
            if k is not None:
                acc = B.g2_add(acc, pts[k])
        assert o.raw == B.g2_compress(acc), pick
