"""oracle/bls_ref.py — TEST INFRASTRUCTURE ONLY: from-spec big-integer restatement of the BLS12-381 path.

Restates, with Python integers, what Lighthouse's `crypto/bls` obtains from the un-vendored crate
blst 0.3.12 (Cargo.lock:1023): min-pk BLS signatures over BLS12-381 with the
BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_ ciphersuite (IETF BLS draft + RFC 9380 §8.8.2), anchored on the
reference's call sites:
    verify_signature_sets ........... crypto/bls/src/impls/blst.rs:37-119   (semantics: SURVEY Appendix C)
    fast_aggregate_verify ........... crypto/bls/src/impls/blst.rs:250-261, generic_aggregate_signature.rs:187-210
    (de)serialisation ............... crypto/bls/src/generic_public_key.rs:12-21,86-94, generic_signature.rs:15-26
    interop secret keys ............. common/eth2_interop_keypairs/src/lib.rs:40-56
Pinned by tests/test_oracle_bls.py against the reference's in-tree vectors: 22 deposit (pk, msg, sig) triples
(validator_manager/test_vectors), 10 interop sk->pk pairs, and the derivable cases of crypto/bls/tests/tests.rs.

Deliberately simple (affine formulas, generic Fp12 multiplication, plain-exponent final exponentiation) so it is
independent of the optimised formulas used in the C oracle (oracle/bls12_381.c) and the CUDA kernels.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import hashlib

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000  # |x|, x negative
DST = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"  # crypto/bls/src/impls/blst.rs:15
H_EFF = 0xBC69F08F2EE75B3584C6A0EA91B352888E2A8E9145AD7689986FF031508FFE1329C2F178731DB956D82BF015D1212B02EC0EC69D7477C1AE954CBC06689F6A359894C0ADEBBF6B4E8020005AAA95551

G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)

# ------------------------------------------------------------------------------------------------ Fp2 = Fp[i]/(i^2+1)
F2_ZERO, F2_ONE = (0, 0), (1, 0)


def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return (-a[0] % P, -a[1] % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_muls(a, s): return (a[0] * s % P, a[1] * s % P)
def f2_conj(a): return (a[0], -a[1] % P)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, -a[1] * n % P)


def f2_pow(a, e):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def fp_sqrt(a):
    """sqrt in Fp (p = 3 mod 4) or None"""
    a %= P
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a else None


def f2_sqrt(a):
    """Some square root of a in Fp2, or None (norm method)."""
    a0, a1 = a
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt(-a0 % P)  # a0 non-residue in Fp => sqrt is purely imaginary
        return (0, s)
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    inv2 = (P + 1) // 2
    for nn in (n, -n % P):
        x0 = fp_sqrt((a0 + nn) * inv2 % P)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            r = (x0, x1)
            assert f2_sqr(r) == (a0 % P, a1 % P)
            return r
    return None


def f2_sgn0(a):
    """RFC 9380 sgn0 for m = 2"""
    s0, z0, s1 = a[0] & 1, a[0] == 0, a[1] & 1
    return s0 | (z0 & s1)


XI = (1, 1)  # 1 + i

# ---------------------------------------------------------------- Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v)
F6_ZERO, F6_ONE = (F2_ZERO, F2_ZERO, F2_ZERO), (F2_ONE, F2_ZERO, F2_ZERO)


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul(XI, f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul(XI, f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):  # multiply by v
    return (f2_mul(XI, a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul(XI, f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul(XI, f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul(XI, f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    return (f6_add(f6_mul(a0, b0), f6_mul_v(f6_mul(a1, b1))), f6_add(f6_mul(a0, b1), f6_mul(a1, b0)))


def f12_sqr(a): return f12_mul(a, a)
def f12_conj(a): return (a[0], f6_neg(a[1]))  # = a^(p^6)


def f12_inv(a):
    a0, a1 = a
    d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, d), f6_neg(f6_mul(a1, d)))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_sqr(a)
        e >>= 1
    return r


# Frobenius on Fp12: coefficients gamma_k = xi^(k (p-1)/6)
_GAMMA = [f2_pow(XI, k * (P - 1) // 6) for k in range(6)]


def f12_frob(a):
    """a^p"""
    (a00, a01, a02), (a10, a11, a12) = a
    # element = sum_{k} c_k w^k with w-degrees: a00:0 a01:2 a02:4 a10:1 a11:3 a12:5
    c = [a00, a10, a01, a11, a02, a12]
    c = [f2_mul(f2_conj(ck), _GAMMA[k]) for k, ck in enumerate(c)]
    return ((c[0], c[2], c[4]), (c[1], c[3], c[5]))


# ------------------------------------------------------------------------------------------------ curves (affine)
# Points: None = infinity, else (x, y).  G1 over Fp ints (b = 4), G2 over Fp2 tuples (b = 4 xi).
class _FpOps:
    zero, one = 0, 1
    @staticmethod
    def add(a, b): return (a + b) % P
    @staticmethod
    def sub(a, b): return (a - b) % P
    @staticmethod
    def mul(a, b): return a * b % P
    @staticmethod
    def neg(a): return -a % P
    @staticmethod
    def inv(a): return pow(a, -1, P)
    @staticmethod
    def muls(a, s): return a * s % P


class _Fp2Ops:
    zero, one = F2_ZERO, F2_ONE
    add, sub, mul, neg, inv, muls = map(staticmethod, (f2_add, f2_sub, f2_mul, f2_neg, f2_inv, f2_muls))


def _ec_add(F, p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        lam = F.mul(F.muls(F.mul(x1, x1), 3), F.inv(F.muls(y1, 2)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    return (x3, F.sub(F.mul(lam, F.sub(x1, x3)), y1))


def _ec_neg(F, p): return None if p is None else (p[0], F.neg(p[1]))


def _ec_mul(F, p, k):
    if k < 0:
        return _ec_mul(F, _ec_neg(F, p), -k)
    r = None
    while k:
        if k & 1:
            r = _ec_add(F, r, p)
        p = _ec_add(F, p, p)
        k >>= 1
    return r


def g1_add(a, b): return _ec_add(_FpOps, a, b)
def g1_neg(a): return _ec_neg(_FpOps, a)
def g1_mul(a, k): return _ec_mul(_FpOps, a, k)
def g2_add(a, b): return _ec_add(_Fp2Ops, a, b)
def g2_neg(a): return _ec_neg(_Fp2Ops, a)
def g2_mul(a, k): return _ec_mul(_Fp2Ops, a, k)

B1 = 4
B2 = (4, 4)


def g1_on_curve(p): return p is None or (p[1] * p[1] - p[0] ** 3 - B1) % P == 0
def g2_on_curve(p): return p is None or f2_sub(f2_sqr(p[1]), f2_add(f2_mul(f2_sqr(p[0]), p[0]), B2)) == F2_ZERO
def g1_in_subgroup(p): return g1_mul(p, R) is None
def g2_in_subgroup(p): return g2_mul(p, R) is None


# psi endomorphism — used only to cross-check the fast forms used on the device
PSI_CX = f2_inv(f2_pow(XI, (P - 1) // 3))
PSI_CY = f2_inv(f2_pow(XI, (P - 1) // 2))


def g2_psi(p):
    if p is None:
        return None
    return (f2_mul(f2_conj(p[0]), PSI_CX), f2_mul(f2_conj(p[1]), PSI_CY))


# ------------------------------------------------------------------------------------------------ serialisation
def g1_compress(p):
    if p is None:
        return bytes([0xC0]) + bytes(47)
    x, y = p
    flag = 0x80 | (0x20 if y > (P - 1) // 2 else 0)
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= flag
    return bytes(b)


def g1_uncompressed(p):
    """96-byte x || y big-endian (blst serialize_uncompressed; validator_pubkey_cache.rs:195-199)"""
    if p is None:
        return bytes([0x40]) + bytes(95)
    return p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")


def g1_decompress(b):
    """-> point or None(infinity); raises ValueError on bad encodings.  No subgroup check."""
    if len(b) != 48:
        raise ValueError("length")
    c, inf, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c:
        raise ValueError("uncompressed flag in 48-byte encoding")
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if inf:
        if x != 0 or s:
            raise ValueError("bad infinity encoding")
        return None
    if x >= P:
        raise ValueError("x >= p")
    y = fp_sqrt((x * x * x + B1) % P)
    if y is None:
        raise ValueError("not on curve")
    if (y > (P - 1) // 2) != bool(s):
        y = P - y
    return (x, y)


def g2_compress(p):
    if p is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), (y0, y1) = p
    big = (y1 > (P - 1) // 2) if y1 != 0 else (y0 > (P - 1) // 2)
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if big else 0)
    return bytes(b)


def g2_decompress(b):
    if len(b) != 96:
        raise ValueError("length")
    c, inf, s = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    if not c:
        raise ValueError("uncompressed flag in 96-byte encoding")
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    x0 = int.from_bytes(b[48:], "big")
    if inf:
        if x0 or x1 or s:
            raise ValueError("bad infinity encoding")
        return None
    if x0 >= P or x1 >= P:
        raise ValueError("coordinate >= p")
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
    if y is None:
        raise ValueError("not on curve")
    big = (y[1] > (P - 1) // 2) if y[1] != 0 else (y[0] > (P - 1) // 2)
    if big != bool(s):
        y = f2_neg(y)
    return (x, y)


# ------------------------------------------------------------------------------------------------ hash to G2 (RFC 9380)
def expand_message_xmd(msg, dst, n):
    ell = (n + 31) // 32
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\0" + dst_prime).digest()
    bs = [hashlib.sha256(b0 + b"\x01" + dst_prime).digest()]
    for i in range(2, ell + 1):
        bs.append(hashlib.sha256(bytes(x ^ y for x, y in zip(b0, bs[-1])) + bytes([i]) + dst_prime).digest())
    return b"".join(bs)[:n]


def hash_to_field_fp2(msg, count=2, dst=DST):
    u = expand_message_xmd(msg, dst, count * 2 * 64)
    out = []
    for i in range(count):
        e = [int.from_bytes(u[64 * (2 * i + j): 64 * (2 * i + j + 1)], "big") % P for j in range(2)]
        out.append((e[0], e[1]))
    return out


SSWU_A = (0, 240)
SSWU_B = (1012, 1012)
SSWU_Z = (P - 2, P - 1)  # -(2 + i)


def sswu_g(x):
    return f2_add(f2_add(f2_mul(f2_sqr(x), x), f2_mul(SSWU_A, x)), SSWU_B)


def map_to_curve_sswu(u):
    """simplified SWU onto E2': y^2 = x^3 + A'x + B' (RFC 9380 §6.6.2, straight-line spec form)"""
    zu2 = f2_mul(SSWU_Z, f2_sqr(u))
    tv1 = f2_add(f2_sqr(zu2), zu2)
    if tv1 == F2_ZERO:
        x1 = f2_mul(SSWU_B, f2_inv(f2_mul(SSWU_Z, SSWU_A)))
    else:
        x1 = f2_mul(f2_mul(f2_neg(SSWU_B), f2_inv(SSWU_A)), f2_add(F2_ONE, f2_inv(tv1)))
    gx1 = sswu_g(x1)
    y = f2_sqrt(gx1)
    if y is not None:
        x = x1
    else:
        x = f2_mul(zu2, x1)
        y = f2_sqrt(sswu_g(x))
        assert y is not None
    if f2_sgn0(u) != f2_sgn0(y):
        y = f2_neg(y)
    return (x, y)


def _c(c0, c1=None): return (c0, c0 if c1 is None else c1)


ISO_XNUM = [
    _c(0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97D6),
    (0, 0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71A),
    (0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71E,
     0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38D),
    (0x171D6541FA38CCFAED6DEA691F5FB614CB14B4E7F4E810AA22D6108F142B85757098E38D0F671C7188E2AAAAAAAA5ED1, 0),
]
ISO_XDEN = [(0, P - 72), (12, P - 12), (1, 0)]
ISO_YNUM = [
    _c(0x1530477C7AB4113B59A4C18B076D11930F7DA5D4A07F649BF54439D87D27E500FC8C25EBF8C92F6812CFC71C71C6D706),
    (0, 0x5C759507E8E333EBB5B7A9A47D7ED8532C52D39FD3A042A88B58423C50AE15D5C2638E343D9C71C6238AAAAAAAA97BE),
    (0x11560BF17BAA99BC32126FCED787C88F984F87ADF7AE0C7F9A208C6B4F20A4181472AAA9CB8D555526A9FFFFFFFFC71C,
     0x8AB05F8BDD54CDE190937E76BC3E447CC27C3D6FBD7063FCD104635A790520C0A395554E5C6AAAA9354FFFFFFFFE38F),
    (0x124C9AD43B6CF79BFBF7043DE3811AD0761B0F37A1E26286B0E977C69AA274524E79097A56DC4BD9E1B371C71C718B10, 0),
]
ISO_YDEN = [(P - 432, P - 432), (0, P - 216), (18, P - 18), (1, 0)]


def _horner(coeffs, x):
    r = F2_ZERO
    for c in reversed(coeffs):
        r = f2_add(f2_mul(r, x), c)
    return r


def iso_map_g2(pt):
    x, y = pt
    xn, xd, yn, yd = (_horner(c, x) for c in (ISO_XNUM, ISO_XDEN, ISO_YNUM, ISO_YDEN))
    if xd == F2_ZERO or yd == F2_ZERO:
        return None
    return (f2_mul(xn, f2_inv(xd)), f2_mul(y, f2_mul(yn, f2_inv(yd))))


def map_to_g2_uncleared(msg, dst=DST):
    u0, u1 = hash_to_field_fp2(msg, 2, dst)
    return g2_add(iso_map_g2(map_to_curve_sswu(u0)), iso_map_g2(map_to_curve_sswu(u1)))


def hash_to_g2(msg, dst=DST):
    return g2_mul(map_to_g2_uncleared(msg, dst), H_EFF)


# ------------------------------------------------------------------------------------------------ pairing
def _embed_line(c_const, c_v, c_vw):
    """c_const + c_v * v + c_vw * (v w)  as an Fp12 element"""
    return ((c_const, c_v, F2_ZERO), (F2_ZERO, c_vw, F2_ZERO))


def _line(T, Q, Pt):
    """Line through T and Q (tangent if equal) on the twist, evaluated at the G1 point Pt, times w^3:
       yP w^3 - lam xP w^2 + (lam xT - yT).  Returns (fp12 line, T + Q)."""
    xP, yP = Pt
    (xT, yT), (xQ, yQ) = T, Q
    if T == Q:
        lam = f2_mul(f2_muls(f2_sqr(xT), 3), f2_inv(f2_muls(yT, 2)))
    else:
        lam = f2_mul(f2_sub(yQ, yT), f2_inv(f2_sub(xQ, xT)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), xT), xQ)
    y3 = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
    l = _embed_line(f2_sub(f2_mul(lam, xT), yT), f2_neg(f2_muls(lam, xP)), (yP, 0))
    return l, (x3, y3)


def miller_loop(Pt, Q):
    """f_{|x|,Q}(P) conjugated (x < 0); Pt in G1 affine, Q in G2 affine; either at infinity -> 1."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for i in range(X_ABS.bit_length() - 2, -1, -1):
        l, T = _line(T, T, Pt)
        f = f12_mul(f12_sqr(f), l)
        if (X_ABS >> i) & 1:
            l, T = _line(T, Q, Pt)
            f = f12_mul(f, l)
    return f12_conj(f)


def final_exp(f):
    """f^((p^12-1)/r), exponent applied literally (easy part via conj/inverse/frobenius)."""
    f = f12_mul(f12_conj(f), f12_inv(f))          # ^(p^6 - 1)
    f = f12_mul(f12_frob(f12_frob(f)), f)         # ^(p^2 + 1)
    return f12_pow(f, (P ** 4 - P ** 2 + 1) // R)


def pairing(Pt, Q): return final_exp(miller_loop(Pt, Q))


def multi_pairing_is_one(pairs):
    f = F12_ONE
    for Pt, Q in pairs:
        f = f12_mul(f, miller_loop(Pt, Q))
    return final_exp(f) == F12_ONE


# ------------------------------------------------------------------------------------------------ BLS scheme (min-pk)
def sk_to_pk(sk): return g1_mul(G1_GEN, sk % R)
def sign(sk, msg): return g2_mul(hash_to_g2(msg), sk % R)


def interop_secret_key(i):
    """common/eth2_interop_keypairs/src/lib.rs:40-56: int_le(SHA256(le32-padded index)) mod r"""
    pre = i.to_bytes(32, "little")
    return int.from_bytes(hashlib.sha256(pre).digest(), "little") % R


def core_verify(pk, msg, sig):
    """e(pk, H(m)) == e(g1, sig)  with pk in G1, sig in G2 (already subgroup-checked by the caller)."""
    if pk is None:
        return False
    return multi_pairing_is_one([(pk, hash_to_g2(msg)), (g1_neg(G1_GEN), sig)])


EMPTY_SIG = bytes(96)  # Lighthouse's "empty" signature (generic_signature.rs:26)


def verify_signature_sets(sets, rands):
    """sets: list of (sig_bytes96, [pubkey points], msg32); rands: nonzero 64-bit ints.
    Semantics of crypto/bls/src/impls/blst.rs:37-119."""
    if not sets:
        return False
    pairs = []
    acc = None
    for (sig_b, pks, msg), r in zip(sets, rands):
        if sig_b == EMPTY_SIG:
            return False
        try:
            sig = g2_decompress(sig_b)
        except ValueError:
            return False
        if not g2_in_subgroup(sig):
            return False
        if not pks:
            return False
        apk = None
        for pk in pks:
            apk = g1_add(apk, pk)
        if apk is None:
            return False
        assert 0 < r < (1 << 64)
        pairs.append((g1_mul(apk, r), hash_to_g2(msg)))
        acc = g2_add(acc, g2_mul(sig, r))
    pairs.append((g1_neg(G1_GEN), acc))
    return multi_pairing_is_one(pairs)
