"""Pins the big-integer BLS oracle (oracle/bls_ref.py) against the reference's in-tree golden vectors
(SURVEY.md §8c).  CPU only."""
import hashlib

from oracle import bls_ref as B
from tests import oracle_lib as O


def deposit_signing_root(d):
    fv = bytes.fromhex(d["fork_version"])
    fdr = hashlib.sha256(fv + bytes(28) + bytes(32)).digest()
    return hashlib.sha256(bytes.fromhex(d["deposit_message_root"]) + bytes([3, 0, 0, 0]) + fdr[:28]).digest()


def test_curve_constants():
    assert B.g1_on_curve(B.G1_GEN) and B.g2_on_curve(B.G2_GEN)
    assert B.g1_in_subgroup(B.G1_GEN) and B.g2_in_subgroup(B.G2_GEN)
    x = -B.X_ABS
    assert B.R == x ** 4 - x ** 2 + 1 and B.P == (x - 1) ** 2 * B.R // 3 + x
    Q = B.g2_mul(B.G2_GEN, 12345)
    assert B.g2_psi(Q) == B.g2_neg(B.g2_mul(Q, B.X_ABS))      # psi(P) == [x]P on G2
    # hard-part identity used by the device final exponentiation
    assert (x - 1) ** 2 * (x + B.P) * (x * x + B.P * B.P - 1) + 3 == 3 * ((B.P ** 4 - B.P ** 2 + 1) // B.R)


def test_interop_keypairs_golden():
    """common/eth2_interop_keypairs/specs/keygen_10_validators.yaml (tests/generation.rs:6-64)"""
    for i, k in enumerate(O.golden_json("interop_keypairs.json")):
        sk = int(k["privkey"], 16)
        assert sk == B.interop_secret_key(i)
        assert B.g1_compress(B.sk_to_pk(sk)).hex() == k["pubkey"][2:]
        assert B.g1_decompress(bytes.fromhex(k["pubkey"][2:])) == B.sk_to_pk(sk)


def test_deposit_signatures_golden():
    """22 triples that the reference asserts valid (validator_manager/src/create_validators.rs:749-752)."""
    deps = O.golden_json("deposit_data.json")
    assert len(deps) == 22
    for i, d in enumerate(deps):
        pk = B.g1_decompress(bytes.fromhex(d["pubkey"]))
        sig = B.g2_decompress(bytes.fromhex(d["signature"]))
        assert B.g1_in_subgroup(pk) and B.g2_in_subgroup(sig)
        assert B.g2_compress(sig).hex() == d["signature"]
        m = deposit_signing_root(d)
        assert B.core_verify(pk, m, sig)
        if i < 2:
            assert not B.core_verify(pk, bytes(32), sig)


def test_bilinearity_and_batch_semantics():
    e1 = B.pairing(B.g1_mul(B.G1_GEN, 5), B.g2_mul(B.G2_GEN, 7))
    assert e1 == B.f12_pow(B.pairing(B.G1_GEN, B.G2_GEN), 35)
    msg = hashlib.sha256(b"a").digest()
    pk, sig = B.sk_to_pk(3), B.g2_compress(B.sign(3, msg))
    assert B.verify_signature_sets([(sig, [pk], msg)], [5])
    assert not B.verify_signature_sets([], [])
    assert not B.verify_signature_sets([(B.EMPTY_SIG, [pk], msg)], [5])
    assert not B.verify_signature_sets([(sig, [], msg)], [5])
    assert not B.verify_signature_sets([(B.g2_compress(None), [pk], msg)], [5])
    assert not B.verify_signature_sets([(sig, [pk, B.g1_neg(pk)], msg)], [5])


# ---- C oracle (oracle/bls12_381.c) vs the big-integer oracle
def test_c_oracle_stages_match_python():
    for i in range(3):
        msg = hashlib.sha256(b"c%d" % i).digest()
        assert O.bls_hash_to_g2(msg) == B.g2_compress(B.hash_to_g2(msg))
        sk = B.interop_secret_key(i)
        assert O.bls_sk_to_pk(sk.to_bytes(32, "big")) == B.g1_uncompressed(B.sk_to_pk(sk))
    assert O.bls_sign((12345).to_bytes(32, "big"), bytes(32)) == B.g2_compress(B.sign(12345, bytes(32)))


def _mk_sets(n, k):
    import numpy as np
    sks = [[B.interop_secret_key(7 * i + j) for j in range(k)] for i in range(n)]
    msgs = [hashlib.sha256(b"set%d" % i).digest() for i in range(n)]
    pks = b"".join(O.bls_sk_to_pk(s.to_bytes(32, "big")) for row in sks for s in row)
    sigs = b"".join(O.bls_sign((sum(row) % B.R).to_bytes(32, "big"), m) for row, m in zip(sks, msgs))
    offs = np.arange(n + 1, dtype=np.uint32) * k
    return sigs, b"".join(msgs), pks, offs


def test_c_oracle_batch_verdicts_and_gt():
    import numpy as np
    sigs, msgs, pks, offs = _mk_sets(5, 3)
    rands = [3, 2 ** 64 - 1, 0x123456789, 7, 0x8000000000000001]
    for threads in (1, 3, 8):
        O.set_threads(threads)
        ok, gt = O.bls_verify_signature_sets(sigs, msgs, pks, offs, rands, want_gt=True)
        assert ok
    O.set_threads(1)
    # python oracle agrees on the same batch
    pts = [[(int.from_bytes(pks[96 * (3 * i + j):96 * (3 * i + j) + 48], "big"),
             int.from_bytes(pks[96 * (3 * i + j) + 48:96 * (3 * i + j) + 96], "big")) for j in range(3)] for i in range(5)]
    sets = [(sigs[96 * i:96 * i + 96], pts[i], msgs[32 * i:32 * i + 32]) for i in range(5)]
    assert B.verify_signature_sets(sets[:2], rands[:2])
    # tampered message: verdict false and the GT value equals the python oracle's cube, limb for limb
    bad = bytearray(msgs); bad[0] ^= 1
    ok, gt = O.bls_verify_signature_sets(sigs[:192], bytes(bad[:64]), pks[:576], offs[:3], rands[:2], want_gt=True)
    assert not ok
    f, acc = B.F12_ONE, None
    for (sb, ps, _), m, r in zip(sets[:2], (bytes(bad[:32]), bytes(bad[32:64])), rands[:2]):
        apk = None
        for p in ps:
            apk = B.g1_add(apk, p)
        f = B.f12_mul(f, B.miller_loop(B.g1_mul(apk, r), B.hash_to_g2(m)))
        acc = B.g2_add(acc, B.g2_mul(B.g2_decompress(sb), r))
    f = B.f12_mul(f, B.miller_loop(B.g1_neg(B.G1_GEN), acc))
    g = B.final_exp(f)
    cube = B.f12_mul(B.f12_sqr(g), g)
    (a, b, c), (d, e, h) = cube
    assert gt == b"".join(x[0].to_bytes(48, "big") + x[1].to_bytes(48, "big") for x in (a, b, c, d, e, h))
    # semantics (Appendix C): empty sig, no keys, infinity sig, apk at infinity, empty batch
    ok, st = O.bls_verify_signature_sets(bytes(96) + sigs[96:192], msgs[:64], pks[:576], offs[:3], rands[:2], want_status=True)
    assert not ok and list(st) == [1, 0]
    ok, st = O.bls_verify_signature_sets(sigs[:96], msgs[:32], b"", np.array([0, 0]), rands[:1], want_status=True)
    assert not ok and list(st) == [4]
    assert not O.bls_verify_signature_sets(B.g2_compress(None), msgs[:32], pks[:288], offs[:2], rands[:1])
    neg = B.g1_uncompressed(B.g1_neg(pts[0][0]))
    ok, st = O.bls_verify_signature_sets(sigs[:96], msgs[:32], pks[:96] + neg, np.array([0, 2]), rands[:1], want_status=True)
    assert not ok and list(st) == [5]
    assert not O.bls_verify_signature_sets(b"", b"", b"", np.array([0]), [])
