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
