"""GPU parity tests for the BLS path: CUDA library (through the C ABI) vs the big-integer oracle
(oracle/bls_ref.py), the reference's golden vectors, and the derivable cases of crypto/bls/tests/tests.rs."""
import hashlib
import json

import numpy as np
import pytest

from tests import oracle_lib as O
from oracle import bls_ref as B

pytestmark = pytest.mark.gpu


def secret_from_u64(i):
    """crypto/bls/tests/tests.rs:15-20: big-endian (i + 1)"""
    return (i + 1).to_bytes(32, "big")


def deposit_signing_root(d):
    fv = bytes.fromhex(d["fork_version"])
    fdr = hashlib.sha256(fv + bytes(28) + bytes(32)).digest()  # deposit domain: zero gvr (chain_spec.rs:498-500)
    return hashlib.sha256(bytes.fromhex(d["deposit_message_root"]) + bytes([3, 0, 0, 0]) + fdr[:28]).digest()


@pytest.fixture(scope="module")
def bls(gpu):
    from lighthouse_b200 import bls as m
    return m


def test_interop_keypairs_golden(bls):
    kp = O.golden_json("interop_keypairs.json")
    sks = b"".join(bytes.fromhex(k["privkey"][2:]) for k in kp)
    pk48, pk96 = bls.sk_to_pk(sks)
    for i, k in enumerate(kp):
        assert pk48[48 * i:48 * i + 48].hex() == k["pubkey"][2:]
        assert pk96[96 * i:96 * i + 96] == B.g1_uncompressed(B.g1_decompress(bytes.fromhex(k["pubkey"][2:])))
        pk = bls.PublicKey.deserialize(bytes.fromhex(k["pubkey"][2:]))
        assert pk.serialize_uncompressed() == pk96[96 * i:96 * i + 96]


def test_sign_matches_oracle(bls):
    for i in range(3):
        sk = int.from_bytes(secret_from_u64(i), "big")
        msg = hashlib.sha256(b"msg%d" % i).digest()
        assert bls.sign(secret_from_u64(i), msg) == B.g2_compress(B.sign(sk, msg))


def test_deposit_vectors_golden(bls):
    """22 (pubkey, message, signature) triples from validator_manager/test_vectors must verify
    (create_validators.rs:749-752), individually and as one batch; a changed message must not."""
    deps = O.golden_json("deposit_data.json")
    sets = []
    for d in deps:
        pk = bls.PublicKey.deserialize(bytes.fromhex(d["pubkey"]))
        sig = bls.Signature.deserialize(bytes.fromhex(d["signature"]))
        sets.append(bls.SignatureSet.single_pubkey(sig, pk, deposit_signing_root(d)))
    assert len(sets) == 22
    for s in sets[:4]:
        assert s.verify()
    assert bls.verify_signature_sets(sets)
    bad = list(sets)
    bad[7] = bls.SignatureSet.single_pubkey(sets[7].signature, sets[7].signing_keys[0], bytes(32))
    assert not bls.verify_signature_sets(bad)
    assert not bad[7].verify()
    # ParallelSignatureSets (block_signature_verifier.rs:84-96,392-418): accumulate, then one batch verify
    acc = bls.ParallelSignatureSets()
    assert not acc.verify()                       # nothing included: verify_signature_sets(empty) is false
    for st in sets:
        acc.push(st)
    assert len(acc) == 22 and acc.verify()
    acc.push(bad[7])
    assert not acc.verify()


def test_gt_value_matches_oracle(bls):
    """Bit-exact intermediate: the final-exponentiated batch product (device computes the cube)."""
    n, k = 3, 2
    sks = [[int.from_bytes(secret_from_u64(3 * i + j), "big") for j in range(k)] for i in range(n)]
    msgs = [hashlib.sha256(b"m%d" % i).digest() for i in range(n)]
    pks = [[B.sk_to_pk(s) for s in row] for row in sks]
    sigs = [B.g2_compress(B.sign(sum(row) % B.R, m)) for row, m in zip(sks, msgs)]
    rands = [0x0123456789ABCDEF, 1, 0xFFFFFFFFFFFFFFFF]
    batch = bls.Batch(n, n * k)
    offs = np.arange(n + 1, dtype=np.uint32) * k
    batch.upload(b"".join(sigs), b"".join(msgs), b"".join(B.g1_uncompressed(p) for row in pks for p in row), offs, rands)
    batch.enqueue()
    assert batch.result() is True
    f = B.F12_ONE
    acc = None
    for row, m, sb, r in zip(pks, msgs, sigs, rands):
        apk = None
        for p in row:
            apk = B.g1_add(apk, p)
        f = B.f12_mul(f, B.miller_loop(B.g1_mul(apk, r), B.hash_to_g2(m)))
        acc = B.g2_add(acc, B.g2_mul(B.g2_decompress(sb), r))
    f = B.f12_mul(f, B.miller_loop(B.g1_neg(B.G1_GEN), acc))
    gt = B.final_exp(f)
    assert gt == B.F12_ONE
    # an invalid batch gives a non-trivial GT value: compare it limb for limb
    batch.upload(b"".join(sigs), b"".join(msgs[::-1]), b"".join(B.g1_uncompressed(p) for row in pks for p in row),
                 offs, rands)
    batch.enqueue()
    assert batch.result() is False
    f = B.F12_ONE
    for row, m, r in zip(pks, msgs[::-1], rands):
        apk = None
        for p in row:
            apk = B.g1_add(apk, p)
        f = B.f12_mul(f, B.miller_loop(B.g1_mul(apk, r), B.hash_to_g2(m)))
    f = B.f12_mul(f, B.miller_loop(B.g1_neg(B.G1_GEN), acc))
    gt = B.final_exp(f)
    cube = B.f12_mul(B.f12_sqr(gt), gt)
    (a, b, c), (d, e, g) = cube
    want = b"".join(x[0].to_bytes(48, "big") + x[1].to_bytes(48, "big") for x in (a, b, c, d, e, g))
    assert batch.gt_bytes() == want
    batch.destroy()


def make_set(bls, signer_ids, msg, valid=True):
    """tests.rs helper: aggregate signature by the given secret_from_u64 signers over msg"""
    sks = b"".join(secret_from_u64(i) for i in signer_ids)
    pk48, pk96 = bls.sk_to_pk(sks)
    keys = [bls.PublicKey(pk48[48 * i:48 * i + 48], pk96[96 * i:96 * i + 96]) for i in range(len(signer_ids))]
    agg_sk = sum(int.from_bytes(secret_from_u64(i), "big") for i in signer_ids) % B.R
    sig = bls.sign(agg_sk.to_bytes(32, "big"), msg if valid else hashlib.sha256(msg).digest())
    return bls.SignatureSet.multiple_pubkeys(bls.AggregateSignature(sig), keys, msg)


def test_verify_signature_sets_cases(bls):
    """crypto/bls/tests/tests.rs:455-510"""
    m = [hashlib.sha256(bytes([i])).digest() for i in range(4)]
    assert not bls.verify_signature_sets([])                                     # empty iterator (blst.rs:42-44)
    assert bls.verify_signature_sets([make_set(bls, [0], m[0])])                 # 1 set x 1 signer
    assert bls.verify_signature_sets([make_set(bls, [0, 1], m[0])])              # 1 set x 2 signers
    assert bls.verify_signature_sets([make_set(bls, list(range(128)), m[0])])    # 1 set x 128 signers
    assert bls.verify_signature_sets([make_set(bls, [0, 1], m[0]), make_set(bls, [2, 3, 4], m[1])])
    assert not bls.verify_signature_sets([make_set(bls, [0, 1], m[0]), make_set(bls, [2, 3], m[1], valid=False)])
    assert not bls.verify_signature_sets([make_set(bls, [0], m[0], valid=False)])
    # infinity signature in the middle (tests.rs:504-510)
    good1, good2 = make_set(bls, [0], m[0]), make_set(bls, [1], m[1])
    inf = bls.SignatureSet.multiple_pubkeys(bls.AggregateSignature.infinity(), good1.signing_keys, m[2])
    assert not bls.verify_signature_sets([good1, inf, good2])
    # empty signature anywhere -> false, with its status code (blst.rs:79-82)
    emp = bls.SignatureSet.multiple_pubkeys(bls.AggregateSignature.empty(), good1.signing_keys, m[2])
    sigs, msgs, pks, offs = bls.flatten_signature_sets([good1, emp, good2])
    ok, st = bls.verify_signature_sets_raw(sigs, msgs, pks, offs, want_status=True)
    assert not ok and list(st) == [0, 1, 0]
    # a set without signing keys -> false (blst.rs:86-89)
    nokeys = bls.SignatureSet(good1.signature, [], m[0])
    sigs, msgs, pks, offs = bls.flatten_signature_sets([good2, nokeys])
    ok, st = bls.verify_signature_sets_raw(sigs, msgs, pks, offs, want_status=True)
    assert not ok and list(st) == [0, 4]
    # same valid set twice, and explicit random scalars
    assert bls.verify_signature_sets([good1, good1, good2], rands=[1, 2**64 - 1, 0x8000000000000000])
    # aggregate key at infinity (pk + (-pk)) -> false (Appendix C item 5)
    pk = B.sk_to_pk(5)
    sigs = good1.signature.serialize()
    ok, st = bls.verify_signature_sets_raw(sigs, m[0], B.g1_uncompressed(pk) + B.g1_uncompressed(B.g1_neg(pk)),
                                           np.array([0, 2]), want_status=True)
    assert not ok and list(st) == [5]
    # duplicate key inside one set exercises the doubling branch of the aggregation
    sk = int.from_bytes(secret_from_u64(9), "big")
    sig2 = B.g2_compress(B.sign(2 * sk % B.R, m[3]))
    pk9 = B.g1_uncompressed(B.sk_to_pk(sk))
    assert bls.verify_signature_sets_raw(sig2, m[3], pk9 + pk9, np.array([0, 2]))


def test_fast_aggregate_verify_cases(bls):
    """crypto/bls/tests/tests.rs:248-342"""
    msg = hashlib.sha256(b"fav").digest()
    for k in (1, 128):
        s = make_set(bls, list(range(k)), msg)
        assert s.signature.fast_aggregate_verify(msg, s.signing_keys)
        assert not s.signature.fast_aggregate_verify(hashlib.sha256(b"x").digest(), s.signing_keys)
    s = make_set(bls, [0, 1, 2], msg)
    assert not s.signature.fast_aggregate_verify(msg, [])                         # 0 keys
    assert not s.signature.fast_aggregate_verify(msg, s.signing_keys[:2])         # missing signer
    inf = bls.AggregateSignature.infinity()
    assert not inf.fast_aggregate_verify(msg, s.signing_keys)
    assert inf.eth_fast_aggregate_verify(msg, [])                                 # :205-209
    assert not inf.eth_fast_aggregate_verify(msg, s.signing_keys)
    assert not bls.AggregateSignature.empty().fast_aggregate_verify(msg, s.signing_keys)
    assert not bls.AggregateSignature.empty().eth_fast_aggregate_verify(msg, [])


def _non_subgroup_g2():
    x = (3, 1)
    while True:
        y = B.f2_sqrt(B.f2_add(B.f2_mul(B.f2_sqr(x), x), B.B2))
        if y:
            return (x, y)
        x = (x[0] + 1, 1)


def test_signature_subgroup_and_decode_status(bls):
    m = hashlib.sha256(b"s").digest()
    good = make_set(bls, [0], m)
    pk = good.signing_keys[0].serialize_uncompressed()
    bad_pt = B.g2_compress(_non_subgroup_g2())
    ok, st = bls.verify_signature_sets_raw(bad_pt, m, pk, np.array([0, 1]), want_status=True)
    assert not ok and list(st) == [3]
    not_on_curve = bytearray(good.signature.serialize()); not_on_curve[95] ^= 1
    ok, st = bls.verify_signature_sets_raw(bytes(not_on_curve), m, pk, np.array([0, 1]), want_status=True)
    assert not ok and st[0] in (2, 3)
    uncompressed_flag = bytes([good.signature.serialize()[0] & 0x7F]) + good.signature.serialize()[1:]
    ok, st = bls.verify_signature_sets_raw(uncompressed_flag, m, pk, np.array([0, 1]), want_status=True)
    assert not ok and list(st) == [2]
    with pytest.raises(bls.BlstError):
        bls.Signature.deserialize(uncompressed_flag)
    with pytest.raises(bls.InvalidByteLength):
        bls.Signature.deserialize(bytes(95))


def test_public_key_deserialize_rules(bls):
    """tests.rs:344-347 and generic_public_key.rs:86-94"""
    with pytest.raises(bls.InvalidInfinityPublicKey):
        bls.PublicKey.deserialize(bls.INFINITY_PUBLIC_KEY)
    with pytest.raises(bls.InvalidByteLength):
        bls.PublicKey.deserialize(bytes(47))
    # a curve point outside the r-order subgroup must be rejected (key_validate)
    x = 1
    while True:
        y = B.fp_sqrt((x ** 3 + 4) % B.P)
        if y is not None and not B.g1_in_subgroup((x, y)):
            break
        x += 1
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize(B.g1_compress((x, y)))
    # x not on the curve
    x = 2
    while B.fp_sqrt((x ** 3 + 4) % B.P) is not None:
        x += 1
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize(bytes([0x80 | (x >> 376)]) + (x & ((1 << 376) - 1)).to_bytes(47, "big"))
    unc, st = bls.decompress_validate_pubkeys(B.g1_compress(B.G1_GEN) + bls.INFINITY_PUBLIC_KEY)
    assert list(st) == [0, 1] and unc[:96] == B.g1_uncompressed(B.G1_GEN)


def test_attestation_batch_config0_shape(bls):
    """BASELINE configs[0] shape at reduced count: aggregate attestation sets with 128 keys; valid -> True,
    one flipped signature / message / key -> False (per-set statuses stay 0: failure is in the pairing)."""
    from lighthouse_b200.synthetic import attestation_batch
    ab = attestation_batch(96, keys_per_set=128, n_validators=1024, seed=3)
    assert bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, ab.offsets)
    # oracle cross-check of one generated set (the generator uses the device's own sign kernel)
    keys = [B.g1_decompress(B.g1_compress(None)) if False else None]
    pts = []
    for j in range(128):
        raw = ab.pks[96 * j:96 * j + 96]
        pts.append((int.from_bytes(raw[:48], "big"), int.from_bytes(raw[48:], "big")))
    assert B.verify_signature_sets([(ab.sigs[:96], pts, ab.msgs[:32])], [0xDEADBEEFCAFEF00D])
    swapped = ab.sigs[96:192] + ab.sigs[:96] + ab.sigs[192:]
    ok, st = bls.verify_signature_sets_raw(swapped, ab.msgs, ab.pks, ab.offsets, want_status=True)
    assert not ok and not st.any()
    msgs = bytearray(ab.msgs); msgs[32 * 50] ^= 1
    assert not bls.verify_signature_sets_raw(ab.sigs, bytes(msgs), ab.pks, ab.offsets)
    pks = bytearray(ab.pks); pks[96 * 1000:96 * 1001] = ab.pks[96 * 2000:96 * 2001]
    assert not bls.verify_signature_sets_raw(ab.sigs, ab.msgs, bytes(pks), ab.offsets)


def test_ragged_sets_and_linearity(bls):
    """ragged key counts (1..300 keys) in one batch, and the size-independent property
    verify(A ++ B) == verify(A) and verify(B)."""
    rng = np.random.default_rng(9)
    sets = []
    for i, k in enumerate([1, 2, 3, 7, 64, 300, 1, 129]):
        ids = [int(x) for x in rng.choice(400, size=k, replace=False)]
        sets.append(make_set(bls, ids, hashlib.sha256(b"r%d" % i).digest()))
    assert bls.verify_signature_sets(sets)
    bad = make_set(bls, [1, 2], hashlib.sha256(b"bad").digest(), valid=False)
    assert not bls.verify_signature_sets(sets + [bad])
    assert bls.verify_signature_sets(sets[:4]) and bls.verify_signature_sets(sets[4:])
    assert not bls.verify_signature_sets(sets[:4] + [bad] + sets[4:])


def test_pubkey_table_indexed_matches_explicit(bls):
    """device-resident pubkey table (ValidatorPubkeyCache mirror) + u32 indices gives the same
    verdicts as shipping the 96-byte keys, including the failure statuses."""
    from lighthouse_b200.synthetic import attestation_batch
    ab = attestation_batch(64, keys_per_set=32, n_validators=512, seed=5)
    table = bls.PubkeyTable(600)
    table.append(ab.pk_table.tobytes())
    assert len(table) == 512
    idx = ab.committees.reshape(-1).astype(np.uint32)
    b = bls.Batch(64, 64 * 32)
    b.upload_indexed(table, ab.sigs, ab.msgs, idx, ab.offsets)
    b.enqueue()
    assert b.result() is True
    gt_indexed = b.gt_bytes()
    rands = np.arange(1, 65, dtype=np.uint64) * 0x9E3779B97F4A7C15
    b.upload_indexed(table, ab.sigs, ab.msgs, idx, ab.offsets, rands); b.enqueue(); assert b.result(); g1 = b.gt_bytes()
    b.upload(ab.sigs, ab.msgs, ab.pks, ab.offsets, rands); b.enqueue(); assert b.result(); g2 = b.gt_bytes()
    assert g1 == g2 and gt_indexed == g1                       # GT == 1 in all three
    bad = idx.copy(); bad[40] = (bad[40] + 1) % 512            # wrong validator in one committee
    b.upload_indexed(table, ab.sigs, ab.msgs, bad, ab.offsets, rands); b.enqueue()
    ok, st = b.result(want_status=True)
    assert not ok and not st.any()
    bad[40] = 512                                              # index past the table
    b.upload_indexed(table, ab.sigs, ab.msgs, bad, ab.offsets, rands); b.enqueue()
    ok, st = b.result(want_status=True)
    assert not ok and st[40 // 32] == 6
    # import validation: infinity and off-curve keys are refused, nothing is appended
    with pytest.raises(bls.Lhb200Error if hasattr(bls, "Lhb200Error") else Exception):
        table.append(bytes([0x40]) + bytes(95))
    off_curve = bytearray(ab.pk_table[0].tobytes()); off_curve[95] ^= 1
    with pytest.raises(Exception):
        table.append(bytes(off_curve))
    assert len(table) == 512
    b.destroy(); table.destroy()


def test_full_size_batch_properties(bls):
    """BASELINE configs[2] at full size (100 000 sets x 128 keys): size-independent properties — a valid batch
    verifies, any single corrupted unit (signature / message / key index) flips the verdict, and the verdict does
    not depend on the random scalars."""
    from lighthouse_b200.synthetic import attestation_batch
    n, k = 100_000, 128
    ab = attestation_batch(n, keys_per_set=k, n_validators=16384, seed=0xBEEF)
    table = bls.PubkeyTable(16384)
    table.append(ab.pk_table.tobytes())
    idx = ab.committees.reshape(-1).astype(np.uint32)
    b = bls.Batch(n, n * k)
    for seed in (1, 2):
        rands = np.random.default_rng(seed).integers(1, 2 ** 63, size=n, dtype=np.uint64) * 2 + 1
        b.upload_indexed(table, ab.sigs, ab.msgs, idx, ab.offsets, rands)
        b.enqueue()
        assert b.result() is True
    sigs = bytearray(ab.sigs); sigs[96 * 77_777:96 * 77_778] = ab.sigs[96 * 5:96 * 6]      # someone else's signature
    b.upload_indexed(table, bytes(sigs), ab.msgs, idx, ab.offsets); b.enqueue(); assert b.result() is False
    msgs = bytearray(ab.msgs); msgs[32 * 99_999 + 31] ^= 0x80
    b.upload_indexed(table, ab.sigs, bytes(msgs), idx, ab.offsets); b.enqueue(); assert b.result() is False
    bad = idx.copy(); bad[12_345 * k + 7] = (bad[12_345 * k + 7] + 1) % 16384
    b.upload_indexed(table, ab.sigs, ab.msgs, bad, ab.offsets); b.enqueue(); assert b.result() is False
    b.destroy(); table.destroy()


def test_streamed_upload_matches_blocking_upload(bls):
    """lhb200_bls_batch_upload_async (key chunks overlapped with the kernels) gives the verdict of the blocking upload:
    valid batch, a wrong key in the LAST chunk, a wrong key in the first, a set without keys; and the one-shot
    lhb200_verify_signature_sets (which uses the streamed path) agrees."""
    from lighthouse_b200.synthetic import attestation_batch
    n, k = 6000, 128                                    # 73.7 MB of keys -> 2 chunks
    ab = attestation_batch(n, keys_per_set=k, n_validators=4096, seed=0xA51C)
    b = bls.Batch(n, n * k)

    def both(sigs, msgs, pks, offs):
        b.upload(sigs, msgs, pks, offs); b.enqueue(); r0 = b.result()
        b.upload_async(sigs, msgs, pks, offs); b.enqueue(); r1 = b.result()
        assert r0 == r1
        return r1

    assert both(ab.sigs, ab.msgs, ab.pks, ab.offsets) is True
    assert bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, ab.offsets) is True
    for victim in (n * k - 1, 5):
        pks = bytearray(ab.pks)
        other = (victim + 1000) % (n * k)
        pks[96 * victim:96 * victim + 96] = ab.pks[96 * other:96 * other + 96]
        if pks != bytearray(ab.pks):
            assert both(ab.sigs, ab.msgs, bytes(pks), ab.offsets) is False
    offs = ab.offsets.copy(); offs[n - 1] = offs[n]      # set n-2 swallows the keys of n-1, which has none
    assert both(ab.sigs, ab.msgs, ab.pks, offs) is False
    b.destroy()


def test_every_key_aggregation_variant(bls):
    """The key sums run in three forms chosen by batch size (bls_host.cu: slice-parallel <= 8 192 sets, one thread per
    set up to ~38 k, TMA ring above; test_full_size_batch_properties covers the last).  Same verdicts from the first two
    on the same kind of inputs: valid, one foreign key, a set without keys, aggregate key = infinity."""
    from lighthouse_b200.synthetic import attestation_batch
    for n, k in ((700, 37), (9000, 3)):
        ab = attestation_batch(n, keys_per_set=k, n_validators=2048, seed=0x5EED + n)
        assert bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, ab.offsets) is True
        victim = (n // 2) * k + k - 1
        pks = bytearray(ab.pks)
        other = 96 * ((victim + 5 * k) % (n * k))
        pks[96 * victim:96 * victim + 96] = ab.pks[other:other + 96]
        assert pks != bytearray(ab.pks)
        ok, st = bls.verify_signature_sets_raw(ab.sigs, ab.msgs, bytes(pks), ab.offsets, want_status=True)
        assert not ok and not st.any()                       # a pairing failure, not a decode failure
        offs = ab.offsets.copy(); offs[n - 1] = offs[n]
        ok, st = bls.verify_signature_sets_raw(ab.sigs, ab.msgs, ab.pks, offs, want_status=True)
        assert not ok and st[n - 1] != 0 and not st[:n - 2].any()
        pks = bytearray(ab.pks)                              # set 7: key list (P, -P, ...) -> apk = infinity if k == 2; use decode failure instead
        pks[96 * 7 * k] ^= 0x80                              # compression flag on an uncompressed key
        ok, st = bls.verify_signature_sets_raw(ab.sigs, ab.msgs, bytes(pks), ab.offsets, want_status=True)
        assert not ok and st[7] != 0 and not np.delete(st, 7).any()


def test_latency_and_lane_modes_agree(bls):
    """Small batches run on the warp-per-item kernels, large ones on the lane-per-set kernels; the selection
    switches are read once per process, so scripts/mode_probe.py runs in two subprocesses: same verdicts, same statuses and
    the SAME GT BYTES (the value after the final exponentiation is unique) from both families of kernels."""
    import os
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "mode_probe.py")

    def run(extra):
        env = dict(os.environ, **extra)
        out = subprocess.run([sys.executable, probe], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        return [l for l in out.stdout.splitlines() if l.split()[0] in ("valid", "swapped", "negated", "infinity")]

    fast = run({})
    lane = run({"LHB_G2_WARP": "0", "LHB_MILLER_WARP": "0", "LHB_FINAL_WARP": "0"})
    assert fast == lane and len(fast) == 4
    assert fast[0].split()[1] == "True" and all(l.split()[1] == "False" for l in fast[1:])


def test_block_signature_batch_shape(bls):
    """BASELINE configs[3] shape at reduced scale: the sets BlockSignatureVerifier::include_all_signatures collects
    for consecutive blocks (block_signature_verifier.rs:141-393) — 1-key sets (proposal, randao, exits,
    BLS-to-execution changes), committee-sized attestation sets and one 512-key sync aggregate per block — in ONE
    verify_signature_sets call (ParallelSignatureSets::verify, :416-418)."""
    rng = np.random.default_rng(33)
    n_validators = 600
    sks = [int.from_bytes(secret_from_u64(i), "big") for i in range(n_validators)]
    pk48, pk96 = bls.sk_to_pk(b"".join(secret_from_u64(i) for i in range(n_validators)))
    keys = [pk96[96 * i:96 * i + 96] for i in range(n_validators)]
    sets_ids = []
    for blk in range(3):
        sets_ids += [[int(rng.integers(n_validators))] for _ in range(2)]                  # proposal, randao
        sets_ids += [sorted(rng.choice(n_validators, size=61, replace=False).tolist()) for _ in range(8)]  # attestations
        sets_ids += [sorted(rng.choice(n_validators, size=512, replace=True).tolist())]    # sync aggregate (repeats allowed)
        sets_ids += [[int(rng.integers(n_validators))] for _ in range(4)]                  # exits, bls changes
    msgs = [hashlib.sha256(b"blk%d" % i).digest() for i in range(len(sets_ids))]
    agg = [sum(sks[i] for i in ids) % B.R for ids in sets_ids]
    sigs = bls.sign(b"".join(a.to_bytes(32, "big") for a in agg), b"".join(msgs))
    pks = b"".join(keys[i] for ids in sets_ids for i in ids)
    offs = np.concatenate([[0], np.cumsum([len(ids) for ids in sets_ids])]).astype(np.uint32)
    assert bls.verify_signature_sets_raw(sigs, b"".join(msgs), pks, offs)
    for victim in (0, 10, len(sets_ids) - 1):                                              # invalid_signature_* cases
        bad = bytearray(sigs); bad[96 * victim:96 * victim + 96] = sigs[96 * ((victim + 1) % len(sets_ids)):][:96]
        assert not bls.verify_signature_sets_raw(bytes(bad), b"".join(msgs), pks, offs)


def test_cpp_host_layer_on_golden_vectors(gpu, tmp_path):
    """The C++ host mirror (include/lhb200.hpp: bls::SignatureSet, verify_signature_sets, MerkleTree ...) on the
    reference's 22 deposit vectors."""
    import os
    import subprocess
    deps = O.golden_json("deposit_data.json")
    p = tmp_path / "vectors.bin"
    with open(p, "wb") as f:
        for d in deps:
            f.write(bytes.fromhex(d["pubkey"]) + deposit_signing_root(d) + bytes.fromhex(d["signature"]))
    subprocess.check_call(["make", "-C", os.path.join(O.ROOT, "tests", "cpp"), "-s"])
    r = subprocess.run([os.path.join(O.ROOT, "tests", "cpp", "host_mirror_test"), str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "OK 22 sets" in r.stdout


def test_aggregate_signature_add_assign_matches_oracle(bls):
    """TAggregateSignature::add_assign / add_assign_aggregate (blst.rs:230-237) built the way
    crypto/bls/tests/tests.rs:170-188 builds its aggregates: infinity, then add_assign every signer's signature.
    The aggregate's bytes must equal the oracle's G2 sum, and verify like tests.rs:248-342 expects."""
    msg = (42).to_bytes(32, "big")                                   # Hash256::from_low_u64_be(42)
    for n in (1, 2, 5, 33):
        sks = [bls.SecretKey.deserialize(secret_from_u64(i)) for i in range(n)]
        agg = bls.AggregateSignature.infinity()
        for sk in sks:
            agg.add_assign(sk.sign(msg))
        acc = None
        h = B.hash_to_g2(msg)
        for i in range(n):
            acc = B.g2_add(acc, B.g2_mul(h, i + 1))
        assert agg.serialize() == B.g2_compress(acc)
        assert bls.AggregateSignature.aggregate([sk.sign(msg) for sk in sks]).serialize() == B.g2_compress(acc)
        keys = [sk.public_key() for sk in sks]
        assert agg.fast_aggregate_verify(msg, keys)
        assert agg.aggregate_verify([msg] * n, keys)                 # tests.rs:232-240
        if n > 2:
            assert not agg.aggregate_verify([msg] * n, keys[1:] + keys[1:2])    # signer 0 replaced by a duplicate
    # tests.rs:196-220 builder cases
    keys = [bls.SecretKey.deserialize(secret_from_u64(i)).public_key() for i in range(2)]
    agg = bls.AggregateSignature.infinity()
    for i in range(2):
        agg.add_assign(bls.SecretKey.deserialize(secret_from_u64(i)).sign(msg))
    before = agg.serialize()
    agg.add_assign(bls.Signature.empty())                            # aggregate_empty_sig: unchanged
    agg.add_assign_aggregate(bls.AggregateSignature.empty())         # aggregate_empty_agg_sig: unchanged
    agg.add_assign(bls.Signature.deserialize(bls.INFINITY_SIGNATURE))  # aggregate_infinity_sig: unchanged
    assert agg.serialize() == before and agg.fast_aggregate_verify(msg, keys)
    e = bls.AggregateSignature.empty()
    e.add_assign(bls.SecretKey.deserialize(secret_from_u64(0)).sign(msg))   # empty + sig = infinity + sig
    assert e.fast_aggregate_verify(msg, keys[:1])
    assert not bls.AggregateSignature.empty().aggregate_verify([msg], keys[:1])
    assert not bls.AggregateSignature.infinity().aggregate_verify([msg], keys[:1])
    assert not agg.aggregate_verify([], [])                          # generic_aggregate_signature.rs:214-216
    with pytest.raises(bls.BlstError):
        bls.aggregate_signatures(before + bytes([0x80]) + bytes(95))  # not a valid G2 encoding


def test_aggregate_verify_distinct_messages(bls):
    """blst.rs:263-273: one aggregate signature over different messages (the EF aggregate_verify shape)."""
    n = 7
    msgs = [hashlib.sha256(b"agv%d" % i).digest() for i in range(n)]
    sks = [bls.SecretKey.deserialize(secret_from_u64(i)) for i in range(n)]
    agg = bls.AggregateSignature.aggregate([sk.sign(m) for sk, m in zip(sks, msgs)])
    keys = [sk.public_key() for sk in sks]
    assert agg.aggregate_verify(msgs, keys)
    assert not agg.aggregate_verify(msgs[::-1], keys)
    assert not agg.aggregate_verify(msgs[:-1], keys[:-1])
    assert B.multi_pairing_is_one([(B.g1_neg(B.G1_GEN), B.g2_decompress(agg.serialize()))] +
                                  [(B.g1_decompress(k.serialize()), B.hash_to_g2(m)) for k, m in zip(keys, msgs)])


def test_aggregate_public_key_and_uncompressed_deserialize(bls):
    """TAggregatePublicKey::aggregate (blst.rs:178-184) and TPublicKey::deserialize_uncompressed (blst.rs:142-150)."""
    n = 9
    keys = [bls.SecretKey.deserialize(secret_from_u64(i)).public_key() for i in range(n)]
    apk = bls.AggregatePublicKey.aggregate(keys).to_public_key()
    ref = B.g1_mul(B.G1_GEN, sum(range(1, n + 1)))
    assert apk.serialize() == B.g1_compress(ref) and apk.serialize_uncompressed() == B.g1_uncompressed(ref)
    with pytest.raises(bls.BlstError):
        bls.AggregatePublicKey.aggregate([])
    pk = bls.PublicKey.deserialize_uncompressed(keys[3].serialize_uncompressed())
    assert pk == keys[3] and pk.serialize() == keys[3].serialize()
    u = bytearray(keys[3].serialize_uncompressed())
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize_uncompressed(bytes([u[0] | 0x80]) + bytes(u[1:]))    # compression flag set
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize_uncompressed(bytes([u[0] | 0x20]) + bytes(u[1:]))    # sort flag set
    u[95] ^= 1
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize_uncompressed(bytes(u))                               # not on the curve
    with pytest.raises(bls.InvalidInfinityPublicKey):
        bls.PublicKey.deserialize_uncompressed(bytes([0x40]) + bytes(95))
    with pytest.raises(bls.BlstError):
        bls.PublicKey.deserialize_uncompressed(bytes([0x40]) + bytes(94) + b"\x01")    # infinity flag with payload
    with pytest.raises(bls.InvalidByteLength):
        bls.PublicKey.deserialize_uncompressed(bytes(48))


def test_concurrent_callers_overlap_on_the_device(bls):
    """Lighthouse calls verify_signature_sets from up to num_cpus blocking workers with <= 64-set gossip batches
    (beacon_processor/src/lib.rs:202-203,256).  Eight threads hammer lhb200_verify_signature_sets with their own 64-set
    batches (one of them carrying a bad set): every verdict must be right, and the eight callers together must finish
    far sooner than eight times a lone caller (each call borrows its own batch handle and stream)."""
    import threading
    import time
    from lighthouse_b200.synthetic import attestation_batch
    n_thr, n_sets, reps = 8, 64, 6
    batches = []
    for t in range(n_thr):
        ab = attestation_batch(n_sets, keys_per_set=16, n_validators=1024, seed=700 + t)
        sigs = bytearray(ab.sigs)
        if t == 3:
            sigs[96 * 17:96 * 18] = ab.sigs[96 * 18:96 * 19]     # set 17 carries set 18's signature
        batches.append((bytes(sigs), ab.msgs, ab.pks, ab.offsets, t != 3))
    for b in batches:                                              # warm the handle pool / caches
        assert bls.verify_signature_sets_raw(*b[:4]) == b[4]

    def worker(b, out, k):
        for _ in range(reps):
            out[k] = out[k] and (bls.verify_signature_sets_raw(*b[:4]) == b[4])

    def run_all(res):
        ths = [threading.Thread(target=worker, args=(batches[k], res, k)) for k in range(n_thr)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return time.perf_counter() - t0

    res = [True] * n_thr
    run_all(res)                                                   # first concurrent pass creates the pool's handles
    t0 = time.perf_counter()
    worker(batches[0], res, 0)
    t_one = time.perf_counter() - t0
    t_all = run_all(res)
    assert all(res)
    speedup = n_thr * t_one / t_all
    print(f"1 caller: {reps * n_sets / t_one:.0f} sets/s; {n_thr} callers: {n_thr * reps * n_sets / t_all:.0f} sets/s ({speedup:.2f}x)")
    assert speedup > 2.5, (t_one, t_all)
