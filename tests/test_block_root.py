"""BeaconBlockDeneb canonical_root.  The oracle's hand-unrolled C restatement is pinned by the
generic from-spec hashlib merkleization in tests/ssz_spec.py (the reference's own pin, EF ssz_static, is not on disk);
the CUDA path is compared bit-exactly with the oracle through the C ABI."""
import pytest

from lighthouse_b200 import ssz_schema as S
from lighthouse_b200 import synthetic
from tests import oracle_lib as O
from tests import ssz_spec

EMPTY = dict(n_attestations=0, n_transactions=0, n_proposer_slashings=0, n_attester_slashings=0, n_deposits=0, n_exits=0,
             n_bls_changes=0, n_withdrawals=0, n_blobs=0, extra_data_len=0)
CASES = {
    "mainnet_like": dict(),
    "empty_body": EMPTY,
    "full_operations": dict(seed=7, n_attester_slashings=2, n_proposer_slashings=16, n_deposits=16, n_exits=16,
                            n_bls_changes=16, extra_data_len=32, n_blobs=64,
                            tx_sizes=[0, 1, 31, 32, 33, 64, 255, 256, 257, 288, 8191, 8192, 8193, 100_000, 1 << 20]),
    "max_committee": dict(seed=9, committee=2048, n_attestations=9, tx_sizes=[5]),
    "bit_boundaries": dict(seed=11, committee=250, n_attestations=16, n_transactions=3),  # bit lengths 250..256
    "one_empty_tx": dict(seed=12, tx_sizes=[0], n_attestations=1),
}


def _block(name):
    return synthetic.beacon_block_deneb(**CASES[name])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_block_root_matches_spec_merkleization(name):
    value, ssz = _block(name)
    want = ssz_spec.hash_tree_root(S.BeaconBlockDeneb, value)
    want_body = ssz_spec.hash_tree_root(S.BeaconBlockBodyDeneb, value["body"])
    assert O.beacon_block_root_deneb(ssz) == (want, want_body)


def _malformed():
    _, ssz = _block("mainnet_like")
    bad = []
    bad.append(ssz[:83])                                          # shorter than the fixed part
    b = bytearray(ssz); b[80:84] = (85).to_bytes(4, "little"); bad.append(bytes(b))      # body offset != 84
    b = bytearray(ssz); b[84 + 208:84 + 212] = (1 << 30).to_bytes(4, "little"); bad.append(bytes(b))  # attestations offset past end
    b = bytearray(ssz); b[-1:] = b""; bad.append(bytes(b))        # kzg commitments no longer a multiple of 48
    return bad


def test_oracle_rejects_malformed_blocks():
    for b in _malformed():
        assert O.beacon_block_root_deneb(b) is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_block_root_matches_oracle(gpu, name):
    from lighthouse_b200 import tree_hash
    _, ssz = _block(name)
    assert tree_hash.beacon_block_root_deneb(ssz, want_body_root=True) == O.beacon_block_root_deneb(ssz)


@pytest.mark.gpu
def test_gpu_block_batch_of_an_epoch_matches_oracle(gpu):
    """32 blocks in one pass (BASELINE configs[3] shape): every root equals the oracle's, order preserved."""
    from lighthouse_b200 import tree_hash
    blocks = [synthetic.beacon_block_deneb(seed=100 + i, n_transactions=100 + 5 * i)[1] for i in range(32)]
    roots, bodies = tree_hash.beacon_block_roots_deneb(blocks, want_body_roots=True)
    want = [O.beacon_block_root_deneb(b) for b in blocks]
    assert roots == [w[0] for w in want] and bodies == [w[1] for w in want]
    assert tree_hash.beacon_block_roots_deneb(blocks[5:6]) == [want[5][0]]


@pytest.mark.gpu
def test_gpu_block_root_large_transaction(gpu):
    from lighthouse_b200 import tree_hash
    _, ssz = synthetic.beacon_block_deneb(seed=3, n_attestations=2, tx_sizes=[3_000_001, 17])
    assert tree_hash.beacon_block_root_deneb(ssz) == O.beacon_block_root_deneb(ssz)[0]


@pytest.mark.gpu
@pytest.mark.parametrize("tx_sizes", [[0] * 50_000, [1, 0, 3] * 9_000], ids=["empty_txs", "tiny_txs"])
def test_gpu_block_of_many_tiny_transactions(gpu, tx_sizes):
    """Valid SSZ with two tree nodes per 4-byte offset (far outside any gas limit): the offset pre-scan sizes the arena
    for it (round 1 refused this shape with EINVAL), alone and inside a batch with ordinary blocks."""
    from lighthouse_b200 import tree_hash
    _, ssz = synthetic.beacon_block_deneb(seed=21, n_attestations=3, tx_sizes=tx_sizes)
    want = O.beacon_block_root_deneb(ssz)
    assert tree_hash.beacon_block_root_deneb(ssz, want_body_root=True) == want
    other = synthetic.beacon_block_deneb(seed=22)[1]
    assert tree_hash.beacon_block_roots_deneb([other, ssz, other]) == [O.beacon_block_root_deneb(other)[0], want[0],
                                                                       O.beacon_block_root_deneb(other)[0]]


@pytest.mark.gpu
def test_gpu_rejects_malformed_blocks(gpu):
    from lighthouse_b200 import tree_hash, Lhb200Error
    from lighthouse_b200._ffi import EINVAL
    for b in _malformed():
        with pytest.raises(Lhb200Error) as e:
            tree_hash.beacon_block_root_deneb(b)
        assert e.value.code == EINVAL


def test_oracle_block_root_random_shapes():
    """Seeded sweep over block shapes (list lengths 0..limit, transaction sizes across chunk and tile boundaries,
    committee sizes across bit-list byte boundaries): oracle == generic spec merkleization, and the generic decoder
    round-trips the serialisation."""
    import numpy as np
    rng = np.random.default_rng(2024)
    for it in range(24):
        kw = dict(seed=1000 + it, n_attestations=int(rng.integers(0, 129)), n_proposer_slashings=int(rng.integers(0, 17)),
                  n_attester_slashings=int(rng.integers(0, 3)), n_deposits=int(rng.integers(0, 17)),
                  n_exits=int(rng.integers(0, 17)), n_bls_changes=int(rng.integers(0, 17)),
                  n_withdrawals=int(rng.integers(0, 17)), n_blobs=int(rng.integers(0, 7)),
                  committee=int(rng.integers(1, 2049)), extra_data_len=int(rng.integers(0, 33)),
                  tx_sizes=[int(x) for x in rng.choice([0, 1, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 8191, 8192, 8193, 40000],
                                                       size=int(rng.integers(0, 40)))])
        value, ssz = synthetic.beacon_block_deneb(**kw)
        assert S.serialize(S.BeaconBlockDeneb, ssz_spec.deserialize(S.BeaconBlockDeneb, ssz)) == ssz
        want = (ssz_spec.hash_tree_root(S.BeaconBlockDeneb, value), ssz_spec.hash_tree_root(S.BeaconBlockBodyDeneb, value["body"]))
        assert O.beacon_block_root_deneb(ssz) == want, kw


def _blinded(name):
    value, ssz = _block(name)
    payload_t = dict(S.ExecutionPayloadDeneb[1])
    p = value["body"]["execution_payload"]
    bv, bssz = synthetic.blind_block_deneb(value, ssz_spec.hash_tree_root(payload_t["transactions"], p["transactions"]),
                                           ssz_spec.hash_tree_root(payload_t["withdrawals"], p["withdrawals"]))
    return ssz, bv, bssz


@pytest.mark.parametrize("name", ["mainnet_like", "empty_body", "full_operations"])
def test_oracle_blinded_block_root_equals_full_block_root(name):
    """BlindedBeaconBlock (beacon_block.rs:80): replacing the payload by its header leaves the root unchanged; the
    oracle's blinded walk agrees with the full-block walk and with the generic spec merkleization."""
    ssz, bv, bssz = _blinded(name)
    full = O.beacon_block_root_deneb(ssz)
    assert O.blinded_beacon_block_root_deneb(bssz) == full
    assert ssz_spec.hash_tree_root(S.BlindedBeaconBlockDeneb, bv) == full[0]
    assert O.blinded_beacon_block_root_deneb(ssz) is None      # a full block is not a valid blinded block


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mainnet_like", "empty_body", "full_operations"])
def test_gpu_blinded_block_root_matches_oracle(gpu, name):
    from lighthouse_b200 import tree_hash
    ssz, _, bssz = _blinded(name)
    want = O.blinded_beacon_block_root_deneb(bssz)
    assert tree_hash.beacon_block_root_deneb(bssz, want_body_root=True, blinded=True) == want
    assert want[0] == tree_hash.beacon_block_root_deneb(ssz)   # == the full block's root on the device as well


@pytest.mark.gpu
@pytest.mark.parametrize("fork", ["altair", "bellatrix", "capella", "deneb"])
def test_gpu_block_roots_every_fork(gpu, fork):
    """The BeaconBlock superstruct variants (beacon_block.rs:41-90, beacon_block_body.rs:43-110) through the
    fork-parametrised describer, full and blinded, single and as a batch: roots and body roots against the GENERIC
    from-spec merkleization of tests/ssz_spec.py over the value (type descriptors in lighthouse_b200/ssz_schema.py)."""
    from lighthouse_b200 import tree_hash
    from lighthouse_b200._ffi import EINVAL
    from lighthouse_b200 import Lhb200Error
    t, bt = S.BEACON_BLOCK_BY_FORK[fork], S.BEACON_BLOCK_BODY_BY_FORK[fork]
    shapes = [dict(seed=31), dict(seed=32, **EMPTY), dict(seed=33, n_attestations=5, tx_sizes=[0, 31, 32, 33, 4097], committee=2048)]
    values, blobs = zip(*[synthetic.beacon_block_deneb(fork=fork, **kw) for kw in shapes])
    want = [ssz_spec.hash_tree_root(t, v) for v in values]
    want_body = [ssz_spec.hash_tree_root(bt, v["body"]) for v in values]
    roots, bodies = tree_hash.beacon_block_roots(blobs, fork, want_body_roots=True)
    assert roots == want and bodies == want_body
    assert tree_hash.beacon_block_roots(blobs[1:2], fork) == want[1:2]
    if fork == "deneb":
        assert tree_hash.beacon_block_roots_deneb(blobs) == want
    if fork != "altair":
        pt = dict(S.EXECUTION_PAYLOAD_BY_FORK[fork][1])
        blinded = []
        for v in values:
            ep = v["body"]["execution_payload"]
            wr = ssz_spec.hash_tree_root(pt["withdrawals"], ep["withdrawals"]) if "withdrawals" in pt else bytes(32)
            blinded.append(synthetic.blind_block_deneb(v, ssz_spec.hash_tree_root(pt["transactions"], ep["transactions"]), wr,
                                                       fork=fork)[1])
        assert tree_hash.beacon_block_roots(blinded, fork, blinded=True) == want
    # the bytes of one fork are not silently hashed as another
    other = {"altair": "capella", "bellatrix": "deneb", "capella": "bellatrix", "deneb": "capella"}[fork]
    try:
        assert tree_hash.beacon_block_roots(blobs[:1], other) != want[:1]
    except Lhb200Error as e:
        assert e.code == EINVAL


@pytest.mark.parametrize("fork", ["altair", "bellatrix", "capella", "deneb"])
def test_block_fork_schemas_roundtrip_and_blinded_root(fork):
    """Host side of the fork variants (no GPU): the generator's SSZ decodes and re-encodes under the fork's type descriptor,
    and the blinded block (payload replaced by its header) has the full block's root under the generic merkleization."""
    t = S.BEACON_BLOCK_BY_FORK[fork]
    v, ssz = synthetic.beacon_block_deneb(seed=41, n_attestations=3, n_transactions=5, fork=fork)
    assert S.serialize(t, ssz_spec.deserialize(t, ssz)) == ssz
    assert len(S.BEACON_BLOCK_BODY_BY_FORK[fork][1]) == {"altair": 9, "bellatrix": 10, "capella": 11, "deneb": 12}[fork]
    if fork != "altair":
        pt = dict(S.EXECUTION_PAYLOAD_BY_FORK[fork][1])
        ep = v["body"]["execution_payload"]
        wr = ssz_spec.hash_tree_root(pt["withdrawals"], ep["withdrawals"]) if "withdrawals" in pt else bytes(32)
        bv, _ = synthetic.blind_block_deneb(v, ssz_spec.hash_tree_root(pt["transactions"], ep["transactions"]), wr, fork=fork)
        assert ssz_spec.hash_tree_root(S.BLINDED_BEACON_BLOCK_BY_FORK[fork], bv) == ssz_spec.hash_tree_root(t, v)


def test_electra_state_schema_roundtrip():
    """BeaconStateElectra as in this reference revision (beacon_state.rs:487-525): 37 fields, fixed part 2 736 713 bytes,
    generator output decodes and re-encodes under the type descriptor."""
    from lighthouse_b200.synthetic import beacon_state_deneb_ssz
    typ = S.BEACON_STATE_BY_FORK["electra"]
    assert len(typ[1]) == 37
    ssz = beacon_state_deneb_ssz(3, seed=5, fork="electra", n_pending=(2, 1, 3))
    value = ssz_spec.deserialize(typ, ssz)
    assert S.serialize(typ, value) == ssz
    assert [len(value[k]) for k in ("pending_balance_deposits", "pending_partial_withdrawals", "pending_consolidations")] == [2, 1, 3]
    assert len(beacon_state_deneb_ssz(0, fork="electra", all_default=True)) == 2736713 + 648
