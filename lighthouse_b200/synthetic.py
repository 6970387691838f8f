"""Synthetic mainnet-shape inputs: BeaconStateDeneb SSZ bytes for the tree-hash path.
Pure byte layout with numpy — no hashing happens here."""
import struct

import numpy as np

DENEB_FIXED = 2736653
U64_MAX = (1 << 64) - 1


def validators_ssz(n, rng, pubkeys=None):
    """n x 121-byte Validator records shaped like consensus/types/benches/benches.rs:11-47."""
    v = np.zeros((n, 121), dtype=np.uint8)
    if pubkeys is None:
        v[:, 0:48] = rng.integers(0, 256, size=(n, 48), dtype=np.uint8)
    else:
        v[:, 0:48] = np.frombuffer(pubkeys, dtype=np.uint8).reshape(n, 48)
    idx = np.arange(n, dtype="<u8")
    v[:, 48 + 24:48 + 32] = idx.view(np.uint8).reshape(n, 8)  # H256::from_low_u64_le(i)
    v[:, 80:88] = np.frombuffer(struct.pack("<Q", 32_000_000_000), dtype=np.uint8)
    v[:, 88] = 0
    v[:, 89:105] = 0  # activation_eligibility_epoch = activation_epoch = 0
    v[:, 105:121] = 0xFF  # exit_epoch = withdrawable_epoch = u64::MAX
    return v.tobytes()


def beacon_state_deneb_ssz(n_validators, seed=1, all_default=False, n_hist_roots=758, n_votes=1024,
                           n_summaries=600, extra_data_len=14):
    """SSZ(BeaconStateDeneb), mainnet preset.  all_default=True leaves every non-validator field zero so the
    ZERO_HASHES ladder paths are exercised."""
    rng = np.random.default_rng(seed)
    V = n_validators

    def rnd(nbytes):
        if all_default:
            return bytes(nbytes)
        return rng.integers(0, 256, size=nbytes, dtype=np.uint8).tobytes()

    if all_default:
        n_hist_roots = n_votes = n_summaries = extra_data_len = 0
    hist = rnd(32 * n_hist_roots)
    votes = rnd(72 * n_votes)
    vals = validators_ssz(V, rng)
    bal = np.arange(V, dtype="<u8").tobytes()
    if all_default:
        pp = bytes(V)
        cp = bytes(V)
        inact = bytes(8 * V)
        slash = bytes(8192 * 8)
    else:
        pp = rng.integers(0, 8, size=V, dtype=np.uint8).tobytes()
        cp = rng.integers(0, 8, size=V, dtype=np.uint8).tobytes()
        inact = rng.integers(0, 64, size=V, dtype="<u8").tobytes()
        slash = rng.integers(0, 1 << 40, size=8192, dtype="<u8").tobytes()
    leph = (rnd(32) + rnd(20) + rnd(32) + rnd(32) + rnd(256) + rnd(32) + rnd(8) + rnd(8) + rnd(8) + rnd(8)
            + struct.pack("<I", 584) + rnd(32) + rnd(32) + rnd(32) + rnd(32) + rnd(8) + rnd(8) + rnd(extra_data_len))
    assert len(leph) == 584 + extra_data_len
    summ = rnd(64 * n_summaries)

    o_hist = DENEB_FIXED
    o_votes = o_hist + len(hist)
    o_val = o_votes + len(votes)
    o_bal = o_val + len(vals)
    o_pp = o_bal + len(bal)
    o_cp = o_pp + len(pp)
    o_inact = o_cp + len(cp)
    o_leph = o_inact + len(inact)
    o_hs = o_leph + len(leph)
    u32 = lambda x: struct.pack("<I", x)
    fixed = b"".join([
        rnd(8), rnd(32), rnd(8),                       # genesis_time, genesis_validators_root, slot
        rnd(16),                                       # fork
        rnd(112),                                      # latest_block_header
        rnd(8192 * 32), rnd(8192 * 32),                # block_roots, state_roots
        u32(o_hist), rnd(72), u32(o_votes), rnd(8),    # historical_roots, eth1_data, eth1_data_votes, deposit idx
        u32(o_val), u32(o_bal),
        rnd(65536 * 32), slash,                        # randao_mixes, slashings
        u32(o_pp), u32(o_cp),
        bytes([0 if all_default else 0x0B]),           # justification_bits
        rnd(40), rnd(40), rnd(40),                     # checkpoints
        u32(o_inact),
        (vals[:48] * 1 if False else rnd(513 * 48)),   # current_sync_committee (512 pubkeys + aggregate)
        rnd(513 * 48),                                 # next_sync_committee
        u32(o_leph), rnd(8), rnd(8), u32(o_hs),
    ])
    assert len(fixed) == DENEB_FIXED, len(fixed)
    return b"".join([fixed, hist, votes, vals, bal, pp, cp, inact, leph, summ])
