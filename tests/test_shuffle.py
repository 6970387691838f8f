"""swap-or-not shuffle: oracle pinned by an independent hashlib restatement of the spec's
compute_shuffled_index and by the reference's own list/index equivalences (shuffle_list.rs tests :163-215);
GPU kernel bit-exact against the oracle."""
import hashlib

import numpy as np
import pytest

from tests import oracle_lib as O


def py_csi(index, n, seed, rounds):
    for r in range(rounds):
        pivot = int.from_bytes(hashlib.sha256(seed + bytes([r])).digest()[:8], "little") % n
        flip = (pivot + n - index) % n
        pos = max(index, flip)
        src = hashlib.sha256(seed + bytes([r]) + (pos >> 8).to_bytes(4, "little")).digest()
        if (src[(pos % 256) // 8] >> (pos % 8)) & 1:
            index = flip
    return index


def test_oracle_matches_spec_and_reference_equivalences():
    seed = hashlib.sha256(b"seed").digest()
    assert O.shuffle_list([], 90, seed, True) is None              # returns_none_for_zero_length_list
    assert O.shuffle_list([1, 2], 0, seed, True) is None
    for n in (1, 2, 3, 100, 257, 1000):
        for rounds in (1, 10, 90):
            csi = [py_csi(i, n, seed, rounds) for i in range(n)]
            assert csi == [O.compute_shuffled_index(i, n, seed, rounds) for i in range(n)]
            inp = list(range(1000, 1000 + n))
            assert O.shuffle_list(inp, rounds, seed, False) == [inp[csi[i]] for i in range(n)]
            fw = [0] * n
            for x in range(n):
                fw[csi[x]] = inp[x]
            assert O.shuffle_list(inp, rounds, seed, True) == fw
            assert O.shuffle_list(fw, rounds, seed, False) == inp   # shuffle then un-shuffle is the identity


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1000, 65_537, 500_000])
def test_gpu_shuffle_matches_oracle(gpu, n):
    from lighthouse_b200.shuffle import shuffle_list
    seed = hashlib.sha256(b"gpu-seed-%d" % n).digest()
    rng = np.random.default_rng(n)
    inp = rng.integers(0, 1 << 40, size=n, dtype=np.uint64).tolist()
    for rounds in ((1, 90, 97, 161, 255) if n < 100_000 else (90,)):   # > 96 rounds: pivot table / seed slot sizing
        for forwards in (False, True):
            assert shuffle_list(inp, rounds, seed, forwards) == O.shuffle_list(inp, rounds, seed, forwards)
    assert shuffle_list(shuffle_list(inp, 90, seed, True), 90, seed, False) == inp


@pytest.mark.gpu
def test_gpu_shuffle_none_cases(gpu):
    from lighthouse_b200.shuffle import shuffle_list
    seed = bytes(32)
    assert shuffle_list([], 90, seed, True) is None
    assert shuffle_list([1, 2, 3], 0, seed, True) is None
