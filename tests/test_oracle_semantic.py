"""CPU: the numpy restatement of the 3DMM window helpers (oracle/semantic.py) against the fixtures produced by the
reference's own functions (tests/golden/semantic_golden.npz, oracle/make_golden_semantic.py), bit-for-bit; plus the
host-side helpers of the drop-in module that need no GPU."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import semantic as osem


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "semantic_golden.npz"))


@pytest.mark.parametrize("tag,dtype,n", [("f32", np.float32, 40), ("f64", np.float64, 17)])
def test_oracle_matches_reference_fixtures(gold, tag, dtype, n):
    table = osem.synth_table(n, seed=3, dtype=dtype)
    ratio = osem.find_crop_norm_ratio(table[0:1], table[1:])
    assert ratio.dtype == dtype and np.array_equal(ratio, gold[tag + "_ratio"])
    frames = gold[tag + "_frames"]
    plain = np.stack([osem.transform_semantic(table, int(i)) for i in frames])
    scaled = np.stack([osem.transform_semantic(table, int(i), ratio) for i in frames])
    assert plain.dtype == np.float32 and np.array_equal(plain, gold[tag + "_plain"])
    assert np.array_equal(scaled, gold[tag + "_scaled"])
    assert not np.array_equal(plain, scaled)
    assert np.array_equal(osem.transform_semantic(table, 3, np.zeros(1, dtype)), gold[tag + "_zero_ratio"])
    assert np.array_equal(gold[tag + "_zero_ratio"], osem.transform_semantic(table, 3))     # `if crop_norm_ratio:` is False for 0


def test_seq_index_edges():
    assert osem.obtain_seq_index(0, 40) == [0] * 14 + list(range(1, 13))
    assert osem.obtain_seq_index(39, 40) == list(range(26, 40)) + [39] * 12
    assert osem.obtain_seq_index(5, 1) == [0] * 26                      # single-frame table


def test_dropin_host_helpers_match_oracle():
    from s2v_b200.futils import inference_utils as iu
    for i, n in ((0, 40), (39, 40), (7, 3)):
        assert iu.obtain_seq_index(i, n) == osem.obtain_seq_index(i, n)
    t = osem.synth_table(25, seed=1)
    assert np.array_equal(iu.find_crop_norm_ratio(t[2:3], t), osem.find_crop_norm_ratio(t[2:3], t))
    assert iu._ratio_args(None) == (0.0, 0) and iu._ratio_args(np.zeros(1, np.float32)) == (0.0, 0)
    assert iu._ratio_args(np.array([1.5], np.float32)) == (1.5, 1)
    with pytest.raises(ValueError):
        iu._ratio_args(np.ones(2))


def test_transform_semantic_property():
    """Against the literal per-element definition (inference_utils.py:73-91) for random tables, frame indices and ratios."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(1, 60), idx=st.integers(0, 59), seed=st.integers(0, 1000), use_ratio=st.booleans())
    def run(n, idx, seed, use_ratio):
        idx = idx % n
        t = osem.synth_table(n, seed=seed)
        ratio = np.array([1.0 + (seed % 7) * 0.125], np.float32) if use_ratio else None
        got = osem.transform_semantic(t, idx, ratio)
        cols = list(range(80, 144)) + [224, 225, 226, 254, 255, 256, 259, 260, 261]
        for j in range(26):
            row = min(max(idx - 13 + j, 0), n - 1)
            for r, c in enumerate(cols):
                v = t[row, c]
                if use_ratio and r == 70:
                    v = np.float32(v * ratio[0])
                assert got[r, j] == np.float32(v)
    run()
