"""CPU: the numpy restatement of the Laplacian-pyramid blend (oracle/blend.py) against fixtures produced by the reference's
own function running on the real cv2 (tests/golden/blend_golden.npz, oracle/make_golden_blend.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import blend


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "blend_golden.npz"))


def test_pyr_down_up_match_cv2(gold):
    for i in range(7):
        x, f = gold[f"down_u8_{i}_in"], gold[f"down_f32_{i}_in"]
        assert np.array_equal(blend.pyr_down(x), gold[f"down_u8_{i}"]), i              # 8-bit: bit-exact
        assert np.abs(blend.pyr_down(f) - gold[f"down_f32_{i}"]).max() <= 1e-6, i
        assert np.abs(blend.pyr_up(f) - gold[f"up_f32_{i}"]).max() <= 1e-6, i


@pytest.mark.parametrize("key,hw,seed,levels", [("blend64_l6", (64, 64), 0, 6), ("blend64_l7", (64, 64), 0, 7), ("blend48x80_l4", (48, 80), 2, 4)])
def test_blend_matches_reference(gold, key, hw, seed, levels):
    A, B, m = blend.synth_images(*hw, seed=seed)
    got = blend.laplacian_pyramid_blending_with_mask(A, B, m, levels)
    assert got.dtype == np.float32 and got.shape == gold[key].shape
    assert np.abs(got - gold[key]).max() <= 1e-4          # 0..255 scale, float32 summation order only


def test_blend_512_levels_10(gold):
    A, B, m = blend.synth_images(512, 512, seed=1)
    got = blend.laplacian_pyramid_blending_with_mask(A, B, m, 10)
    assert np.abs(got[::37] - gold["blend512_l10_rows"]).max() <= 1e-4
    assert abs(got.astype(np.float64).sum() - float(gold["blend512_l10_sum"])) <= 1e-6 * abs(float(gold["blend512_l10_sum"]))
    # size-independent properties: identical images blend to themselves; mask 1 -> A, mask 0 -> B
    same = blend.laplacian_pyramid_blending_with_mask(A, A, m, 10)
    assert np.abs(same - A.astype(np.float32)).max() <= 2e-3
    assert np.abs(blend.laplacian_pyramid_blending_with_mask(A, B, np.ones_like(m), 10) - A.astype(np.float32)).max() <= 2e-3
    assert np.abs(blend.laplacian_pyramid_blending_with_mask(A, B, np.zeros_like(m), 10) - B.astype(np.float32)).max() <= 2e-3
