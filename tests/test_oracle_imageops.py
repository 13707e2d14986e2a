"""CPU: the image-glue restatement (oracle/imageops.py) against the reference's own lines executed with the real cv2
(tests/golden/imageops_golden.npz, oracle/make_golden_imageops.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import imageops as io


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "imageops_golden.npz"))


def test_resize_u8_bit_exact_and_f32(gold):
    i = 0
    while f"rs_u8_{i}_in" in gold:
        ref = gold[f"rs_u8_{i}"]
        assert np.array_equal(io.resize_linear_u8(gold[f"rs_u8_{i}_in"], *ref.shape[:2]), ref), i
        f = gold[f"rs_f32_{i}_in"]
        assert np.abs(io.resize_linear_f32(f, *ref.shape[:2]) - gold[f"rs_f32_{i}"]).max() <= 1e-4, i
        assert np.abs(io.resize_linear_f32(f[:, :, 0], *ref.shape[:2]) - gold[f"rs_f32c1_{i}"]).max() <= 1e-4, i
        # OpenCV's own (non-IPP) code path uses float coordinates: within the float rounding of the coordinate (~1e-5 * 255 per unit slope)
        assert np.abs(io.resize_linear_f32(f, *ref.shape[:2], coords="f32") - gold[f"rs_f32_{i}"]).max() <= 2e-2, i
        i += 1
    assert i == 8


def test_resize_u8_random_shapes_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(60):
        h, w = rng.integers(1, 50, 2)
        oh, ow = rng.integers(1, 70, 2)
        x = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(io.resize_linear_u8(x, int(oh), int(ow)), cv2.resize(x, (int(ow), int(oh))).reshape(oh, ow, 3)), (h, w, oh, ow)


def test_fake_to_bgr_and_face_batch_and_compose(gold):
    assert np.array_equal(io.fake_to_bgr_u8(gold["fake_in"][0]), gold["fake_bgr"])
    ib, orig = io.face_batch([gold[f"oface_{i}"] for i in range(3)], [gold[f"face_{i}"] for i in range(3)], 48)
    assert ib.dtype == np.float32 and np.array_equal(ib, gold["img_batch"]) and np.array_equal(orig, gold["img_original"])
    assert np.array_equal(io.compose_pred_u8(gold["pred_in"], ib, orig, True), gold["pred_u8_composed"])
    assert np.array_equal(io.compose_pred_u8(gold["pred_in"], ib, orig, False), gold["pred_u8_plain"])


def test_paste_and_blend_back(gold):
    ff = io.paste_resized(gold["pred_u8_composed"][0], gold["frame_in"], tuple(int(v) for v in gold["box"]))
    assert np.array_equal(ff, gold["frame_pasted"])
    pp = io.blend_paste_back(gold["restored_in"], ff, gold["mouse_mask_in"], 10)
    d = np.abs(pp.astype(np.int32) - gold["blend_back"].astype(np.int32))
    # float32 pyramid / resize arithmetic, then truncation to uint8: off by one LSB where the float lands next to an integer
    assert d.max() <= 1 and (d > 0).mean() < 0.01
