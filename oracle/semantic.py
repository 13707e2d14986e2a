"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the reference's 3DMM coefficient-window helpers.

Follows /root/reference futils/inference_utils.py:73-76 (obtain_seq_index), :78-91 (transform_semantic) and
:93-99 (find_crop_norm_ratio).  Pinned: tests/golden/semantic_golden.npz holds the outputs of the reference's own
functions (their source executed unmodified by tests/golden/make_semantic_golden.py) on seeded tables; the oracle must
reproduce them bit-for-bit (tests/test_oracle_semantic.py).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module.
"""
import numpy as np


def obtain_seq_index(index, num_frames):
    """inference_utils.py:73-76"""
    return [min(max(i, 0), num_frames - 1) for i in range(index - 13, index + 13)]


def transform_semantic(semantic, frame_index, crop_norm_ratio=None):
    """inference_utils.py:78-91 -> float32 [73, 26] (numpy instead of torch.Tensor(...).permute(1, 0))."""
    rows = semantic[obtain_seq_index(frame_index, semantic.shape[0]), ...]           # fancy index: a copy
    ex, ang, tr, crop = rows[:, 80:144], rows[:, 224:227], rows[:, 254:257], rows[:, 259:262].copy()
    if crop_norm_ratio is not None and bool(np.asarray(crop_norm_ratio).reshape(-1)[0]):   # `if crop_norm_ratio:` (:87)
        crop[:, -3] = crop[:, -3] * crop_norm_ratio
    return np.ascontiguousarray(np.concatenate([ex, ang, tr, crop], 1).astype(np.float32).T)


def find_crop_norm_ratio(source_coeff, target_coeffs):
    """inference_utils.py:93-99"""
    alpha = 0.3
    exp_diff = np.mean(np.abs(target_coeffs[:, 80:144] - source_coeff[:, 80:144]), 1)
    angle_diff = np.mean(np.abs(target_coeffs[:, 224:227] - source_coeff[:, 224:227]), 1)
    index = np.argmin(alpha * exp_diff + (1 - alpha) * angle_diff)
    return source_coeff[:, -3] / target_coeffs[index:index + 1, -3]


def synth_table(n_frames: int, seed: int = 0, dtype=np.float32, d: int = 262):
    """Seeded synthetic coefficient table [n_frames, 262] (id | exp | tex | angle | gamma | trans | crop params)."""
    rng = np.random.default_rng(seed)
    t = rng.standard_normal((n_frames, d)) * 0.5
    t[:, 257:262] = np.abs(t[:, 257:262]) * 100.0 + 50.0        # crop / alignment parameters are positive pixel-scale numbers
    return t.astype(dtype)
