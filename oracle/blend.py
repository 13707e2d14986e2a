"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy) of the reference's Laplacian-pyramid blend.

Follows /root/reference futils/inference_utils.py:181-222 (Laplacian_Pyramid_Blending_with_mask, called at
inference.py:312 with 512 x 512 uint8 images, a float32 mask and num_levels = 10).  The arithmetic lives in a third-party
dependency, OpenCV (cv2.pyrDown / cv2.pyrUp / cv2.add; `opencv-python` is unpinned in requirements.txt, 4.13.0 in this
image); its published algorithm is restated here:
  pyrDown: 5x5 kernel [1 4 6 4 1] x [1 4 6 4 1] / 256 on every second pixel, BORDER_REFLECT_101, output (n+1)//2;
           8-bit images in integer arithmetic with (sum + 128) >> 8 rounding (bit-exact), float images in float32;
  pyrUp:   zero insertion + the same kernel x 4, i.e. even outputs (x[i-1] + 6 x[i] + x[i+1]) / 8 and odd outputs
           (x[i] + x[i+1]) / 2 per axis, index -1 mirrored to 1 and index n clamped to n-1, output 2n.
Pinned: tests/golden/blend_golden.npz holds outputs of the reference's own function (its source executed unmodified with the
real cv2, oracle/make_golden_blend.py) and of cv2.pyrDown / pyrUp on seeded inputs; this restatement must match them - the
8-bit pyramids bit-for-bit, float results within 1e-4 on the 0..255 scale (tests/test_oracle_blend.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np


def _refl101(p, n):
    if n == 1:
        return np.zeros_like(p)
    p = np.abs(p)                                   # -p for p < 0
    return np.where(p >= n, 2 * n - 2 - p, p)       # one reflection is enough for |offset| <= 2 and n >= 2


def _down_axis(x, axis, acc_dtype):
    n = x.shape[axis]
    o = (n + 1) // 2
    idx = 2 * np.arange(o)[:, None] + np.arange(-2, 3)[None, :]
    idx = _refl101(idx, n)
    t = np.take(x.astype(acc_dtype), idx, axis=axis)             # [..., o, 5, ...]
    t = np.moveaxis(t, axis + 1, -1)
    return t[..., 2] * 6 + (t[..., 1] + t[..., 3]) * 4 + t[..., 0] + t[..., 4]


def pyr_down(x):
    """cv2.pyrDown for uint8 (bit-exact) and float32 images [H, W] or [H, W, C]."""
    if x.dtype == np.uint8:
        s = _down_axis(_down_axis(x, 1, np.int64), 0, np.int64)
        return ((s + 128) >> 8).astype(np.uint8)
    s = _down_axis(_down_axis(x, 1, np.float32), 0, np.float32)
    return (s * np.float32(1.0 / 256.0)).astype(np.float32)


def _up_axis(x, axis):
    n = x.shape[axis]
    x = np.moveaxis(x, axis, 0)
    lo = np.concatenate([x[min(1, n - 1):min(1, n - 1) + 1], x[:-1]], 0)      # x[i-1], index -1 -> 1 (0 when n == 1)
    hi = np.concatenate([x[1:], x[n - 1:n]], 0)                               # x[i+1], index n -> n-1
    out = np.empty((2 * n,) + x.shape[1:], np.float32)
    out[0::2] = lo + x * 6 + hi
    out[1::2] = (x + hi) * 4
    return np.moveaxis(out, 0, axis)


def pyr_up(x):
    """cv2.pyrUp for float32 images (output exactly twice the input size)."""
    x = x.astype(np.float32)
    return (_up_axis(_up_axis(x, 1), 0) * np.float32(1.0 / 64.0)).astype(np.float32)


def laplacian_pyramid_blending_with_mask(A, B, m, num_levels=6):
    """inference_utils.py:181-222.  A, B: uint8 (or float32) [H, W, 3]; m float32 [H, W]; -> float32 [H, W, 3]."""
    gpA, gpB, gpM = [A], [B], [m]
    GA, GB, GM = A, B, m
    for _ in range(num_levels):                                    # :190-196 (the last level is computed but never used)
        GA, GB, GM = pyr_down(GA), pyr_down(GB), pyr_down(GM)
        gpA.append(np.float32(GA)); gpB.append(np.float32(GB)); gpM.append(np.float32(GM))
    lpA, lpB, gpMr = [gpA[num_levels - 1]], [gpB[num_levels - 1]], [gpM[num_levels - 1]]
    for i in range(num_levels - 1, 0, -1):                         # :202-209
        lpA.append(np.subtract(gpA[i - 1], pyr_up(gpA[i])))
        lpB.append(np.subtract(gpB[i - 1], pyr_up(gpB[i])))
        gpMr.append(gpM[i - 1])
    LS = []
    for la, lb, gm in zip(lpA, lpB, gpMr):                         # :212-216
        gm = gm[:, :, np.newaxis]
        LS.append(la * gm + lb * (1.0 - gm))
    ls_ = LS[0]
    for i in range(1, num_levels):                                 # :219-222
        ls_ = pyr_up(ls_) + LS[i]
    return ls_.astype(np.float32)


def synth_images(h, w, seed=0, n=None):
    """Seeded synthetic (A, B, mask): two smooth-ish uint8 images and a soft float32 mask in [0, 1]."""
    rng = np.random.default_rng(seed)
    shape = (h, w) if n is None else (n, h, w)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    base = (127 + 90 * np.sin(6.0 * xx + 3.0 * yy))[..., None]
    A = np.clip(base + rng.normal(0, 25, shape + (3,)), 0, 255).astype(np.uint8)
    B = np.clip(255 - base + rng.normal(0, 25, shape + (3,)), 0, 255).astype(np.uint8)
    m = np.clip(1.2 - 2.4 * np.hypot(xx - 0.5, yy - 0.55) + rng.normal(0, 0.02, shape), 0, 1).astype(np.float32)
    return A, B, m
