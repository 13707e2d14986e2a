"""numpy restatement of the per-frame image glue around the networks (TEST INFRASTRUCTURE).

The reference does these steps inline, per frame, on the CPU with numpy / OpenCV (`opencv-python`, un-pinned in requirements.txt;
4.13.0 in this image):

* ``fake_to_bgr_u8``      preprocessing/facing.py:190-192   DNet's fake_image -> clamp(-1,1) -> uint8 -> RGB2BGR
* ``face_batch``          inference.py:388-399, :260-262     datagen's batch: cv2.resize of the face / reference crops to img_size,
                                                             lower half masked, concatenated, / 255 -> [B,6,S,S] and img_original
* ``compose_pred_u8``     inference.py:267, :282-288, :292   clamp(pred,0,1), the `without_rl1` mask mix, * 255, astype(uint8)
* ``paste_resized``       inference.py:292-297               cv2.resize(p, (x2-x1, y2-y1)); ff = xf.copy(); ff[y1:y2, x1:x2] = p
* ``blend_paste_back``    inference.py:308-313               the three cv2.resize to 512 x 512, the Laplacian blend, clip, resize back, uint8
* ``resize_linear_u8`` / ``resize_linear_f32``               cv2.resize(..., interpolation=INTER_LINEAR) - OpenCV's published algorithm
  (modules/imgproc/src/resize.cpp): source coordinate fx = (dx + 0.5) * scale - 0.5 in float, taps clamped at the borders;
  8-bit images use 11-bit fixed-point weights (cvRound(w * 2048)) horizontally and
  ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`` vertically; an exact 2 x 2 down-scale is INTER_AREA
  ((a + b + c + d + 2) >> 2).  float32 images: the same taps in float; OpenCV's own code uses float coordinates, its IPP path
  (the default in the pip wheels) double ones - ``coords`` selects, default "f64" = what ``cv2.resize`` returns here.

Pinned: oracle/make_golden_imageops.py runs the real cv2 on seeded inputs (tests/golden/imageops_golden.npz); the 8-bit resize must
match bit for bit, the float32 one to 1e-4 on the 0..255 scale.
"""
from __future__ import annotations

import numpy as np


def _coords(n_src: int, n_dst: int, f64: bool):
    scale = n_src / n_dst
    d = np.arange(n_dst, dtype=np.float64)
    f = (d + 0.5) * scale - 0.5
    if not f64:
        f = f.astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(f.dtype)
    return s, f


def _clamp_x(sx, fx, w):
    lo = sx < 0
    fx, sx = np.where(lo, 0, fx), np.where(lo, 0, sx)
    hi = sx >= w - 1
    return np.where(hi, w - 1, sx), np.where(hi, 0, fx).astype(fx.dtype)


def resize_linear_u8(src: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """cv2.resize(src, (ow, oh)) for uint8 [H,W,C] (or [H,W]); bit-exact."""
    sq = src.ndim == 2
    s = src[:, :, None] if sq else src
    h, w, _ = s.shape
    s32 = s.astype(np.int32)
    if h == 2 * oh and w == 2 * ow:
        out = ((s32[0::2, 0::2] + s32[0::2, 1::2] + s32[1::2, 0::2] + s32[1::2, 1::2] + 2) >> 2).astype(np.uint8)
        return out[:, :, 0] if sq else out
    sx, fx = _clamp_x(*_coords(w, ow, False), w)
    fx = fx.astype(np.float32)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int32)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int32)
    sx1 = np.minimum(sx + 1, w - 1)
    hz = s32[:, sx] * a0[None, :, None] + s32[:, sx1] * a1[None, :, None]
    sy, fy = _coords(h, oh, False)
    fy = fy.astype(np.float32)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int32)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int32)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    out = (((b0[:, None, None] * (hz[y0] >> 4)) >> 16) + ((b1[:, None, None] * (hz[y1] >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if sq else out


def resize_linear_f32(src: np.ndarray, oh: int, ow: int, coords: str = "f64") -> np.ndarray:
    """cv2.resize(src, (ow, oh)) for float32 [H,W,C] (or [H,W])."""
    sq = src.ndim == 2
    s = (src[:, :, None] if sq else src).astype(np.float32)
    h, w, _ = s.shape
    if h == 2 * oh and w == 2 * ow:
        out = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2]) * np.float32(0.25)
        return out[:, :, 0] if sq else out
    f64 = coords == "f64"
    acc = np.float64 if f64 else np.float32
    sx, fx = _clamp_x(*_coords(w, ow, f64), w)
    sx1 = np.minimum(sx + 1, w - 1)
    sa = s.astype(acc)
    fx = fx.astype(acc)
    hz = sa[:, sx] * (1 - fx)[None, :, None] + sa[:, sx1] * fx[None, :, None]
    sy, fy = _coords(h, oh, f64)
    fy = fy.astype(acc)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    out = (hz[y0] * (1 - fy)[:, None, None] + hz[y1] * fy[:, None, None]).astype(np.float32)
    return out[:, :, 0] if sq else out


def fake_to_bgr_u8(fake: np.ndarray) -> np.ndarray:
    """facing.py:190-192: fake [3,H,W] float32 in (-1,1) -> uint8 BGR [H,W,3]."""
    x = np.clip(fake.astype(np.float32), -1, 1).transpose(1, 2, 0)
    rgb = np.uint8((x + 1) / 2. * 255)
    return np.ascontiguousarray(rgb[:, :, ::-1])


def face_batch(ofaces, faces, img_size: int = 384):
    """inference.py:388-399 + :260-262: lists of uint8 [h,w,3] crops (the frame's face ``oface`` and the stabilised reference
    ``face``) -> img_batch float32 [B,6,S,S] (masked face | reference, / 255) and img_original float32 [B,3,S,S] (/ 255)."""
    of = np.asarray([resize_linear_u8(o, img_size, img_size) for o in ofaces])
    fa = np.asarray([resize_linear_u8(f, img_size, img_size) for f in faces])
    masked = of.copy()
    masked[:, img_size // 2:] = 0
    img_batch = np.concatenate((masked, fa), axis=3) / 255.                       # float64, like the reference
    img_batch = np.transpose(img_batch, (0, 3, 1, 2)).astype(np.float32)          # torch.FloatTensor(...)
    img_original = np.transpose(of, (0, 3, 1, 2)).astype(np.float32) / np.float32(255.)   # FloatTensor(...) / 255.
    return img_batch, img_original


def compose_pred_u8(pred: np.ndarray, img_batch: np.ndarray, img_original: np.ndarray, compose: bool = True) -> np.ndarray:
    """inference.py:267, :282-288: pred [B,3,S,S] float32 -> uint8 [B,S,S,3]; with ``compose`` the pixels the mask left visible
    (incomplete != 0) come from img_original."""
    p = np.clip(pred.astype(np.float32), 0, 1)
    if compose:
        inc = img_batch[:, :3]
        mask = np.where(inc == 0, np.float32(1), np.float32(0))
        p = p * mask + img_original.astype(np.float32) * (1 - mask)
    return (p.transpose(0, 2, 3, 1) * np.float32(255.)).astype(np.uint8)


def paste_resized(p_u8: np.ndarray, frame: np.ndarray, box) -> np.ndarray:
    """inference.py:292-297: p uint8 [S,S,3], frame uint8 [H,W,3], box (y1, y2, x1, x2) -> the frame copy with the resized face."""
    y1, y2, x1, x2 = box
    ff = frame.copy()
    ff[y1:y2, x1:x2] = resize_linear_u8(p_u8, y2 - y1, x2 - x1)
    return ff


def blend_paste_back(restored: np.ndarray, ff: np.ndarray, mask: np.ndarray, num_levels: int = 10, coords: str = "f64") -> np.ndarray:
    """inference.py:308-313: restored / ff uint8 [H,W,3], mask float32 [H,W,3] -> uint8 [H,W,3]."""
    from . import blend
    h, w = ff.shape[:2]
    r5, f5 = resize_linear_u8(restored, 512, 512), resize_linear_u8(ff, 512, 512)
    m5 = resize_linear_f32(np.float32(mask), 512, 512, coords)
    img = blend.laplacian_pyramid_blending_with_mask(r5, f5, m5[:, :, 0], num_levels)
    return np.uint8(resize_linear_f32(np.clip(img, 0, 255).astype(np.float32), h, w, coords))
