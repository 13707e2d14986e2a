"""Generates tests/golden/imageops_golden.npz (TEST INFRASTRUCTURE; run in the build container only).

    python -m oracle.make_golden_imageops

The per-frame image glue of the reference is inline code (preprocessing/facing.py:190-192, inference.py:260-262, :267, :282-297,
:308-313, :388-399), so the generator executes those LINES literally - same numpy / torch / cv2 calls, same order, with the real
``cv2`` of this image - on seeded inputs, and the reference's own Laplacian_Pyramid_Blending_with_mask (cut out of the unmodified
file, oracle/make_golden_blend.py).  cv2.resize itself is recorded on edge shapes (1-pixel sources, exact 2 x down-scale, up-scale).
"""
from __future__ import annotations

import os

import cv2
import numpy as np
import torch

from . import make_golden_blend, weights


def main():
    rng = np.random.default_rng(11)
    out = {"cv2_version": np.array(cv2.__version__)}
    # ---- cv2.resize (INTER_LINEAR default), uint8 3-channel and float32 1- / 3-channel ------------------------------------------
    shapes = [((17, 23), (40, 31)), ((40, 31), (17, 23)), ((32, 48), (16, 24)), ((1, 5), (4, 9)), ((6, 1), (3, 7)), ((24, 24), (96, 96)),
              ((50, 70), (49, 71)), ((9, 9), (1, 1))]
    for i, ((h, w), (oh, ow)) in enumerate(shapes):
        x = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        f = (rng.random((h, w, 3), dtype=np.float32) * 300 - 20)
        out[f"rs_u8_{i}_in"], out[f"rs_u8_{i}"] = x, cv2.resize(x, (ow, oh)).reshape(oh, ow, 3)
        out[f"rs_f32_{i}_in"], out[f"rs_f32_{i}"] = f, cv2.resize(f, (ow, oh)).reshape(oh, ow, 3)
        out[f"rs_f32c1_{i}"] = cv2.resize(f[:, :, 0], (ow, oh)).reshape(oh, ow)
    # ---- facing.py:190-192 ----------------------------------------------------------------------------------------------------
    fake = torch.from_numpy((rng.standard_normal((1, 3, 24, 40)) * 0.8).astype(np.float32))
    out["fake_in"] = fake.numpy()
    img_stablized = np.uint8((fake.clone().squeeze(0).permute(1, 2, 0).cpu().clamp_(-1, 1).numpy() + 1) / 2. * 255)
    out["fake_bgr"] = cv2.cvtColor(img_stablized, cv2.COLOR_RGB2BGR)
    # ---- datagen's batch, inference.py:388-399 and :260-262 (img_size = 48 here; the reference's default is 384) ---------------------
    img_size = 48
    ofaces = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((61, 55), (30, 37), (96, 96))]
    faces = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((61, 55), (30, 37), (96, 96))]
    img_batch, ref_batch = [], []
    for oface, face in zip(ofaces, faces):
        face = cv2.resize(face, (img_size, img_size))
        oface = cv2.resize(oface, (img_size, img_size))
        img_batch.append(oface)
        ref_batch.append(face)
    img_batch, ref_batch = np.asarray(img_batch), np.asarray(ref_batch)
    img_masked = img_batch.copy()
    img_original = img_batch.copy()
    img_masked[:, img_size // 2:] = 0
    img_batch = np.concatenate((img_masked, ref_batch), axis=3) / 255.
    t_img_batch = torch.FloatTensor(np.transpose(img_batch, (0, 3, 1, 2)))
    t_img_original = torch.FloatTensor(np.transpose(img_original, (0, 3, 1, 2))) / 255.
    for i in range(3):
        out[f"oface_{i}"], out[f"face_{i}"] = ofaces[i], faces[i]
    out["img_batch"], out["img_original"] = t_img_batch.numpy(), t_img_original.numpy()
    # ---- inference.py:267, :282-288 -----------------------------------------------------------------------------------------
    pred = torch.from_numpy((rng.random((3, 3, img_size, img_size), dtype=np.float32) * 1.4 - 0.2))
    out["pred_in"] = pred.numpy()
    pred = torch.clamp(pred, 0, 1)
    incomplete, reference = torch.split(t_img_batch, 3, dim=1)
    mask = torch.where(incomplete == 0, torch.ones_like(incomplete), torch.zeros_like(incomplete))
    pred_c = pred * mask + t_img_original * (1 - mask)
    out["pred_u8_composed"] = (pred_c.cpu().numpy().transpose(0, 2, 3, 1) * 255.).astype(np.uint8)
    out["pred_u8_plain"] = (pred.cpu().numpy().transpose(0, 2, 3, 1) * 255.).astype(np.uint8)
    # ---- inference.py:292-297 and :308-313 ------------------------------------------------------------------------------------
    blend_fn = make_golden_blend.load_reference_function()
    xf = rng.integers(0, 256, (90, 120, 3), dtype=np.uint8)
    c = (20, 71, 33, 95)
    p = out["pred_u8_composed"][0]
    y1, y2, x1, x2 = c
    p = cv2.resize(p.astype(np.uint8), (x2 - x1, y2 - y1))
    ff = xf.copy()
    ff[y1:y2, x1:x2] = p
    out["frame_in"], out["box"], out["frame_pasted"] = xf, np.array(c), ff
    restored_img = np.clip(ff.astype(np.int32) + rng.integers(-20, 20, ff.shape), 0, 255).astype(np.uint8)      # stands in for GFPGAN's output
    mouse_mask = np.zeros_like(restored_img)
    tmp_mask = rng.integers(0, 2, (256, 256), dtype=np.uint8) * 255                                             # stands in for the face parser's map
    tmp_mask = cv2.GaussianBlur(tmp_mask, (31, 31), 7)
    mouse_mask[y1:y2, x1:x2] = cv2.resize(tmp_mask, (x2 - x1, y2 - y1))[:, :, np.newaxis] / 255.
    out["restored_in"], out["mouse_mask_in"] = restored_img, np.float32(mouse_mask)
    height, width = ff.shape[:2]
    restored_img, ff, full_mask = [cv2.resize(x, (512, 512)) for x in (restored_img, ff, np.float32(mouse_mask))]
    img = blend_fn(restored_img, ff, full_mask[:, :, 0], 10)
    pp = np.uint8(cv2.resize(np.clip(img, 0, 255), (width, height)))
    out["blend_back"] = pp
    np.savez_compressed(os.path.join(weights._GOLDEN, "imageops_golden.npz"), **out)
    print("written", len(out), "arrays")


if __name__ == "__main__":
    main()
