"""Batched GPU forms of the per-frame image glue the reference runs inline on the CPU (numpy / OpenCV), either side of the
networks.  Each function names the reference lines it replaces; uint8 images are channels-last [N,H,W,3] CUDA tensors exactly as
cv2 holds them (BGR stays BGR), network tensors float32 NCHW.  There is no CPU path.

    fake_to_bgr_u8      preprocessing/facing.py:190-192   DNet's fake_image -> the uint8 BGR frames of the stabilised video
    resize_u8           cv2.resize(x, (w, h))              bit-exact INTER_LINEAR for 8-bit images (inference.py:292, :308, :392-393)
    resize_f32          cv2.resize(x, (w, h))              float32 images / masks (inference.py:308, :313)
    face_batch          inference.py:388-399, :260-262     datagen's LNet / ENet input batch from the face and reference crops
    compose_pred_u8     inference.py:267, :282-290         the generated faces as uint8 images
    paste_faces         inference.py:292-297               resize every face to its box and paste it into a copy of its frame
    blend_paste_back    inference.py:308-313               resize to 512 x 512, Laplacian-pyramid blend, clip, resize back, uint8
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .futils import inference_utils


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.S2VError("CUDA tensors required: this package has no CPU path")


def fake_to_bgr_u8(fake: torch.Tensor) -> torch.Tensor:
    """fake [N,3,H,W] float32 -> uint8 [N,H,W,3] BGR: ``np.uint8((fake.clamp(-1,1) + 1) / 2. * 255)`` + cv2.COLOR_RGB2BGR."""
    _need_cuda(fake)
    fake = fake.contiguous().float()
    n, c, h, w = fake.shape
    assert c == 3
    out = torch.empty(n, h, w, 3, dtype=torch.uint8, device=fake.device)
    with torch.cuda.device(fake.device):
        L.check(L.require_device(fake.device.index).s2v_fake_to_bgr_u8(fake.data_ptr(), n, h, w, out.data_ptr(), _stream()), "s2v_fake_to_bgr_u8")
    return out


def resize_u8(x: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    """cv2.resize(x, (ow, oh)) of uint8 [N,H,W,C] (C = 1 or 3), bit-exact."""
    _need_cuda(x)
    x = x.contiguous()
    n, h, w, c = x.shape
    assert x.dtype == torch.uint8
    out = torch.empty(n, oh, ow, c, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.require_device(x.device.index).s2v_resize_linear_u8(x.data_ptr(), n, h, w, c, out.data_ptr(), oh * ow * c, ow * c, None, oh, ow, 0,
                                                                    _stream()), "s2v_resize_linear_u8")
    return out


def resize_f32(x: torch.Tensor, oh: int, ow: int, clip: bool = False, to_u8: bool = False) -> torch.Tensor:
    """cv2.resize(x, (ow, oh)) of float32 [N,H,W,C] (C = 1 or 3).  ``clip``: np.clip(x, 0, 255) first; ``to_u8``: np.uint8() after."""
    _need_cuda(x)
    x = x.contiguous().float()
    n, h, w, c = x.shape
    out = torch.empty(n, oh, ow, c, dtype=torch.uint8 if to_u8 else torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.require_device(x.device.index).s2v_resize_linear_f32(x.data_ptr(), n, h, w, c, out.data_ptr(), oh, ow, int(clip), int(to_u8),
                                                                     _stream()), "s2v_resize_linear_f32")
    return out


def face_batch(ofaces: torch.Tensor, faces: torch.Tensor):
    """ofaces / faces: uint8 [N,S,S,3] crops already at img_size (resize_u8) -> (img_batch [N,6,S,S], img_original [N,3,S,S]) float32:
    the frame's face with its lower half zeroed | the stabilised reference, / 255 (inference.py:394-399, :260-262)."""
    _need_cuda(ofaces, faces)
    ofaces, faces = ofaces.contiguous(), faces.contiguous()
    n, s, s2, c = ofaces.shape
    assert s == s2 and c == 3 and faces.shape == ofaces.shape and ofaces.dtype == faces.dtype == torch.uint8
    ib = torch.empty(n, 6, s, s, dtype=torch.float32, device=ofaces.device)
    io = torch.empty(n, 3, s, s, dtype=torch.float32, device=ofaces.device)
    with torch.cuda.device(ofaces.device):
        L.check(L.require_device(ofaces.device.index).s2v_face_batch(ofaces.data_ptr(), faces.data_ptr(), n, s, ib.data_ptr(), io.data_ptr(), _stream()),
                "s2v_face_batch")
    return ib, io


def compose_pred_u8(pred: torch.Tensor, img_batch: torch.Tensor = None, img_original: torch.Tensor = None) -> torch.Tensor:
    """pred [N,3,S,S] -> uint8 [N,S,S,3]: clamp(pred, 0, 1) (inference.py:267); with img_batch / img_original the pixels the mask
    left visible come from img_original (the `without_rl1` branch, :282-288); ``* 255`` and ``astype(np.uint8)`` (:290, :292)."""
    _need_cuda(pred, img_batch, img_original)
    pred = pred.contiguous().float()
    n, c, s, s2 = pred.shape
    assert c == 3 and s == s2
    compose = img_batch is not None
    if compose:
        img_batch, img_original = img_batch.contiguous().float(), img_original.contiguous().float()
        assert tuple(img_batch.shape) == (n, 6, s, s) and tuple(img_original.shape) == (n, 3, s, s)
    out = torch.empty(n, s, s, 3, dtype=torch.uint8, device=pred.device)
    with torch.cuda.device(pred.device):
        L.check(L.require_device(pred.device.index).s2v_compose_pred_u8(pred.data_ptr(), img_batch.data_ptr() if compose else None,
                                                                      img_original.data_ptr() if compose else None, n, s, int(compose), out.data_ptr(),
                                                                      _stream()), "s2v_compose_pred_u8")
    return out


def paste_faces(p_u8: torch.Tensor, frames: torch.Tensor, boxes) -> torch.Tensor:
    """p_u8 uint8 [N,S,S,3], frames uint8 [N,H,W,3], boxes [N][4] = (y1, y2, x1, x2) (a list / int tensor) -> copies of the frames
    with every face resized to its box and pasted (inference.py:292-297): one launch, the resize writes into the window."""
    _need_cuda(p_u8, frames)
    p_u8 = p_u8.contiguous()
    n, s, s2, c = p_u8.shape
    ff = frames.contiguous().clone()
    assert ff.shape[0] == n and ff.dtype == p_u8.dtype == torch.uint8 and c == 3
    b = torch.as_tensor(boxes, dtype=torch.int32).reshape(n, 4)
    bh, bw = (b[:, 1] - b[:, 0]), (b[:, 3] - b[:, 2])
    if n and (int(b[:, 0].min()) < 0 or int(b[:, 2].min()) < 0 or int(b[:, 1].max()) > ff.shape[1] or int(b[:, 3].max()) > ff.shape[2] or
              int(bh.min()) <= 0 or int(bw.min()) <= 0):
        raise ValueError("boxes must lie inside the frames")
    bd = b.to(ff.device)
    h, w = ff.shape[1], ff.shape[2]
    with torch.cuda.device(ff.device):
        L.check(L.require_device(ff.device.index).s2v_resize_linear_u8(p_u8.data_ptr(), n, s, s2, 3, ff.data_ptr(), h * w * 3, w * 3, bd.data_ptr(), 0, 0,
                                                                     int((bh * bw).max()) if n else 0, _stream()), "s2v_resize_linear_u8")
    return ff


def blend_paste_back(restored: torch.Tensor, ff: torch.Tensor, mask: torch.Tensor, num_levels: int = 10) -> torch.Tensor:
    """inference.py:308-313 for a batch: restored / ff uint8 [N,H,W,3], mask float32 [N,H,W,3] (the mouth mask inside the face box)
    -> uint8 [N,H,W,3]: the three cv2.resize to 512 x 512, Laplacian_Pyramid_Blending_with_mask(restored, ff, mask[...,0], 10),
    np.clip(0, 255), cv2.resize back to the frame size, np.uint8."""
    _need_cuda(restored, ff, mask)
    h, w = ff.shape[1], ff.shape[2]
    r5, f5 = resize_u8(restored, 512, 512), resize_u8(ff, 512, 512)
    m5 = resize_f32(mask, 512, 512)
    img = inference_utils.laplacian_blend(r5, f5, m5[..., 0].contiguous(), num_levels)
    return resize_f32(img, h, w, clip=True, to_u8=True)
