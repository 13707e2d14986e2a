"""Drop-in for the 3DMM-coefficient helpers of the reference's futils/inference_utils.py (:73-99): the step in front
of DNet.forward that turns the per-frame coefficient table into DNet's ``driving_source`` windows.

Same names, argument meaning and return types as the reference (``transform_semantic`` returns a CPU float32 tensor
[73, 26]); ``semantic_windows`` is the batched form the pipeline uses (one launch for a whole frame range, result stays
in HBM).  The gather runs in libs2v's ``s2v_semantic_windows`` kernel - no CPU path, a CUDA device is required.
``find_crop_norm_ratio`` is a once-per-clip host reduction over the table (as in the reference: numpy).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L


def obtain_seq_index(index, num_frames):
    """inference_utils.py:73-76: the 26 table rows of frame ``index`` (index-13 .. index+12, clamped)."""
    seq = list(range(index - 13, index + 13))
    return [min(max(item, 0), num_frames - 1) for item in seq]


_EXP, _ANGLE = slice(80, 144), slice(224, 227)          # expression / pose columns of the coefficient table


def find_crop_norm_ratio(source_coeff, target_coeffs):
    """inference_utils.py:93-99 (host, once per clip): the source frame's crop scale divided by the crop scale of the
    target frame that is closest to it in expression (weight 0.3) and pose (weight 0.7), distances being mean absolute
    coefficient differences.  numpy in, numpy [1] out, dtype preserved (same operation order as the reference, so the
    argmin - an index, hence bit-exact - sees the same floats)."""
    w_exp = 0.3
    d_exp = np.abs(target_coeffs[:, _EXP] - source_coeff[:, _EXP]).mean(axis=1)
    d_pose = np.abs(target_coeffs[:, _ANGLE] - source_coeff[:, _ANGLE]).mean(axis=1)
    best = int(np.argmin(w_exp * d_exp + (1 - w_exp) * d_pose))
    return source_coeff[:, -3] / target_coeffs[best:best + 1, -3]


def _ratio_args(crop_norm_ratio):
    """The reference tests ``if crop_norm_ratio:`` (:87): None, 0 and [0.] leave the crop untouched."""
    if crop_norm_ratio is None:
        return 0.0, 0
    r = np.asarray(crop_norm_ratio.detach().cpu().numpy() if torch.is_tensor(crop_norm_ratio) else crop_norm_ratio)
    if r.size != 1:
        raise ValueError("crop_norm_ratio must hold one value (the truth value of a longer array is ambiguous)")
    r = float(r.reshape(-1)[0])
    return r, int(bool(r))


def upload_semantic(semantic, device) -> torch.Tensor:
    """The [T, D] coefficient table as a device tensor (float32 or float64 kept as given; D >= 262)."""
    t = semantic if torch.is_tensor(semantic) else torch.from_numpy(np.ascontiguousarray(semantic))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.float()
    if t.dim() != 2 or t.shape[1] < 262:
        raise ValueError("semantic must be [T, D >= 262], got %s" % (tuple(t.shape),))
    return t.to(device).contiguous()


def semantic_windows(semantic: torch.Tensor, frames, crop_norm_ratio=None) -> torch.Tensor:
    """semantic: CUDA table [T, D]; frames: (first, count) or an int sequence / tensor of frame indices.
    Returns CUDA float32 [count, 73, 26] - transform_semantic of every frame, stacked."""
    if not semantic.is_cuda:
        raise L.S2VError("semantic must be a CUDA tensor: this package has no CPU path (see upload_semantic)")
    lib = L.require_device(semantic.device.index)
    ratio, use = _ratio_args(crop_norm_ratio)
    if isinstance(frames, tuple) and len(frames) == 2:
        first, count, idx = int(frames[0]), int(frames[1]), None
    else:
        idx = torch.as_tensor(frames, dtype=torch.int32).to(semantic.device).contiguous()
        first, count = 0, int(idx.numel())
    out = torch.empty(count, 73, 26, dtype=torch.float32, device=semantic.device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    with torch.cuda.device(semantic.device):
        for s in range(0, count, 65535):
            n = min(65535, count - s)
            L.check(lib.s2v_semantic_windows(semantic.data_ptr(), int(semantic.dtype == torch.float64), semantic.shape[0],
                                             semantic.shape[1], None if idx is None else idx[s:].data_ptr(), first + s, n,
                                             ratio, use, out[s:].data_ptr(), stream), "s2v_semantic_windows")
    return out


def transform_semantic(semantic, frame_index, crop_norm_ratio=None):
    """inference_utils.py:78-91: numpy table [T, D] + frame index -> CPU float32 tensor [73, 26]."""
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise L.S2VError("a CUDA device is required: this package has no CPU path")
    table = semantic if torch.is_tensor(semantic) and semantic.is_cuda else upload_semantic(semantic, dev)
    return semantic_windows(table, (int(frame_index), 1), crop_norm_ratio)[0].cpu()


# ---- Laplacian-pyramid blend (inference_utils.py:181-222, called at inference.py:312) --------------------------------------
def laplacian_blend(A: torch.Tensor, B: torch.Tensor, m: torch.Tensor, num_levels: int = 6) -> torch.Tensor:
    """Batched form: A, B uint8 CUDA [N,H,W,C] (C = 1, 3 or 4, channels-last as cv2 holds images), m float32 CUDA [N,H,W]
    -> float32 CUDA [N,H,W,C].  The 8-bit Gaussian pyramids are bit-exact cv2.pyrDown; each level of the collapse is one
    fused launch (Laplacian levels of A and B, mask blend and reconstruction) - ~3 * num_levels launches per batch."""
    if not (A.is_cuda and B.is_cuda and m.is_cuda):
        raise L.S2VError("laplacian_blend needs CUDA tensors: this package has no CPU path")
    if A.dtype != torch.uint8 or B.dtype != torch.uint8:
        raise TypeError("A and B must be uint8 images (what inference.py:311-312 passes)")
    if A.dim() != 4 or A.shape != B.shape or tuple(m.shape) != tuple(A.shape[:3]) or A.shape[3] not in (1, 3, 4):
        raise ValueError("expected A, B [N,H,W,C] and m [N,H,W], got %s %s %s" % (tuple(A.shape), tuple(B.shape), tuple(m.shape)))
    if num_levels < 1:
        raise IndexError("num_levels must be >= 1")          # the reference indexes gp[num_levels - 1]
    n, h, w, c = A.shape
    div = 1 << (num_levels - 1)
    if h % div or w % div:
        # the reference fails in np.subtract(gp[i-1], cv2.pyrUp(gp[i])) when a level's size is odd
        raise ValueError("operands could not be broadcast together: H and W must be multiples of 2**(num_levels-1) = %d" % div)
    lib = L.require_device(A.device.index)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    gA, gB, gM = [A.contiguous()], [B.contiguous()], [m.contiguous().float()]
    with torch.cuda.device(A.device):
        for i in range(1, num_levels):            # the reference also builds level num_levels, which nothing reads
            hh, ww = gA[-1].shape[1], gA[-1].shape[2]
            oh, ow = (hh + 1) // 2, (ww + 1) // 2
            a = torch.empty(n, oh, ow, c, dtype=torch.uint8, device=A.device)
            b = torch.empty_like(a)
            mm = torch.empty(n, oh, ow, dtype=torch.float32, device=A.device)
            L.check(lib.s2v_pyrdown_u8(gA[-1].data_ptr(), n, hh, ww, c, a.data_ptr(), st), "s2v_pyrdown_u8")
            L.check(lib.s2v_pyrdown_u8(gB[-1].data_ptr(), n, hh, ww, c, b.data_ptr(), st), "s2v_pyrdown_u8")
            L.check(lib.s2v_pyrdown_f32(gM[-1].data_ptr(), n, hh, ww, 1, mm.data_ptr(), st), "s2v_pyrdown_f32")
            gA.append(a); gB.append(b); gM.append(mm)
        top = num_levels - 1
        out = torch.empty(gA[top].shape, dtype=torch.float32, device=A.device)
        L.check(lib.s2v_lap_blend_level(None, gA[top].data_ptr(), gB[top].data_ptr(), gM[top].data_ptr(), None, None,
                                        n, gA[top].shape[1], gA[top].shape[2], c, out.data_ptr(), st), "s2v_lap_blend_level")
        for i in range(top, 0, -1):
            fine = torch.empty(gA[i - 1].shape, dtype=torch.float32, device=A.device)
            L.check(lib.s2v_lap_blend_level(out.data_ptr(), gA[i - 1].data_ptr(), gB[i - 1].data_ptr(), gM[i - 1].data_ptr(),
                                            gA[i].data_ptr(), gB[i].data_ptr(), n, fine.shape[1], fine.shape[2], c,
                                            fine.data_ptr(), st), "s2v_lap_blend_level")
            out = fine
    return out


def Laplacian_Pyramid_Blending_with_mask(A, B, m, num_levels=6):
    """inference_utils.py:181-222, same signature: numpy uint8 images [H,W,3] + float32 mask [H,W] -> numpy float32 [H,W,3]."""
    if not torch.cuda.is_available():
        raise L.S2VError("a CUDA device is required: this package has no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device())
    a = torch.from_numpy(np.ascontiguousarray(A)).to(dev)[None]
    b = torch.from_numpy(np.ascontiguousarray(B)).to(dev)[None]
    mm = torch.from_numpy(np.ascontiguousarray(m, dtype=np.float32)).to(dev)[None]
    return laplacian_blend(a, b, mm, num_levels)[0].cpu().numpy()
