"""Drop-in for the reference's futils/flow_util.py, backed by libs2v's fused warp kernel.

Same function names, argument meaning and shapes (flow_util.py:3-15, :17-38, :41-56).
CUDA float32 tensors only (no CPU path).
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib as L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prep(t, name):
    if not t.is_cuda:
        raise L.S2VError("%s must be a CUDA tensor: this package has no CPU path" % name)
    return t.contiguous().float()


def convert_flow_to_deformation(flow):
    """flow [B,2,h,w] -> deformation [B,h,w,2] (flow_util.py:3-15)."""
    flow = _prep(flow, "flow")
    b, c, h, w = flow.shape
    lib = L.require_device(flow.device.index)
    out = torch.empty(b, h, w, 2, dtype=torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        L.check(lib.s2v_flow_to_deformation_f32(flow.data_ptr(), out.data_ptr(), b, h, w, _stream()), "s2v_flow_to_deformation_f32")
    return out


def make_coordinate_grid(flow):
    """grid [B,h,w,2] in [-1,1] with the same size as the flow field (flow_util.py:17-38)."""
    return convert_flow_to_deformation(torch.zeros_like(flow))


def warp_image(source_image, deformation):
    """source [B,C,H,W], deformation [B,h,w,2] -> [B,C,H,W] (flow_util.py:41-56:
    bilinear grid resize when sizes differ, then grid_sample bilinear/zeros/align_corners=False)."""
    src = _prep(source_image, "source_image")
    d = _prep(deformation, "deformation")
    b, c, h, w = src.shape
    _, hd, wd, _ = d.shape
    lib = L.require_device(src.device.index)
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        L.check(lib.s2v_warp_deformation_f32(src.data_ptr(), d.data_ptr(), out.data_ptr(), b, c, h, w, hd, wd, _stream()),
                "s2v_warp_deformation_f32")
    return out


def warp_flow(source_image, flow):
    """Fused convert_flow_to_deformation + warp_image (what DNet.forward does, models/DNet.py:88-89)."""
    src = _prep(source_image, "source_image")
    fl = _prep(flow, "flow")
    b, c, h, w = src.shape
    lib = L.require_device(src.device.index)
    out = torch.empty_like(src)
    with torch.cuda.device(src.device):
        L.check(lib.s2v_flow_warp_f32(src.data_ptr(), fl.data_ptr(), out.data_ptr(), b, c, h, w, fl.shape[2], fl.shape[3],
                                      None, 0, _stream()), "s2v_flow_warp_f32")
    return out
