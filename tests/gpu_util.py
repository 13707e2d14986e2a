"""Helpers shared by the -m gpu parity tests (the CUDA path is always called through the C ABI)."""
import os

import torch

import s2v_b200  # noqa: F401
from s2v_b200 import _lib as L
from s2v_b200 import ops


def lib():
    return L.require_device(torch.cuda.current_device())


def nhwc(t):
    """NCHW float -> contiguous fp16 NHWC"""
    return t.permute(0, 2, 3, 1).contiguous().half()


def nchw(t):
    return t.permute(0, 3, 1, 2).float()


def report(name, got, ref):
    d = (got.float() - ref.float()).abs()
    scale = ref.float().abs().max().item() + 1e-12
    msg = "%-34s max_abs=%.3e  rel_to_max=%.3e  mean_abs=%.3e" % (name, d.max().item(), d.max().item() / scale, d.mean().item())
    print(msg)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write(msg + "\n")
    return d.max().item(), d.max().item() / scale


def psnr(got, ref, peak):
    mse = ((got.double() - ref.double()) ** 2).mean().item()
    import math
    return 10 * math.log10(peak * peak / max(mse, 1e-30))
