"""TEST INFRASTRUCTURE ONLY - functional PyTorch restatement of the reference's ENet (the 96 -> 384
upsampler that wraps LNet and is the object inference.py:266 really calls).  The CUDA path (s2v_b200.models.ENet) is checked
against this module and against the golden output of the unmodified reference in tests/test_gpu_enet.py.

Follows /root/reference models/ENet.py:82-139 (forward) and models/base_blocks.py:29-49 (ResBlock), :460-508
(ModulatedConv2d), :515-536 (StyleConv), :539-553 (ToRGB).  Pinned: tests/test_oracle_enet.py compares it with the unmodified
reference imported from /root/reference (when present) and with tests/golden/enet_seed0_b1_out.npz produced by that
reference (oracle/make_golden_enet.py).  The reference draws the StyleConv noise from the global RNG inside forward
(base_blocks.py:528-530); both sides draw it the same way, in the same order, under torch.manual_seed(noise_seed), and the
restatement also accepts explicit noise tensors (the hook the CUDA path will use).
Only tests/ may import this module.
"""
from __future__ import annotations

import json
import math
import os

import torch
import torch.nn.functional as F

from . import nets, weights


def load_schema() -> dict:
    """Full reference key -> shape map: ``low_res.*`` (LNet's schema, registered first) followed by ENet's own 64 tensors."""
    with open(os.path.join(weights._GOLDEN, "enet_schema.json")) as f:
        own = json.load(f)
    full = {"low_res." + k: v for k, v in weights.load_schema("lnet").items()}
    full.update(own)
    return full


def make_state_dict(seed: int = 0) -> dict:
    """Reference-schema ENet state_dict: ``low_res.*`` = the LNet factory weights; ENet's own 64 tensors seeded per key:
    conv / linear U(+-1/sqrt(fan_in)); modulated-conv weight N(0,1)/sqrt(Cin k^2) (base_blocks.py:483-485); modulation bias 1
    (:481); StyleConv noise strength 0.05 (non-zero so that the noise path is exercised; the reference initialises 0)."""
    schema = load_schema()
    lnet = weights.make_state_dict("lnet", seed)
    sd = {}
    for key, shape in schema.items():
        if key.startswith("low_res."):
            sd[key] = lnet[key[len("low_res."):]]
            continue
        g = weights._gen(seed, "enet." + key)
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("modulated_conv.weight"):
            _, co, ci, kh, kw = shape
            sd[key] = torch.randn(shape, generator=g) / math.sqrt(ci * kh * kw)
        elif key.endswith("modulation.bias"):
            sd[key] = torch.ones(shape)
        elif leaf == "weight" and shape == [1]:
            sd[key] = torch.full(shape, 0.05)
        elif leaf == "bias" and len(shape) == 4:
            sd[key] = weights._uniform(shape, -0.1, 0.1, g)
        elif leaf == "weight":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[key] = weights._uniform(shape, -fan_in ** -0.5, fan_in ** -0.5, g)
        elif leaf == "bias":
            wshape = schema[key[:-4] + "weight"]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            sd[key] = weights._uniform(shape, -fan_in ** -0.5, fan_in ** -0.5, g)
        else:
            raise KeyError("no init rule for " + key)
    return sd


def res_block_down(x, sd, p):                        # base_blocks.py:29-49, mode='down'
    out = F.leaky_relu(F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1), 0.2)
    out = F.interpolate(out, scale_factor=0.5, mode="bilinear", align_corners=False)
    out = F.leaky_relu(F.conv2d(out, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1), 0.2)
    skip = F.conv2d(F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False), sd[p + ".skip.weight"])
    return out + skip


def modulated_conv(x, style, sd, p, demodulate, upsample, eps=1e-8):   # base_blocks.py:487-508
    b, c, _, _ = x.shape
    w = sd[p + ".weight"]                                               # [1, Cout, Cin, k, k]
    co, k = w.shape[1], w.shape[3]
    s = F.linear(style, sd[p + ".modulation.weight"], sd[p + ".modulation.bias"]).view(b, 1, c, 1, 1)
    w = w * s
    if demodulate:
        w = w * torch.rsqrt(w.pow(2).sum([2, 3, 4]) + eps).view(b, co, 1, 1, 1)
    if upsample:
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    h, wd = x.shape[2:]
    out = F.conv2d(x.reshape(1, b * c, h, wd), w.view(b * co, c, k, k), padding=k // 2, groups=b)
    return out.view(b, co, h, wd)


def style_conv(x, style, sd, p, upsample, noise=None):                 # base_blocks.py:515-536
    out = modulated_conv(x, style, sd, p + ".modulated_conv", True, upsample) * 2 ** 0.5
    if noise is None:
        b, _, h, w = out.shape
        noise = out.new_empty(b, 1, h, w).normal_()                     # the reference's draw (:528-530), global RNG
    out = out + sd[p + ".weight"] * noise + sd[p + ".bias"]
    return F.leaky_relu(out, 0.2)


def to_rgb(x, style, skip, sd, p):                                      # base_blocks.py:539-553
    out = modulated_conv(x, style, sd, p + ".modulated_conv", False, False) + sd[p + ".bias"]
    return out + F.interpolate(skip, scale_factor=2, mode="bilinear", align_corners=False)


def enet_forward(sd, audio, face, gt, noises=None):
    """models/ENet.py:82-139, 4-D form: audio [B,1,80,16], face [B,6,96,96], gt [B,3,96,96] -> ([B,3,384,384], [B,3,96,96]).
    ``noises``: optional list of 4 tensors [B,1,h,w] (200, 200, 400, 400) for the four StyleConvs, else drawn like the reference."""
    inp, ref = torch.split(face, 3, dim=1)
    feat = F.leaky_relu(F.conv2d(F.interpolate(ref, size=(256, 256), mode="bilinear"), sd["conv_body_first.weight"],
                                 sd["conv_body_first.bias"]), 0.2)
    for i in range(6):
        feat = res_block_down(feat, sd, f"conv_body_down.{i}")
    feat = F.leaky_relu(F.conv2d(feat, sd["final_conv.weight"], sd["final_conv.bias"], padding=1), 0.2)
    style = F.linear(feat.reshape(feat.size(0), -1), sd["final_linear.weight"], sd["final_linear.bias"])
    style = style.reshape(style.size(0), -1, 512)                       # [B, 1, 512]
    lnet_in = F.interpolate(torch.cat([inp, gt], dim=1), size=(96, 96), mode="bilinear")
    lsd = {k[len("low_res."):]: v for k, v in sd.items() if k.startswith("low_res.")}
    low = nets.lnet_forward(lsd, audio, lnet_in)
    out = F.pad(low, (2, 2, 2, 2), "reflect")
    skip = out
    for j in range(2):
        n0 = None if noises is None else noises[2 * j]
        n1 = None if noises is None else noises[2 * j + 1]
        out = style_conv(out, style, sd, f"style_convs.{2 * j}", True, n0)
        out = style_conv(out, style, sd, f"style_convs.{2 * j + 1}", False, n1)
        skip = to_rgb(out, style, skip, sd, f"to_rgbs.{j}")
    return skip[:, :, 8:-8, 8:-8], low
