"""Deterministic seeded weight factory (TEST INFRASTRUCTURE).

The reference ships no checkpoints (.gitignore:3), and its raw random init is
numerically degenerate for LNet (spectral-norm u,v are random => sigma ~ -0.01,
SURVEY B.3).  Parity therefore runs on "identical random-init weights" made by
this factory: a ``state_dict`` with exactly the reference's keys and shapes
(``tests/golden/{lnet,dnet}_schema.json``, dumped from the instantiated reference
modules by oracle/make_golden.py), filled per key from a key-derived seed:

* conv / linear ``weight`` / ``weight_orig`` / ``bias``: U(-b, b), b = 1/sqrt(fan_in)
  (PyTorch's default init bound).
* spectral norm ``weight_u`` / ``weight_v``: 30 power iterations on
  ``weight_orig.flatten(1)`` from a seeded start, i.e. what a trained/"warmed"
  module holds (sigma = u^T W v ~ top singular value).
* LayerNorm2d / nn.LayerNorm / BatchNorm affine: weight U(0.8,1.2), bias U(-0.1,0.1).
* BatchNorm running stats: mean N(0, 0.05^2), var U(0.25, 0.45); counters = 10.

Same torch version + CPU generator => bit-identical tensors in the build
container and on the GPU box, which is what lets the committed golden outputs
(generated here by the real reference) check the CUDA path there.
"""
from __future__ import annotations

import json
import os
import zlib

import torch
import torch.nn.functional as F

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_schema(net: str) -> dict:
    with open(os.path.join(_GOLDEN, f"{net}_schema.json")) as f:
        return json.load(f)


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _is_norm_affine(key: str, shape) -> bool:
    # LayerNorm2d [C,1,1]; nn.LayerNorm / BatchNorm [C] under *.norm*, *.bn, conv_block.1, conv1.1
    if len(shape) == 3 and shape[1:] == [1, 1]:
        return True
    parts = key.split(".")
    owner = parts[-2]
    if owner in ("normx", "normy", "norm", "bn"):
        return True
    if len(shape) == 1 and owner == "1" and parts[-3] in ("conv_block", "conv1"):
        return True
    return False


def make_state_dict(net: str, seed: int = 0, dtype=torch.float32) -> dict:
    schema = load_schema(net)
    sd: dict[str, torch.Tensor] = {}
    for key, shape in schema.items():
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.tensor(10, dtype=torch.int64)
        elif leaf == "running_mean":
            sd[key] = torch.randn(shape, generator=g) * 0.05
        elif leaf == "running_var":
            sd[key] = _uniform(shape, 0.25, 0.45, g)
        elif leaf in ("weight_u", "weight_v"):
            continue                                   # filled after weight_orig
        elif _is_norm_affine(key, shape):
            sd[key] = _uniform(shape, 0.8, 1.2, g) if leaf == "weight" else _uniform(shape, -0.1, 0.1, g)
        elif leaf in ("weight", "weight_orig"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            b = fan_in ** -0.5
            sd[key] = _uniform(shape, -b, b, g)
        elif leaf == "bias":
            wkey = key[:-4] + "weight"
            wshape = schema.get(wkey) or schema.get(key[:-4] + "weight_orig")
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            b = fan_in ** -0.5
            sd[key] = _uniform(shape, -b, b, g)
        else:
            raise KeyError("no init rule for " + key)
    for key in schema:
        if key.endswith(".weight_orig"):
            p = key[: -len(".weight_orig")]
            w = sd[key].flatten(1).double()
            g = _gen(seed, p + ".weight_u")
            u = F.normalize(torch.randn(w.shape[0], generator=g, dtype=torch.float64), dim=0, eps=1e-12)
            v = F.normalize(torch.mv(w.t(), u), dim=0, eps=1e-12)
            for _ in range(30):
                v = F.normalize(torch.mv(w.t(), u), dim=0, eps=1e-12)
                u = F.normalize(torch.mv(w, v), dim=0, eps=1e-12)
            sd[p + ".weight_u"] = u.float()
            sd[p + ".weight_v"] = v.float()
    out = {}
    for key in schema:                                  # schema order == reference order
        t = sd[key]
        out[key] = t.to(dtype) if t.is_floating_point() else t
    return out
