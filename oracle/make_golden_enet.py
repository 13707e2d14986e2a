"""Generates tests/golden/enet_schema.json and enet_seed0_b1_out.npz (TEST INFRASTRUCTURE; build container only).

    python -m oracle.make_golden_enet

Imports the UNMODIFIED reference ENet (models/ENet.py wrapping models/LNet.py) through oracle/ref_shim.py, dumps its
state_dict schema, loads the seeded factory weights (oracle/enet.py make_state_dict, strict=True) and records its output on
seeded synthetic inputs with the StyleConv noise drawn under torch.manual_seed(7).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import enet, ref_shim, synth, weights


def reference_enet():
    LNet, _, _ = ref_shim.load()
    from models.ENet import ENet          # noqa: the reference's module (ref_shim put /root/reference on sys.path)
    return ENet(lnet=LNet()).eval()


def inputs(b=1, seed=0):
    mel, face = synth.lnet_inputs(b, seed=seed)
    g = torch.Generator().manual_seed(100 + seed)
    return mel, face, torch.rand(b, 3, 96, 96, generator=g)


def main():
    net = reference_enet()
    full = {k: list(v.shape) for k, v in net.state_dict().items()}
    lnet_schema = weights.load_schema("lnet")
    assert list(full)[:len(lnet_schema)] == ["low_res." + k for k in lnet_schema]       # low_res.* first, in LNet's own order
    schema = {k: v for k, v in full.items() if not k.startswith("low_res.")}             # only ENet's own tensors are stored
    with open(os.path.join(weights._GOLDEN, "enet_schema.json"), "w") as f:
        json.dump(schema, f, indent=0)
    net.load_state_dict(enet.make_state_dict(0), strict=True)
    mel, face, gt = inputs(1, 0)
    with torch.no_grad():
        torch.manual_seed(7)
        out, low = net(mel, face, gt)
    np.savez_compressed(os.path.join(weights._GOLDEN, "enet_seed0_b1_out.npz"), low=low.numpy(),
                        out_mean=np.array(float(out.double().mean())), out_absmax=np.array(float(out.abs().max())),
                        out_rows=out.numpy()[0, :, ::16].copy())           # every 16th of the 384 rows (the full image is 1.8 MB)
    print("enet own tensors", len(schema), "out", tuple(out.shape), float(out.min()), float(out.max()))


if __name__ == "__main__":
    main()
