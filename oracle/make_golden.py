"""Generates tests/golden (TEST INFRASTRUCTURE; run in the build container only).

    python -m oracle.make_golden

Imports the UNMODIFIED reference from /root/reference (oracle/ref_shim.py),
instantiates LNet / DNet, dumps their state_dict key/shape schema, loads the
factory weights (oracle/weights.py, strict=True) and records the reference's
own outputs on seeded synthetic inputs (oracle/synth.py).  The fixtures are what
lets the CUDA path be checked against the reference on the GPU box, where the
reference tree does not exist.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import mel as omel
from . import ref_shim, synth, weights

GOLDEN = weights._GOLDEN


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    LNet, DNet, flow_util = ref_shim.load()
    torch.manual_seed(0)
    lnet, dnet = LNet().eval(), DNet().eval()
    for name, net in (("lnet", lnet), ("dnet", dnet)):
        schema = {k: list(v.shape) for k, v in net.state_dict().items()}
        with open(os.path.join(GOLDEN, f"{name}_schema.json"), "w") as f:
            json.dump(schema, f, indent=0)
        print(name, "schema:", len(schema), "tensors")

    with torch.no_grad():
        # ---- LNet, factory seed 0, B=2 -------------------------------------------------
        lnet.load_state_dict(weights.make_state_dict("lnet", 0), strict=True)
        mel, face = synth.lnet_inputs(2, seed=0)
        out = lnet(mel, face)
        np.save(os.path.join(GOLDEN, "lnet_seed0_b2_out.npy"), out.numpy())
        print("lnet out", out.shape, float(out.min()), float(out.max()), float(out.mean()))
        # ---- DNet, factory seed 0, B=1 -------------------------------------------------
        dnet.load_state_dict(weights.make_state_dict("dnet", 0), strict=True)
        src, coeff = synth.dnet_inputs(1, seed=0)
        o = dnet(src, coeff)
        np.savez_compressed(os.path.join(GOLDEN, "dnet_seed0_b1_out.npz"),
                            flow_field=o["flow_field"].numpy(),
                            warp_image=o["warp_image"].numpy().astype(np.float16),
                            fake_image=o["fake_image"].numpy().astype(np.float16))
        print("dnet flow", float(o["flow_field"].abs().max()), "fake", float(o["fake_image"].min()),
              float(o["fake_image"].max()))
        # ---- flow_util, small ----------------------------------------------------------
        s, fl = synth.warp_inputs(2, seed=0, c=3, hw=64, fhw=16)
        w = flow_util.warp_image(s, flow_util.convert_flow_to_deformation(fl))
        np.save(os.path.join(GOLDEN, "warp_seed0_b2_64_16.npy"), w.numpy())
    # ---- mel (oracle restatement; librosa absent => not a reference output) -------------
    wav = synth.wav(1.0, seed=0)
    m = omel.melspectrogram(wav)
    np.save(os.path.join(GOLDEN, "mel_oracle_seed0_1s.npy"), m.astype(np.float32))
    counts = {str(sec): len(omel.mel_window_starts(1 + int(sec * 16000) // 200)) for sec in (5, 60, 600)}
    with open(os.path.join(GOLDEN, "mel_window_counts.json"), "w") as f:
        json.dump({"counts": counts, "starts_5s": omel.mel_window_starts(401)}, f)
    print("window counts", counts)


if __name__ == "__main__":
    main()
