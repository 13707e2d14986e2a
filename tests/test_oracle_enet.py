"""CPU: SURVEY 8f #1 (ENet): the functional restatement (oracle/enet.py) against the golden output of the unmodified
reference ENet (tests/golden/enet_seed0_b1_out.npz, oracle/make_golden_enet.py) and, when /root/reference is present,
against the imported reference itself.  The CUDA ENet is checked against both in tests/test_gpu_enet.py."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import enet, make_golden_enet, ref_shim


@pytest.fixture(scope="module")
def sd():
    return enet.make_state_dict(0)


def test_schema_and_factory(sd):
    schema = enet.load_schema()
    assert len(schema) == 1785 and sum(1 for k in schema if not k.startswith("low_res.")) == 64
    assert list(sd) == list(schema) and all(list(sd[k].shape) == schema[k] for k in schema)
    assert float(sd["style_convs.0.weight"]) == pytest.approx(0.05) and bool((sd["to_rgbs.1.modulated_conv.modulation.bias"] == 1).all())


def test_restatement_vs_golden(sd):
    gold = np.load(os.path.join(GOLDEN, "enet_seed0_b1_out.npz"))
    mel, face, gt = make_golden_enet.inputs(1, 0)
    with torch.no_grad():
        torch.manual_seed(7)                       # the StyleConv noise is drawn from the global RNG, like the reference
        out, low = enet.enet_forward(sd, mel, face, gt)
    assert out.shape == (1, 3, 384, 384) and low.shape == (1, 3, 96, 96)
    assert np.abs(low.numpy() - gold["low"]).max() < 2e-5
    assert np.abs(out.numpy()[0, :, ::16] - gold["out_rows"]).max() < 1e-4 * max(1.0, float(gold["out_absmax"]))
    assert abs(float(out.double().mean()) - float(gold["out_mean"])) < 1e-5
    # explicit noise (the hook a CUDA path needs) == the same draws made by hand, in the reference's order
    torch.manual_seed(7)
    noises = [torch.empty(1, 1, s, s).normal_() for s in (200, 200, 400, 400)]
    with torch.no_grad():
        out2, _ = enet.enet_forward(sd, mel, face, gt, noises=noises)
    assert torch.equal(out, out2)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_restatement_vs_imported_reference(sd):
    net = make_golden_enet.reference_enet()
    net.load_state_dict(sd, strict=True)
    mel, face, gt = make_golden_enet.inputs(1, 3)
    with torch.no_grad():
        torch.manual_seed(11)
        ref_out, ref_low = net(mel, face, gt)
        torch.manual_seed(11)
        out, low = enet.enet_forward(sd, mel, face, gt)
    assert (out - ref_out).abs().max().item() < 1e-4 and (low - ref_low).abs().max().item() < 2e-5
