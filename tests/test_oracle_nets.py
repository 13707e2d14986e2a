"""Pins the torch restatement (oracle/nets.py) and the weight factory against
(1) the committed golden outputs produced by the real reference
(oracle/make_golden.py) - runs anywhere - and (2) the unmodified reference
imported from /root/reference when that tree exists (build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import nets, ref_shim, synth, weights

from conftest import GOLDEN

torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def lnet_sd():
    return weights.make_state_dict("lnet", 0)


@pytest.fixture(scope="module")
def dnet_sd():
    return weights.make_state_dict("dnet", 0)


def test_factory_schema_and_sigma(lnet_sd):
    schema = weights.load_schema("lnet")
    assert list(lnet_sd.keys()) == list(schema.keys()) and len(schema) == 1721
    for k, shp in schema.items():
        assert list(lnet_sd[k].shape) == shp, k
    # warmed spectral norm: u^T W v equals the top singular value
    p = "decoder.final.model.0"
    w = lnet_sd[p + ".weight_orig"].flatten(1)
    sigma = torch.dot(lnet_sd[p + ".weight_u"], w @ lnet_sd[p + ".weight_v"])
    top = torch.linalg.svdvals(w)[0]
    assert abs(float(sigma / top) - 1) < 1e-2
    again = weights.make_state_dict("lnet", 0)
    assert all(torch.equal(again[k], lnet_sd[k]) for k in schema)


def test_lnet_restatement_vs_golden(lnet_sd):
    mel, face = synth.lnet_inputs(2, seed=0)
    out = nets.lnet_forward(lnet_sd, mel, face).numpy()
    gold = np.load(os.path.join(GOLDEN, "lnet_seed0_b2_out.npy"))
    assert out.shape == gold.shape == (2, 3, 96, 96)
    assert np.abs(out - gold).max() < 2e-5
    assert 0.05 < gold.std() and gold.min() > 0 and gold.max() < 1     # not saturated / degenerate


def test_dnet_restatement_vs_golden(dnet_sd):
    src, coeff = synth.dnet_inputs(1, seed=0)
    o = nets.dnet_forward(dnet_sd, src, coeff)
    g = np.load(os.path.join(GOLDEN, "dnet_seed0_b1_out.npz"))
    assert np.abs(o["flow_field"].numpy() - g["flow_field"]).max() < 2e-4
    assert np.abs(o["warp_image"].numpy() - g["warp_image"].astype(np.float32)).max() < 2e-3
    assert np.abs(o["fake_image"].numpy() - g["fake_image"].astype(np.float32)).max() < 2e-3
    assert "fake_image" not in nets.dnet_forward(dnet_sd, src, coeff, stage="warp")


def test_warp_restatement_and_closed_form_vs_golden():
    s, fl = synth.warp_inputs(2, seed=0, c=3, hw=64, fhw=16)
    gold = np.load(os.path.join(GOLDEN, "warp_seed0_b2_64_16.npy"))
    a = nets.warp_image(s, nets.convert_flow_to_deformation(fl)).numpy()
    assert np.abs(a - gold).max() < 1e-6
    b = nets.warp_closed_form(s.double(), fl.double()).numpy()
    assert np.abs(b - gold).max() < 2e-5          # gold is an fp32 computation


def test_lnet_5d_form(lnet_sd):
    mel, face = synth.lnet_inputs(4, seed=5)
    flat = nets.lnet_forward(lnet_sd, mel, face)                     # [T*B] time-major, B=2,T=2
    a5 = torch.stack([mel[:2], mel[2:]], 1)                          # [B,T,1,80,16]
    f5 = torch.stack([face[:2], face[2:]], 2)                        # [B,6,T,96,96]
    out5 = nets.lnet_forward(lnet_sd, a5, f5)
    assert out5.shape == (2, 3, 2, 96, 96)
    assert torch.allclose(out5[:, :, 1], flat[2:], atol=1e-6)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_restatement_vs_imported_reference(lnet_sd, dnet_sd):
    LNet, DNet, flow_util = ref_shim.load()
    lnet, dnet = LNet().eval(), DNet().eval()
    lnet.load_state_dict(lnet_sd, strict=True)
    dnet.load_state_dict(dnet_sd, strict=True)
    mel, face = synth.lnet_inputs(1, seed=7)
    assert (lnet(mel, face) - nets.lnet_forward(lnet_sd, mel, face)).abs().max() < 2e-5
    src, coeff = synth.dnet_inputs(1, seed=7)
    r, o = dnet(src, coeff), nets.dnet_forward(dnet_sd, src, coeff)
    for k in ("flow_field", "warp_image", "fake_image"):
        assert (r[k] - o[k]).abs().max() < 2e-4, k
    s, fl = synth.warp_inputs(1, seed=2, hw=32, fhw=8)
    assert (flow_util.warp_image(s, flow_util.convert_flow_to_deformation(fl))
            - nets.warp_image(s, nets.convert_flow_to_deformation(fl))).abs().max() < 1e-6
