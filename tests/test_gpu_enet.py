"""-m gpu parity of the CUDA ENet (the 96 -> 384 upsampler wrapping LNet, models/ENet.py:82-139) through the drop-in
module, against the golden output of the UNMODIFIED reference (tests/golden/enet_seed0_b1_out.npz) and the oracle restatement.
The StyleConv noise is passed explicitly (the hook): the same draws the reference made under torch.manual_seed(7)."""
import math
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def env():
    import gpu_util
    from oracle import enet as oenet
    from s2v_b200.models.ENet import ENet
    from s2v_b200.models.LNet import LNet
    gpu_util.lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = oenet.make_state_dict(0)
    net = ENet(lnet=LNet()).cuda().eval()
    net.load_state_dict(sd, strict=True)
    return gpu_util, oenet, sd, net


def _psnr(a, b, peak):
    return 10 * math.log10(peak * peak / max(((a.double() - b.double()) ** 2).mean().item(), 1e-30))


def test_enet_state_dict_schema(env):
    G, oenet, sd, net = env
    out = net.state_dict()
    assert list(out.keys()) == list(sd.keys()) and len(out) == 1785
    assert all(torch.equal(out[k].cpu(), sd[k]) for k in sd)


def test_enet_vs_reference_golden(env):
    G, oenet, sd, net = env
    from oracle import make_golden_enet
    gold = np.load(os.path.join(GOLDEN, "enet_seed0_b1_out.npz"))
    mel, face, gt = make_golden_enet.inputs(1, 0)
    torch.manual_seed(7)
    noises = [torch.empty(1, 1, s, s).normal_().cuda() for s in (200, 200, 400, 400)]      # the reference's CPU draws
    out, low = net(mel.cuda(), face.cuda(), gt.cuda(), noises=noises)
    assert out.shape == (1, 3, 384, 384) and low.shape == (1, 3, 96, 96) and out.dtype == torch.float32
    peak = float(gold["out_absmax"])
    p_low = _psnr(low.cpu(), torch.from_numpy(gold["low"]), 1.0)
    p_out = _psnr(out.cpu()[0, :, ::16], torch.from_numpy(gold["out_rows"]), peak)
    m = (out.cpu()[0, :, ::16] - torch.from_numpy(gold["out_rows"])).abs().max().item()
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write("ENet vs reference golden B=1: out PSNR %.2f dB (peak %.3f) max_abs %.5f, low PSNR %.2f dB, mean %.6f vs %.6f\n"
                % (p_out, peak, m, p_low, float(out.double().mean()), float(gold["out_mean"])))
    print("ENet out PSNR %.2f dB (peak %.3f), max_abs %.5f; low %.2f dB" % (p_out, peak, m, p_low))
    assert p_low >= 45.0 and p_out >= 45.0
    assert torch.equal(net(mel.cuda(), face.cuda(), gt.cuda(), noises=noises)[0], out)     # graph replay is bit-identical


def test_enet_vs_oracle_batch_resized_inputs(env):
    """B = 3 with 384 x 384 face / gt inputs (what inference.py:262-266 passes): exercises both resize paths and the per-sample
    modulation across a batch; per-sample == batched bit for bit."""
    G, oenet, sd, net = env
    g = torch.Generator().manual_seed(5)
    from oracle import synth
    mel, _ = synth.lnet_inputs(3, seed=8)
    face = torch.rand(3, 6, 384, 384, generator=g)
    face[:, :3, 192:] = 0
    gt = torch.rand(3, 3, 384, 384, generator=g)
    noises = [torch.randn(3, 1, s, s, generator=g).cuda() for s in (200, 200, 400, 400)]
    mel, face, gt = mel.cuda(), face.cuda(), gt.cuda()
    out, low = net(mel, face, gt, noises=noises)
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref_out, ref_low = oenet.enet_forward(sdc, mel, face, gt, noises=noises)
    peak = float(ref_out.abs().max())
    p_out, p_low = _psnr(out, ref_out, peak), _psnr(low, ref_low, 1.0)
    m, _ = G.report("ENet out vs oracle B=3 (384 inputs)", out, ref_out)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write("ENet vs oracle B=3, 384x384 inputs: out PSNR %.2f dB (peak %.3f) max_abs %.5f, low PSNR %.2f dB\n" % (p_out, peak, m, p_low))
    assert p_out >= 45.0 and p_low >= 45.0
    one_out, one_low = net(mel[1:2], face[1:2], gt[1:2], noises=[n[1:2] for n in noises])
    assert torch.equal(one_out, out[1:2]) and torch.equal(one_low, low[1:2])
    # default noise path: torch's global RNG, the reference's draw order
    torch.manual_seed(3)
    a, _ = net(mel, face, gt)
    torch.manual_seed(3)
    exp = [torch.empty(3, 1, s, s, device="cuda").normal_() for s in (200, 200, 400, 400)]
    b, _ = net(mel, face, gt, noises=exp)
    assert torch.equal(a, b)
    e_out, e_low = net(mel[:0], face[:0], gt[:0])
    assert e_out.shape == (0, 3, 384, 384) and e_low.shape == (0, 3, 96, 96)


def test_load_network_round_trip(env, tmp_path):
    """models.load_network (models/__init__.py:29-35): LNet checkpoint + ENet checkpoint whose low_res.* keys are skipped."""
    G, oenet, sd, net = env
    from s2v_b200 import models
    lsd = {"module." + k[len("low_res."):]: v for k, v in sd.items() if k.startswith("low_res.")}
    torch.save({"state_dict": lsd}, tmp_path / "lnet.pth")
    esd = {"module." + k: (torch.zeros_like(v) if k.startswith("low_res.") else v) for k, v in sd.items()}   # stale low_res copies must be ignored
    torch.save({"state_dict": esd}, tmp_path / "enet.pth")
    args = types.SimpleNamespace(LNet_path=str(tmp_path / "lnet.pth"), ENet_path=str(tmp_path / "enet.pth"))
    model = models.load_network(args)
    assert not model.training
    got = model.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
