"""-m gpu parity of the CUDA LNet (drop-in module, through the C ABI) against
(1) the committed golden output of the real reference and (2) the oracle restatement.
Gate: PSNR >= 45 dB (peak 1) vs the fp32 reference, max-abs reported."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402


@pytest.fixture(scope="module")
def env():
    import gpu_util
    from oracle import weights
    from s2v_b200.models.LNet import LNet
    gpu_util.lib()
    sd = weights.make_state_dict("lnet", 0)
    net = LNet().cuda().eval()
    missing = net.load_state_dict(sd, strict=True)
    return gpu_util, sd, net


def _check(G, name, got, ref, min_psnr=45.0):
    m, _ = G.report(name, got, ref)
    p = G.psnr(got, ref, 1.0)
    print("%s PSNR %.2f dB, max_abs %.4f" % (name, p, m))
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write("%s PSNR %.2f dB max_abs %.5f\n" % (name, p, m))
    assert p >= min_psnr, "%s: PSNR %.2f dB below %.1f" % (name, p, min_psnr)
    return p


def test_state_dict_schema(env):
    G, sd, net = env
    out = net.state_dict()
    assert list(out.keys()) == list(sd.keys())
    assert all(torch.equal(out[k].cpu(), sd[k]) for k in sd)


def test_lnet_vs_reference_golden(env):
    G, sd, net = env
    from oracle import synth
    mel, face = synth.lnet_inputs(2, seed=0)
    gold = torch.from_numpy(np.load(os.path.join(GOLDEN, "lnet_seed0_b2_out.npy")))
    out = net(mel.cuda(), face.cuda())
    assert out.shape == (2, 3, 96, 96) and out.dtype == torch.float32
    _check(G, "LNet tc vs reference golden B=2", out.cpu(), gold)
    # replays (the third call runs from the captured CUDA graph) are bit-identical
    o2 = net(mel.cuda(), face.cuda())
    o3 = net(mel.cuda(), face.cuda())
    o4 = net(mel.cuda(), face.cuda())
    assert torch.equal(o2, out) and torch.equal(o3, out) and torch.equal(o4, out)


def test_lnet_vs_oracle_ragged_batch_and_5d(env):
    G, sd, net = env
    from oracle import nets, synth
    sdc = {k: v.cuda() for k, v in sd.items()}
    mel, face = synth.lnet_inputs(11, seed=3)          # 11: ragged vs the 8-image box of the 12x12 tiles
    mel, face = mel.cuda(), face.cuda()
    ref = nets.lnet_forward(sdc, mel, face)
    out = net(mel, face)
    _check(G, "LNet tc vs oracle B=11", out, ref)
    # frames are independent: a sub-batch gives the same frames bit-for-bit (shard equivalence)
    sub = net(mel[3:8], face[3:8])
    assert torch.equal(sub, out[3:8])
    # 5-D training form (models/LNet.py:125-127,134-136): B=2, T=2
    a5 = torch.stack([mel[:2], mel[2:4]], 1)
    f5 = torch.stack([face[:2], face[2:4]], 2)
    out5 = net(a5, f5)
    assert out5.shape == (2, 3, 2, 96, 96)
    assert torch.equal(out5[:, :, 0], out[:2]) and torch.equal(out5[:, :, 1], out[2:4])


def test_lnet_empty_and_single_frame(env):
    """Edge cases: an empty batch gives an empty tensor (no launch), a single frame equals that frame of a larger batch."""
    G, sd, net = env
    from oracle import synth
    mel, face = synth.lnet_inputs(3, seed=5)
    mel, face = mel.cuda(), face.cuda()
    e = net(mel[:0], face[:0])
    assert e.shape == (0, 3, 96, 96) and e.dtype == torch.float32
    full = net(mel, face)
    one = net(mel[1:2], face[1:2])
    assert one.shape == (1, 3, 96, 96) and torch.equal(one, full[1:2])


def test_lnet_simt_path_agrees(env):
    """The SIMT convolution path (every conv on CUDA cores, fp32 weights) is an independent
    implementation of the same layers; both must sit on the oracle."""
    G, sd, net = env
    from oracle import nets, synth
    from s2v_b200.models.LNet import LNet
    net2 = LNet(conv_impl="simt", use_graph=False).cuda().eval()
    net2.load_state_dict(sd, strict=True)
    mel, face = synth.lnet_inputs(2, seed=0)
    ref = nets.lnet_forward({k: v.cuda() for k, v in sd.items()}, mel.cuda(), face.cuda())
    out = net2(mel.cuda(), face.cuda())
    _check(G, "LNet simt vs oracle B=2", out, ref)


def test_lnet_errors(env):
    G, sd, net = env
    from s2v_b200 import _lib as L
    from s2v_b200.models.LNet import LNet
    with pytest.raises(L.S2VError):
        LNet().eval()(torch.zeros(1, 1, 80, 16), torch.zeros(1, 6, 96, 96))     # CPU module: no fallback


@pytest.mark.parametrize("seed", [0, 1])
def test_lnet_bench_config_b128_vs_oracle(seed):
    """BASELINE.json configs[1] as benchmarked: batch 128 (full 8-image boxes at 12 x 12, CTA pairs / resident plans keyed on
    the tile count), weight seeds 0 and 1, against the fp32 oracle on the GPU with TF32 off.  Gate: PSNR >= 45 dB, max-abs printed."""
    import gpu_util as G
    from oracle import nets, synth, weights
    from s2v_b200.models.LNet import LNet
    G.lib()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = weights.make_state_dict("lnet", seed)
        net = LNet().cuda().eval()
        net.load_state_dict(sd, strict=True)
        mel, face = synth.lnet_inputs(128, seed=100 + seed)
        mel, face = mel.cuda(), face.cuda()
        out = net(mel, face)
        out2 = net(mel, face)                              # graph replay
        assert torch.equal(out, out2)
        sdc = {k: v.cuda() for k, v in sd.items()}
        ref = torch.cat([nets.lnet_forward(sdc, mel[i:i + 32], face[i:i + 32]) for i in range(0, 128, 32)])
        p = _check(G, "LNet tc vs oracle B=128 weight seed %d" % seed, out, ref)
        worst = min(G.psnr(out[i], ref[i], 1.0) for i in range(128))
        print("worst single frame %.2f dB" % worst)
        assert worst >= 45.0
        # the frames of the big batch equal the same frames computed in a small batch, bit for bit
        assert torch.equal(net(mel[40:48], face[40:48]), out[40:48])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
