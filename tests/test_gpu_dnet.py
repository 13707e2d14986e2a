"""-m gpu parity of the CUDA DNet (drop-in module, through the C ABI) against the committed golden
output of the real reference and the oracle restatement.  Gate: PSNR >= 45 dB (peak 2, images in
[-1,1]) for warp_image / fake_image vs the fp32 reference; flow_field max-abs reported and bounded."""
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
    from s2v_b200.models.DNet import DNet
    gpu_util.lib()
    sd = weights.make_state_dict("dnet", 0)
    net = DNet().cuda().eval()
    net.load_state_dict(sd, strict=True)
    return gpu_util, sd, net


def _check(G, name, got, ref, peak, min_psnr):
    m, _ = G.report(name, got, ref)
    p = G.psnr(got, ref, peak)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write("%s PSNR %.2f dB (peak %g) max_abs %.5f\n" % (name, p, peak, m))
    print("%s PSNR %.2f dB max_abs %.5f" % (name, p, m))
    assert p >= min_psnr, "%s: PSNR %.2f dB below %.1f" % (name, p, min_psnr)


def test_dnet_vs_reference_golden(env):
    G, sd, net = env
    from oracle import synth
    src, coeff = synth.dnet_inputs(1, seed=0)
    g = np.load(os.path.join(GOLDEN, "dnet_seed0_b1_out.npz"))
    out = net(src.cuda(), coeff.cuda())
    assert set(out) == {"flow_field", "warp_image", "fake_image"}
    assert out["flow_field"].shape == (1, 2, 64, 64) and out["fake_image"].shape == (1, 3, 256, 256)
    fl = torch.from_numpy(g["flow_field"])
    _check(G, "DNet flow_field vs golden", out["flow_field"].cpu(), fl, float(fl.abs().max()), 45.0)
    _check(G, "DNet warp_image vs golden", out["warp_image"].cpu(), torch.from_numpy(g["warp_image"].astype(np.float32)), 2.0, 45.0)
    _check(G, "DNet fake_image vs golden", out["fake_image"].cpu(), torch.from_numpy(g["fake_image"].astype(np.float32)), 2.0, 45.0)


def test_dnet_vs_oracle_batch_and_stage(env):
    G, sd, net = env
    from oracle import nets, synth
    sdc = {k: v.cuda() for k, v in sd.items()}
    src, coeff = synth.dnet_inputs(3, seed=5, t=27)             # T=27: any T >= 25 is legal
    src, coeff = src.cuda(), coeff.cuda()
    ref = nets.dnet_forward(sdc, src, coeff)
    out = net(src, coeff)
    fl = ref["flow_field"]
    _check(G, "DNet flow_field vs oracle B=3", out["flow_field"], fl, float(fl.abs().max()), 45.0)
    _check(G, "DNet warp_image vs oracle B=3", out["warp_image"], ref["warp_image"], 2.0, 45.0)
    _check(G, "DNet fake_image vs oracle B=3", out["fake_image"], ref["fake_image"], 2.0, 45.0)
    w = net(src, coeff, stage="warp")
    assert set(w) == {"flow_field", "warp_image"}
    assert torch.equal(w["flow_field"], out["flow_field"]) and torch.equal(w["warp_image"], out["warp_image"])
    # frames are independent: per-sample == batched, bit for bit
    one = net(src[1:2], coeff[1:2])
    assert torch.equal(one["fake_image"], out["fake_image"][1:2])
    # the warp of the module equals the fused warp kernel applied to the module's own flow
    from s2v_b200.futils import flow_util
    again = flow_util.warp_image(src, flow_util.convert_flow_to_deformation(out["flow_field"]))
    assert (again - out["warp_image"]).abs().max().item() < 1e-5


def test_dnet_empty_batch(env):
    G, sd, net = env
    out = net(torch.zeros(0, 3, 256, 256, device="cuda"), torch.zeros(0, 73, 26, device="cuda"))
    assert out["flow_field"].shape == (0, 2, 64, 64) and out["warp_image"].shape == (0, 3, 256, 256) and out["fake_image"].shape == (0, 3, 256, 256)
    out = net(torch.zeros(0, 3, 256, 256, device="cuda"), torch.zeros(0, 73, 26, device="cuda"), stage="warp")
    assert "fake_image" not in out


def test_dnet_errors(env):
    G, sd, net = env
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 256, 256, device="cuda"), torch.zeros(1, 73, 20, device="cuda"))


@pytest.mark.parametrize("seed", [0, 1])
def test_dnet_bench_config_b64_vs_oracle(seed):
    """BASELINE.json configs[2] as benchmarked: batch 64 at 256 x 256, weight seeds 0 and 1, against the fp32 oracle on the
    GPU with TF32 off.  Gate: PSNR >= 45 dB (peak 2) on warp_image / fake_image, flow_field relative to its own peak."""
    import gpu_util as G
    from oracle import nets, synth, weights
    from s2v_b200.models.DNet import DNet
    G.lib()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sd = weights.make_state_dict("dnet", seed)
        net = DNet().cuda().eval()
        net.load_state_dict(sd, strict=True)
        src, coeff = synth.dnet_inputs(64, seed=200 + seed)
        src, coeff = src.cuda(), coeff.cuda()
        out = net(src, coeff)
        sdc = {k: v.cuda() for k, v in sd.items()}
        refs = [nets.dnet_forward(sdc, src[i:i + 16], coeff[i:i + 16]) for i in range(0, 64, 16)]
        ref = {k: torch.cat([r[k] for r in refs]) for k in refs[0]}
        fl = ref["flow_field"]
        tag = "B=64 weight seed %d" % seed
        _check(G, "DNet flow_field vs oracle " + tag, out["flow_field"], fl, float(fl.abs().max()), 45.0)
        _check(G, "DNet warp_image vs oracle " + tag, out["warp_image"], ref["warp_image"], 2.0, 45.0)
        _check(G, "DNet fake_image vs oracle " + tag, out["fake_image"], ref["fake_image"], 2.0, 45.0)
        worst = min(G.psnr(out["fake_image"][i], ref["fake_image"][i], 2.0) for i in range(64))
        print("worst single frame %.2f dB" % worst)
        assert worst >= 45.0
        one = net(src[17:18], coeff[17:18])
        assert torch.equal(one["fake_image"], out["fake_image"][17:18])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
