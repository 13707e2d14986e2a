"""Whole-network C entry points (include/s2v.h: s2v_plan_*, s2v_lnet_forward, s2v_dnet_forward): a plan file exported by
the Python package, replayed through the C ABI only, must reproduce the Python engine bit for bit - same launchers, same
arguments, same tile shapes; only the addresses differ.  Also driven from a C program without Python (examples/run_plan)."""
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def G():
    import gpu_util
    from oracle import synth, weights
    gpu_util.lib()
    gpu_util.synth, gpu_util.weights = synth, weights      # seeded inputs / weights only: the checker here is the Python engine
    return gpu_util


def _lnet(G, seed=0):
    from s2v_b200.models.LNet import LNet
    net = LNet().cuda().eval()
    net.load_state_dict(G.weights.make_state_dict("lnet", seed), strict=True)
    return net


def _dnet(G, seed=0):
    from s2v_b200.models.DNet import DNet
    net = DNet().cuda().eval()
    net.load_state_dict(G.weights.make_state_dict("dnet", seed), strict=True)
    return net


def test_lnet_plan_file_replays_bit_exact(G, tmp_path):
    from s2v_b200 import plan_export
    net = _lnet(G)
    mel, face = G.synth.lnet_inputs(8, 0)
    mel, face = mel.cuda(), face.cuda()
    want = net(mel, face)
    path = str(tmp_path / "lnet_b8.s2vplan")
    info = plan_export.export_lnet(net, 8, path)
    assert info["ops"] == net.engine().launches_per_forward(8) and info["io"] == ["mel", "face", "out"]
    p = plan_export.NativePlan(path, "cuda:0")
    assert p.num_ops() == info["ops"]
    got = p.lnet_forward(mel, face)
    torch.cuda.synchronize()
    assert torch.equal(got, want), (got - want).abs().max().item()
    got2 = p.lnet_forward(mel.flip(0), face.flip(0))          # replay on other inputs: still the engine's result
    assert torch.equal(got2, net(mel.flip(0), face.flip(0)))
    G.report("LNet B=8 plan file through s2v_lnet_forward (C ABI, %d launches) vs Python engine" % info["ops"], got, want)
    p.close()


@pytest.mark.parametrize("stage", [None, "warp"])
def test_dnet_plan_file_replays_bit_exact(G, tmp_path, stage):
    from s2v_b200 import plan_export
    net = _dnet(G)
    src, coeff = G.synth.dnet_inputs(8, 0)
    src, coeff = src.cuda(), coeff.cuda()
    want = net(src, coeff, stage=stage)
    path = str(tmp_path / "dnet_b8.s2vplan")
    plan_export.export_dnet(net, 8, path, T=coeff.shape[2], stage=stage)
    p = plan_export.NativePlan(path, "cuda:0")
    got = p.dnet_forward(src, coeff)
    torch.cuda.synchronize()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), (k, (got[k] - want[k]).abs().max().item())
    p.close()


def test_fresh_workspace_contents_do_not_matter(G, tmp_path):
    """Bind the same plan to a workspace full of NaN bit patterns: the plan's own init list must cover every region it
    reads before writing."""
    import ctypes as C
    from s2v_b200 import plan_export, _lib as L
    net = _lnet(G)
    mel, face = G.synth.lnet_inputs(8, 1)
    mel, face = mel.cuda(), face.cuda()
    want = net(mel, face)
    path = str(tmp_path / "lnet_b8.s2vplan")
    plan_export.export_lnet(net, 8, path)
    p = plan_export.NativePlan(path, "cuda:0")
    p.work.fill_(0xFF)
    L.check(p.lib.s2v_plan_bind(p.h, p.const.data_ptr(), p.work.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "rebind")
    got = p.lnet_forward(mel, face)
    assert torch.equal(got, want), (got - want).abs().max().item()
    p.close()


def test_c_host_without_python(G, tmp_path):
    """examples/run_plan.c: plan file + raw float32 inputs in, raw outputs out; equal to the Python engine bit for bit."""
    from conftest import ROOT
    from s2v_b200 import plan_export
    exe = os.path.join(ROOT, "examples", "run_plan")
    if not os.path.exists(exe):
        pytest.skip("examples/run_plan not built (python __graft_entry__.py)")
    net = _lnet(G)
    mel, face = G.synth.lnet_inputs(8, 0)
    want = net(mel.cuda(), face.cuda()).cpu().numpy()
    path = str(tmp_path / "lnet_b8.s2vplan")
    plan_export.export_lnet(net, 8, path)
    mel.numpy().astype(np.float32).tofile(tmp_path / "mel.f32")
    face.numpy().astype(np.float32).tofile(tmp_path / "face.f32")
    r = subprocess.run([exe, path, str(tmp_path / "o"), "mel=%s" % (tmp_path / "mel.f32"), "face=%s" % (tmp_path / "face.f32")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "o.out.f32", dtype=np.float32).reshape(want.shape)
    assert np.array_equal(got, want), np.abs(got - want).max()
