"""CPU (-m "not gpu") tests of the host logic: the C-ABI library loads and exports every symbol the
header declares, the drop-in modules keep the reference's state_dict schema, plans build with
consistent shapes, box selection, weight packing, window index helper (host C function)."""
import json
import os
import re

import pytest
import torch

import s2v_b200
from s2v_b200 import _lib as L
from s2v_b200 import ops

from conftest import GOLDEN, ROOT


def test_library_exports_every_declared_symbol():
    lib = s2v_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "s2v.h")).read()
    declared = set(re.findall(r"\b(s2v_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libs2v.so does not export %s" % name
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert lib.s2v_version() >= 100
    assert lib.s2v_strerror(-1).decode().startswith("invalid")


def test_window_helpers_match_oracle():
    from oracle import mel as omel
    from s2v_b200.futils import audio
    for T in (16, 17, 31, 401, 4801, 48001):
        for fps in (25.0, 23.976, 30.0, 60.0):
            assert audio.mel_window_starts(T, fps) == omel.mel_window_starts(T, fps)
    assert audio.mel_window_count(401) == 122 and audio.mel_window_count(4801) == 1497 and audio.mel_window_count(48001) == 14997
    with pytest.raises(ValueError):
        audio.mel_window_count(15)


def test_mel_basis_matches_oracle():
    import numpy as np
    from oracle import mel as omel
    from s2v_b200.futils import audio
    assert np.array_equal(audio._build_mel_basis(), omel.mel_basis())


def test_no_cpu_fallback():
    import numpy as np
    from s2v_b200.futils import audio, flow_util
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    with pytest.raises(L.S2VError):
        audio.melspectrogram(np.zeros(1600, np.float32))
    with pytest.raises(L.S2VError):
        flow_util.warp_image(torch.zeros(1, 3, 8, 8), torch.zeros(1, 8, 8, 2))
    from s2v_b200.futils import inference_utils as iu
    with pytest.raises(L.S2VError):
        iu.transform_semantic(np.zeros((30, 262), np.float32), 3)
    with pytest.raises(L.S2VError):
        iu.semantic_windows(torch.zeros(30, 262), (0, 4))
    with pytest.raises(L.S2VError):
        iu.Laplacian_Pyramid_Blending_with_mask(np.zeros((64, 64, 3), np.uint8), np.zeros((64, 64, 3), np.uint8), np.zeros((64, 64), np.float32))
    with pytest.raises(L.S2VError):
        iu.laplacian_blend(torch.zeros(1, 64, 64, 3, dtype=torch.uint8), torch.zeros(1, 64, 64, 3, dtype=torch.uint8), torch.zeros(1, 64, 64))


def test_choose_box():
    assert ops.choose_box(96, 96, 128) == (32, 4, 1)
    assert ops.choose_box(48, 48, 128) == (16, 8, 1)
    assert ops.choose_box(24, 24, 128) == (8, 8, 2)
    assert ops.choose_box(12, 12, 128) == (4, 4, 8)
    assert ops.choose_box(1, 10752, 1) == (128, 1, 1)
    for h, w, n in ((256, 256, 64), (64, 64, 3), (8, 8, 5), (7, 12, 9)):
        bw, bh, bn = ops.choose_box(h, w, n)
        assert bw * bh * bn == 128


def test_weight_packing_layouts():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_w_tc(w)
    assert p.shape == (8, 9 * 64) and p.dtype == torch.float16 and (p[2:] == 0).all()
    assert p[1, 4 * 64 + 2].item() == w[1, 2, 1, 1].item() and p[1, 4 * 64 + 3].item() == 0     # one chunk: [tap][64]
    w2 = torch.randn(8, 128, 3, 3)
    p2 = ops.pack_w_tc(w2)                          # chunk-major: [cout][chunk][tap][64]
    assert p2.shape == (8, 2 * 9 * 64) and p2[3, (1 * 9 + 5) * 64 + 7].item() == w2[3, 64 + 7, 1, 2].half().item()
    s = ops.pack_w_simt(w, 8)
    assert s.shape == (9, 8, 4) and s[5, 1, 1].item() == w[1, 1, 1, 2].item()


@pytest.mark.parametrize("net", ["lnet", "dnet"])
def test_schema_enumeration_matches_reference(net):
    from s2v_b200.models import _schema
    spec = _schema.lnet_spec() if net == "lnet" else _schema.dnet_spec()
    ref = json.load(open(os.path.join(GOLDEN, f"{net}_schema.json")))
    assert [n for n, _, _ in spec] == list(ref.keys())
    assert all(list(s) == ref[n] for n, s, _ in spec)


def test_lnet_module_state_dict_and_plan():
    from oracle import weights
    from s2v_b200.models.LNet import LNet, LNetEngine
    sd = weights.make_state_dict("lnet", 0)
    net = LNet().eval()
    net.load_state_dict(sd, strict=True)
    assert list(net.state_dict().keys()) == list(sd.keys())
    for impl in ("tc", "simt"):
        eng = LNetEngine(sd, torch.device("cpu"), conv_impl=impl)       # plan build only: validates shapes/ABI structs
        ent = eng._get_plan(3, eng._build(3))
        assert len(ent["plan"]) > (400 if impl == "tc" else 500)      # tc: statistics fused into the convs
        assert ent["io"]["out"].shape == (3, 3, 96, 96)


def test_dnet_module_state_dict_and_plan():
    from oracle import weights
    from s2v_b200.models.DNet import DNet, DNetEngine
    sd = weights.make_state_dict("dnet", 0)
    net = DNet().eval()
    net.load_state_dict(sd, strict=True)
    assert list(net.state_dict().keys()) == list(sd.keys())
    eng = DNetEngine(sd, torch.device("cpu"))                          # plan build only
    full = eng._get_plan((2, 26, "full"), eng._build(2, 26, "full"))
    warp = eng._get_plan((2, 26, "warp"), eng._build(2, 26, "warp"))
    assert "fake" in full["io"] and "fake" not in warp["io"]
    gf = sum(getattr(op, "alg_flops", 0.0) for op in full["plan"].ops) / 2 / 1e9
    assert abs(gf - 101.45) < 0.1                                       # useful GFLOP/frame (dead conv1 skipped)


def test_convT_and_up2_phase_weights_match_torch():
    import torch.nn.functional as F
    from s2v_b200.models.DNet import convT_phase_weights
    from s2v_b200.models._engine import up2_phase_weights
    torch.manual_seed(0)
    x = torch.randn(1, 5, 6, 7, dtype=torch.float64)
    wt = torch.randn(5, 4, 3, 3, dtype=torch.float64)
    ref = F.conv_transpose2d(x, wt, stride=2, padding=1, output_padding=1)
    out = torch.zeros_like(ref)
    for (p, q), w4 in convT_phase_weights(wt).items():
        kh, kw = w4.shape[2:]
        out[:, :, p::2, q::2] = F.conv2d(F.pad(x, (0, kw - 1, 0, kh - 1)), w4)
    assert (out - ref).abs().max() < 1e-12
    w3 = torch.randn(4, 5, 3, 3, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2), w3, padding=1)
    out = torch.zeros_like(ref)
    for (p, q), w4 in up2_phase_weights(w3).items():
        out[:, :, p::2, q::2] = F.conv2d(F.pad(x, (1 - q, q, 1 - p, p)), w4)
    assert (out - ref).abs().max() < 1e-12


def test_enet_module_schema_and_loaders(tmp_path):
    """models.ENet keeps the reference's 1 785-tensor schema (low_res.* first); load_checkpoint strips `module.`, skips `low_res`
    keys and falls back to a plain state_dict file (models/__init__.py:12-27); load_network wires LNet into ENet (:29-35)."""
    import types
    from s2v_b200 import models
    from s2v_b200.models.ENet import ENet
    from s2v_b200.models.LNet import LNet
    with open(os.path.join(GOLDEN, "enet_schema.json")) as f:
        own = json.load(f)
    with open(os.path.join(GOLDEN, "lnet_schema.json")) as f:
        lnet_schema = json.load(f)
    net = ENet(lnet=LNet())
    sd = net.state_dict()
    assert list(sd)[:len(lnet_schema)] == ["low_res." + k for k in lnet_schema]
    assert list(sd)[len(lnet_schema):] == list(own) and all(list(sd[k].shape) == own[k] for k in own)
    with pytest.raises(L.S2VError):
        ENet(lnet=LNet(), concat=True)
    with pytest.raises(TypeError):
        ENet()
    # checkpoints: wrapped ({"state_dict": {"module.x": ...}}) and plain
    lsd = {k: torch.randn_like(v) if v.is_floating_point() else v for k, v in LNet().state_dict().items()}
    torch.save({"state_dict": {"module." + k: v for k, v in lsd.items()}}, tmp_path / "l.pth")
    esd = {k: torch.randn_like(v) for k, v in sd.items() if not k.startswith("low_res.")}
    stale = {"module.low_res." + k: torch.zeros_like(v) for k, v in lsd.items()}
    torch.save({"state_dict": {**stale, **{"module." + k: v for k, v in esd.items()}}}, tmp_path / "e.pth")
    model = models.load_network(types.SimpleNamespace(LNet_path=str(tmp_path / "l.pth"), ENet_path=str(tmp_path / "e.pth")))
    got = model.state_dict()
    assert not model.training
    assert all(torch.equal(got["low_res." + k], lsd[k]) for k in lsd) and all(torch.equal(got[k], esd[k]) for k in esd)
    torch.save(lsd, tmp_path / "plain.pth")                       # the reference's fallback branch: a bare state_dict
    l2 = models.load_checkpoint(str(tmp_path / "plain.pth"), LNet())
    assert all(torch.equal(l2.state_dict()[k], lsd[k]) for k in lsd)
    with pytest.raises(L.S2VError):                               # CPU module: no fallback
        model(torch.zeros(1, 1, 80, 16), torch.zeros(1, 6, 96, 96), torch.zeros(1, 3, 96, 96))


def test_pipeline_batches_and_plan_cache_bound():
    from s2v_b200.pipeline import balanced_batches
    assert balanced_batches(1497, 256) == [256] * 5 + [217] and balanced_batches(188, 192) == [188]
    assert balanced_batches(0, 64) == [] and balanced_batches(64, 64) == [64] and sum(balanced_batches(14997, 192)) == 14997
    # the per-engine plan cache is an LRU bounded by count (and bytes): plans built on a CPU "device" only validate shapes
    from oracle import weights
    from s2v_b200.models.LNet import LNetEngine
    eng = LNetEngine(weights.make_state_dict("lnet", 0), torch.device("cpu"))
    eng.max_plans = 2
    for b in (8, 16, 24):
        eng.plan_for(b)
    info = eng.plan_cache_info()
    assert info["plans"] == 2 and info["keys"] == [16, 24] and info["bytes"] > 0
    eng.plan_for(16)
    assert eng.plan_cache_info()["keys"] == [24, 16]              # a hit moves the plan to the young end


def _tiny_plan_bytes():
    """A hand-written one-launch plan (s2v_add on three 8-channel views of the workspace): exercises the host-side parser
    of csrc/plan.cu without a GPU."""
    import ctypes as C
    import struct
    view = L.View(None, 1, 2, 2, 8, 32, 16, 8)
    blob = bytes(view)
    def arg_view(off):
        return struct.pack("<II", 3, len(blob)) + blob + struct.pack("<II", 1, 0) + struct.pack("<IIQ", 0, 1, off)
    op = struct.pack("<40sII", b"s2v_add", 3, 0) + arg_view(0) + arg_view(256) + arg_view(512)
    io = struct.pack("<32sQQII", b"a", 0, 64, 0, 0) + struct.pack("<32sQQII", b"y", 512, 64, 1, 0)
    init = struct.pack("<QQIf", 256, 64, 0, 0.0)
    image = bytes(range(256))
    head = b"S2VPLAN1" + struct.pack("<IIQQIIII", 1, 1, len(image), 768, 2, 1, 0, 0)
    return head + io + init + op + image


def test_plan_file_parser_host_side():
    """s2v_plan_load_memory / s2v_plan_*: header, I/O table and op table of a plan file are parsed and bounds-checked on
    the host; malformed files are rejected with S2V_EINVAL (never a crash)."""
    import ctypes as C
    lib = s2v_b200.load_library()
    data = _tiny_plan_bytes()
    h = C.c_void_p()
    assert lib.s2v_plan_load_memory(data, len(data), C.byref(h)) == 0
    assert lib.s2v_plan_num_ops(h) == 1 and lib.s2v_plan_num_io(h) == 2
    assert lib.s2v_plan_const_bytes(h) == 256 and lib.s2v_plan_workspace_bytes(h) == 768
    name, off, nb, out = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int()
    assert lib.s2v_plan_io_info(h, 1, C.byref(name), C.byref(off), C.byref(nb), C.byref(out)) == 0
    assert (name.value, off.value, nb.value, out.value) == (b"y", 512, 64, 1)
    assert lib.s2v_plan_io_info(h, 2, None, None, None, None) == -1
    assert lib.s2v_plan_run(h, None) == -1                       # not bound to device buffers yet
    lib.s2v_plan_free(h)
    bad = [b"", b"S2VPLAN0" + data[8:], data[:-1], data + b"\0", data[:100],
           data.replace(b"s2v_add", b"s2v_xyz"),                 # unknown launcher
           data[:16] + (1 << 40).to_bytes(8, "little") + data[24:]]          # constant image larger than the file
    for b in bad:
        h = C.c_void_p()
        assert lib.s2v_plan_load_memory(b, len(b), C.byref(h)) == -1
        assert not h.value
    # a relocation that points outside its arena
    i = data.index(struct_reloc := (512).to_bytes(8, "little"), 200)
    broken = data[:i] + (1 << 30).to_bytes(8, "little") + data[i + 8:]
    h = C.c_void_p()
    assert lib.s2v_plan_load_memory(broken, len(broken), C.byref(h)) == -1
    assert lib.s2v_plan_load(b"/nonexistent/plan.s2vplan", C.byref(h)) == -1
