"""-m gpu parity tests of the individual kernels, called through the C ABI, against the
oracle (oracle/) or a plain PyTorch fp32 reference of the same op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from conftest import GOLDEN  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def G():
    import gpu_util
    gpu_util.lib()
    return gpu_util


def test_flow_warp_vs_oracle_and_golden(G):
    import os
    from oracle import nets, synth
    lib = G.lib()
    # golden produced by the real reference (small)
    s, fl = synth.warp_inputs(2, seed=0, c=3, hw=64, fhw=16)
    gold = torch.from_numpy(np.load(os.path.join(GOLDEN, "warp_seed0_b2_64_16.npy")))
    s, fl = s.cuda(), fl.cuda()
    out = torch.empty_like(s)
    G.ops.op_flow_warp(lib, s, fl, out).run()
    m, _ = G.report("warp vs reference golden 64/16", out.cpu(), gold)
    assert m < 1e-4          # fp32; coordinate rounding (ulp(64)*|dsrc|) bounds this
    # BASELINE config 3(iii) shape, vs the torch restatement of flow_util on the GPU
    s, fl = synth.warp_inputs(4, seed=1)
    s, fl = s.cuda(), fl.cuda()
    out = torch.empty_like(s)
    G.ops.op_flow_warp(lib, s, fl, out).run()
    ref = nets.warp_image(s, nets.convert_flow_to_deformation(fl))
    m, _ = G.report("warp vs oracle 256/64", out, ref)
    assert m < 2e-4 and (out - ref).abs().mean().item() < 2e-6
    # same-size grid (no resize branch) + fp16 NHWC side output
    fl2 = torch.randn(2, 2, 32, 32, device="cuda") * 2
    s2 = torch.rand(2, 3, 32, 32, device="cuda")
    out2 = torch.empty_like(s2)
    side = torch.zeros(2, 32, 32, 8, dtype=torch.float16, device="cuda")
    G.ops.op_flow_warp(lib, s2, fl2, out2, side, 3).run()
    ref2 = nets.warp_image(s2, nets.convert_flow_to_deformation(fl2))
    assert G.report("warp same-size", out2, ref2)[0] < 1e-4
    assert (side[..., 3:6].permute(0, 3, 1, 2).float() - out2).abs().max().item() < 1e-3
    assert side[..., :3].abs().max().item() == 0


@pytest.mark.parametrize("hw,fhw", [(256, 64), (32, 32)])
def test_flow_warp_texel_pack_equals_pack_then_warp(G, hw, fhw):
    """S2V_WARP_PACK_SRC (DNet's stem input torch.cat([input_image, warp_image], 1), models/DNet.py:104): the warp launch that also
    writes the source into the texel gives bit for bit what the separate pack launch + the plain warp launch give - in the 4-pixel
    kernel (256 / 64, DNet's geometry) and in the fallback (same-size grid)."""
    lib = G.lib()
    torch.manual_seed(11)
    s = (torch.rand(3, 3, hw, hw, device="cuda") * 2 - 1).contiguous()
    fl = torch.randn(3, 2, fhw, fhw, device="cuda") * 3
    pad = torch.full((3, hw + 6, hw + 8, 8), 7.0, dtype=torch.float16, device="cuda")      # a padded stem buffer: strided interior view
    a, b = pad.clone(), pad.clone()
    ia, ib = a[:, 3:3 + hw, 3:3 + hw, :], b[:, 3:3 + hw, 3:3 + hw, :]
    oa, ob = torch.empty_like(s), torch.empty_like(s)
    G.ops.op_pack(lib, s, ia, 0, 8).run()
    G.ops.op_flow_warp(lib, s, fl, oa, ia, 3).run()
    G.ops.op_flow_warp(lib, s, fl, ob, ib, 3, pack_src=True).run()
    torch.cuda.synchronize()
    assert torch.equal(oa, ob)
    assert torch.equal(a, b)
    assert torch.equal(ib[..., :3].permute(0, 3, 1, 2), s.half()) and ib[..., 6:].abs().max().item() == 0
    assert torch.equal(b[:, :3], pad[:, :3]) and torch.equal(b[:, :, :3], pad[:, :, :3])       # the border is untouched


def test_mel_vs_oracle(G):
    from oracle import mel as omel, synth
    from s2v_b200.futils import audio
    for seconds, seed in ((1.0, 0), (5.0, 0), (0.3, 3)):
        wav = synth.wav(seconds, seed=seed)
        ref = omel.melspectrogram(wav)
        got = audio.melspectrogram(wav)
        assert got.dtype == np.float64 and got.shape == ref.shape
        d = np.abs(got - ref)
        tol = 1e-4 * np.maximum(np.abs(ref), 1.0)
        print("mel %.1fs max_abs=%.3e" % (seconds, d.max()))
        assert (d <= tol).all()
    # reflect padding variant and the windows
    wav = synth.wav(5.0, seed=0)
    ref = omel.melspectrogram(wav, pad_mode="reflect")
    got = audio.melspectrogram(wav, pad_mode="reflect")
    assert (np.abs(got - ref) <= 1e-4 * np.maximum(np.abs(ref), 1.0)).all()
    mel_dev = audio.melspectrogram_device(torch.from_numpy(wav).cuda())
    win = audio.mel_windows(mel_dev, fps=25.0)
    assert win.shape == (122, 1, 80, 16)
    starts = omel.mel_window_starts(mel_dev.shape[1], 25.0)
    assert audio.mel_window_starts(mel_dev.shape[1], 25.0) == starts          # bit-exact indices
    m = mel_dev.cpu()
    for i in (0, 1, 5, 60, 120, 121):
        assert torch.equal(win[i, 0].cpu(), m[:, starts[i]:starts[i] + 16])   # pure gather: bit-exact
    for fps in (23.976, 30.0, 60.0):
        for T in (16, 17, 401, 4801):
            assert audio.mel_window_starts(T, fps) == omel.mel_window_starts(T, fps)


def test_pack_unpack(G):
    lib = G.lib()
    src = torch.randn(3, 3, 20, 12, device="cuda")
    dst = torch.full((3, 20, 12, 16), 7.0, dtype=torch.float16, device="cuda")
    G.ops.op_pack(lib, src, dst, c_off=8, c_fill=8, scale=2.0, shift=-1.0).run()
    wide = torch.randn(3, 6, 20, 12, device="cuda")
    d2 = torch.zeros(3, 20, 12, 8, dtype=torch.float16, device="cuda")
    G.ops.op_pack(lib, wide[:, 3:6], d2, 0, 8).run()
    assert torch.equal(d2[..., :3], wide[:, 3:6].permute(0, 2, 3, 1).half()) and (d2[..., 3:] == 0).all()
    assert (dst[..., :8] == 7).all() and (dst[..., 11:] == 0).all()
    assert torch.equal(dst[..., 8:11].permute(0, 3, 1, 2), (src * 2 - 1).half())
    back = torch.empty(3, 3, 20, 12, device="cuda")
    G.ops.op_unpack(lib, dst, 8, 3, back).run()
    assert torch.equal(back, dst[..., 8:11].permute(0, 3, 1, 2).float())


@pytest.mark.parametrize("case", [
    dict(n=2, cin=3, cout=64, k=7, pad=3, h=20, w=24),
    dict(n=2, cin=64, cout=3, k=7, pad=3, h=12, w=12, f32=True, act="sigmoid"),
    dict(n=3, cin=32, cout=64, k=3, pad=1, h=20, w=16, stride=(3, 1), act="relu", scale=True),
    dict(n=2, cin=64, cout=64, k=3, pad=1, h=9, w=6, act="relu", res1=True, scale=True),
    dict(n=2, cin=16, cout=32, k=3, pad=1, h=10, w=10, reflect=True),
    dict(n=2, cin=16, cout=32, k=3, pad=1, h=6, w=6, up2=True, act="lrelu"),
    dict(n=2, cin=32, cout=64, k=4, pad=1, h=16, w=16, stride=(2, 2)),
    dict(n=4, cin=512, cout=128, k=1, pad=0, h=1, w=1, act="relu"),
    dict(n=2, cin=256, cout=2, k=7, pad=3, h=8, w=8, f32=True),
    dict(n=2, cin=80, cout=48, k=(1, 3), pad=(0, 0), h=1, w=20, dil=(1, 3), res2=True),
])
def test_conv_simt(G, case):
    lib = G.lib()
    L = G.L
    torch.manual_seed(1)
    n, cin, cout, h, w = case["n"], case["cin"], case["cout"], case["h"], case["w"]
    k = case["k"] if isinstance(case["k"], tuple) else (case["k"], case["k"])
    pad = case["pad"] if isinstance(case["pad"], tuple) else (case["pad"], case["pad"])
    stride, dil = case.get("stride", (1, 1)), case.get("dil", (1, 1))
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, *k, device="cuda") / (cin * k[0] * k[1]) ** 0.5
    bias = torch.randn(cout, device="cuda")
    scale = torch.rand(cout, device="cuda") + 0.5 if case.get("scale") else None
    cin_p = -(-cin // 8) * 8
    xh = torch.zeros(n, h, w, cin_p, dtype=torch.float16, device="cuda")
    xh[..., :cin] = G.nhwc(x)
    xr = xh[..., :cin].permute(0, 3, 1, 2).float()
    xin = F.interpolate(xr, scale_factor=2) if case.get("up2") else xr
    if case.get("reflect"):
        ref = F.conv2d(F.pad(xin, (pad[1], pad[1], pad[0], pad[0]), mode="reflect"), wt, None, stride, 0, dil)
    else:
        ref = F.conv2d(xin, wt, None, stride, pad, dil)
    if scale is not None:
        ref = ref * scale[None, :, None, None]
    ref = ref + bias[None, :, None, None]
    oh, ow = ref.shape[2:]
    r1 = r2 = None
    if case.get("res1"):
        r1 = torch.randn(n, oh, ow, cout, device="cuda").half()
        ref = ref + r1.permute(0, 3, 1, 2).float()
    act = {"relu": L.ACT_RELU, "lrelu": L.ACT_LRELU, "sigmoid": L.ACT_SIGMOID, None: L.ACT_NONE}[case.get("act")]
    ref = {"relu": F.relu, "lrelu": lambda t: F.leaky_relu(t, 0.1), "sigmoid": torch.sigmoid, None: lambda t: t}[case.get("act")](ref)
    if case.get("res2"):
        r2 = torch.randn(n, oh, ow, cout, device="cuda").half()
        ref = ref + r2.permute(0, 3, 1, 2).float()
    wp = G.ops.pack_w_simt(wt, cin_p)
    kw = dict(k=k, stride=stride, pad=pad, dil=dil, pad_mode=L.PAD_REFLECT if case.get("reflect") else L.PAD_ZERO,
              up2=1 if case.get("up2") else 0, scale=scale, bias=bias, res1=r1, res2=r2, act=act, act_param=0.1, impl="simt")
    if case.get("f32"):
        y = torch.zeros(n, cout, oh, ow, device="cuda")
        G.ops.op_conv(lib, xh, wp, None, y_f32=y, out_shape=(n, cout, oh, ow), **kw).run()
        got = y
        tol = 2e-4
    else:
        y = torch.zeros(n, oh, ow, -(-cout // 8) * 8, dtype=torch.float16, device="cuda")
        yv = y[..., :cout] if cout % 8 == 0 else y
        G.ops.op_conv(lib, xh, wp, yv, **kw).run()
        got = G.nchw(y[..., :cout])
        tol = 4e-3
    torch.cuda.synchronize()
    m, rel = G.report("conv_simt %s" % case, got, ref)
    assert rel < tol


def _norm_inputs(n, c, h, w):
    torch.manual_seed(2)
    x = torch.randn(n, c, h, w, device="cuda") * 1.7 + 0.3
    return x, G_nhwc(x)


def G_nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().half()


@pytest.mark.parametrize("n,c,h,w,pool", [(3, 64, 24, 24, 0), (2, 128, 16, 16, 1), (2, 512, 12, 12, 0), (1, 1024, 12, 12, 0)])
def test_layernorm2d(G, n, c, h, w, pool):
    lib = G.lib()
    L = G.L
    x, xh = _norm_inputs(n, c, h, w)
    xf = xh.permute(0, 3, 1, 2).float()
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
    ref = F.leaky_relu(F.layer_norm(xf, xf.shape[1:], gamma[:, None, None].expand(c, h, w), beta[:, None, None].expand(c, h, w)), 0.1)
    if pool:
        ref = F.avg_pool2d(ref, 2)
    res = torch.randn(n, ref.shape[2], ref.shape[3], c, device="cuda").half()
    ref = ref + res.permute(0, 3, 1, 2).float()
    chunks = G.ops.stats_chunks(n, h * w, c)
    partial = torch.empty(n, chunks, c, 2, device="cuda")
    a, b = torch.empty(n, c, device="cuda"), torch.empty(n, c, device="cuda")
    y = torch.empty(n, ref.shape[2], ref.shape[3], c, dtype=torch.float16, device="cuda")
    G.ops.op_chan_stats(lib, xh, chunks, partial).run()
    G.ops.op_ln2d_finalize(lib, partial, n, chunks, c, h * w, gamma, beta, a, b).run()
    G.ops.op_affine_act(lib, xh, a, b, y, act=L.ACT_LRELU, act_param=0.1, pool2=pool, res=res).run()
    m, rel = G.report("layernorm2d n%d c%d %dx%d pool%d" % (n, c, h, w, pool), G.nchw(y), ref)
    assert rel < 2e-3


@pytest.mark.parametrize("n,c,h,w", [(3, 64, 12, 12), (2, 1024, 12, 12), (2, 256, 24, 24), (2, 32, 40, 40)])
def test_adain_reflect(G, n, c, h, w):
    lib = G.lib()
    L = G.L
    x, xh = _norm_inputs(n, c, h, w)
    xf = xh.permute(0, 3, 1, 2).float()
    gb = torch.randn(n, 2 * c + 5, device="cuda") * 0.3
    gamma, beta = gb[:, :c], gb[:, c:2 * c]
    ref = F.leaky_relu(F.instance_norm(xf, eps=1e-5) * (1 + gamma[:, :, None, None]) + beta[:, :, None, None], 0.01)
    res = torch.randn(n, h, w, c, device="cuda").half()
    ref = ref + res.permute(0, 3, 1, 2).float()
    refp = F.pad(ref, (1, 1, 1, 1), mode="reflect")
    chunks = G.ops.stats_chunks(n, h * w, c)
    partial = torch.empty(n, chunks, c, 2, device="cuda")
    a, b = torch.empty(n, c, device="cuda"), torch.empty(n, c, device="cuda")
    yp = torch.zeros(n, h + 2, w + 2, c, dtype=torch.float16, device="cuda")
    G.ops.op_chan_stats(lib, xh, chunks, partial).run()
    G.ops.op_adain_finalize(lib, partial, n, chunks, c, h * w, gamma, beta, gb.stride(0), a, b).run()
    G.ops.op_affine_act(lib, xh, a, b, yp[:, 1:-1, 1:-1, :], act=L.ACT_LRELU, act_param=0.01, res=res, reflect1=1).run()
    m, rel = G.report("adain+reflect n%d c%d %dx%d" % (n, c, h, w), G.nchw(yp), refp)
    assert rel < 2e-3
    # standalone border fill
    yp2 = yp.clone()
    yp2[:, 0] = 0; yp2[:, -1] = 0; yp2[:, :, 0] = 0; yp2[:, :, -1] = 0
    G.ops.op_reflect_border(lib, yp2[:, 1:-1, 1:-1, :]).run()
    assert torch.equal(yp2, yp)


@pytest.mark.parametrize("n,c,h,w,res", [(3, 1024, 12, 12, True), (2, 256, 24, 24, False), (2, 128, 48, 48, True), (2, 64, 20, 20, True)])
def test_adain_fused_single_pass(G, n, c, h, w, res):
    lib, L = G.lib(), G.L
    x, xh = _norm_inputs(n, c, h, w)
    xf = xh.permute(0, 3, 1, 2).float()
    gb = torch.randn(n, 2 * c + 3, device="cuda") * 0.3
    gamma, beta = gb[:, :c], gb[:, c:2 * c]
    ref = F.leaky_relu(F.instance_norm(xf, eps=1e-5) * (1 + gamma[:, :, None, None]) + beta[:, :, None, None], 0.01)
    r = torch.randn(n, h, w, c, device="cuda").half() if res else None
    if res:
        ref = ref + r.permute(0, 3, 1, 2).float()
    refp = F.pad(ref, (1, 1, 1, 1), mode="reflect")
    assert lib.s2v_adain_fused_fits(h, w, c) > 0
    yp = torch.zeros(n, h + 2, w + 2, c, dtype=torch.float16, device="cuda")
    G.ops.op_adain_fused(lib, xh, gamma, beta, gb.stride(0), yp[:, 1:-1, 1:-1, :], act=L.ACT_LRELU, act_param=0.01, res=r, reflect1=1).run()
    m, rel = G.report("adain fused n%d c%d %dx%d" % (n, c, h, w), G.nchw(yp), refp)
    assert rel < 2e-3
    assert lib.s2v_adain_fused_fits(256, 256, 32) == 0


def test_token_layernorm_add_mean(G):
    lib = G.lib()
    torch.manual_seed(3)
    x = torch.randn(2, 12, 12, 1024, device="cuda").half()
    xs = x[..., :512]
    g, b = torch.rand(512, device="cuda") + 0.5, torch.randn(512, device="cuda") * 0.1
    y = torch.empty(2, 12, 12, 512, dtype=torch.float16, device="cuda")
    G.ops.op_token_ln(lib, xs, g, b, y).run()
    ref = F.layer_norm(xs.float(), (512,), g, b)
    assert G.report("token_layernorm", y, ref)[1] < 2e-3
    a2 = torch.randn(2, 6, 6, 64, device="cuda").half()
    b2 = torch.randn(2, 6, 6, 64, device="cuda").half()
    y2 = torch.empty_like(a2)
    G.ops.op_add(lib, a2, b2, y2).run()
    assert torch.equal(y2, (a2.float() + b2.float()).half())
    xm = torch.randn(3, 1, 2, 256, device="cuda").half()
    ym = torch.empty(3, 1, 1, 256, dtype=torch.float16, device="cuda")
    G.ops.op_mean_over_w(lib, xm, ym).run()
    assert (ym.float() - xm.float().mean(2, keepdim=True)).abs().max().item() < 2e-3


@pytest.mark.parametrize("s,c", [(12, 384), (24, 96), (48, 48), (12, 8), (24, 24)])
def test_fft2(G, s, c):
    lib = G.lib()
    torch.manual_seed(4)
    n = 3
    x = torch.randn(n, c, s, s, device="cuda")
    xh = G.nhwc(x)
    xf = xh.permute(0, 3, 1, 2).float()
    spec = torch.empty(n, s, s // 2 + 1, 2 * c, dtype=torch.float16, device="cuda")
    G.ops.op_rfft2(lib, xh, spec).run()
    ff = torch.fft.rfftn(xf, dim=(-2, -1), norm="ortho")
    ref = torch.stack((ff.real, ff.imag), dim=-1).permute(0, 1, 4, 2, 3).reshape(n, 2 * c, s, s // 2 + 1)
    assert G.report("rfft2 %dx%d c%d" % (s, s, c), G.nchw(spec), ref)[1] < 2e-3
    # inverse on a NON-Hermitian spectrum (post-ReLU in the reference), + residual add
    z = F.relu(torch.randn(n, 2 * c, s, s // 2 + 1, device="cuda"))
    zh = G.nhwc(z)
    zf = zh.permute(0, 3, 1, 2).float().reshape(n, c, 2, s, s // 2 + 1).permute(0, 1, 3, 4, 2)
    refi = torch.fft.irfftn(torch.complex(zf[..., 0].contiguous(), zf[..., 1].contiguous()), s=(s, s), dim=(-2, -1), norm="ortho")
    add = torch.randn(n, s, s, c, device="cuda").half()
    y = torch.empty(n, s, s, c, dtype=torch.float16, device="cuda")
    G.ops.op_irfft2(lib, zh, add, y).run()
    assert G.report("irfft2 %dx%d c%d" % (s, s, c), G.nchw(y), refi + add.permute(0, 3, 1, 2).float())[1] < 2e-3


def test_fft2_mma_path():
    """The tensor-core DFT variant of the same entry points (csrc/fft2d_mma.cu, S2V_FFT_MMA=63: both directions, all three sizes) against torch.fft.
    The switch is read once per process, so the parametrised test above is re-run in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, S2V_FFT_MMA="63")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", "test_fft2 and not mma_path"],
                       env=env, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "5 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_attention(G):
    lib = G.lib()
    torch.manual_seed(5)
    n, t, heads, dh = 3, 144, 4, 64
    qk = torch.randn(n, 1, t, 2 * heads * dh, device="cuda").half()
    v = torch.randn(n, 1, t, heads * dh, device="cuda").half()
    o = torch.empty(n, 1, t, heads * dh, dtype=torch.float16, device="cuda")
    G.ops.op_attention(lib, qk[..., :256], qk[..., 256:], v, o, heads, dh ** -0.5).run()
    q, k = (qk[..., i * 256:(i + 1) * 256].float().reshape(n, t, heads, dh).permute(0, 2, 1, 3) for i in (0, 1))
    vv = v.float().reshape(n, t, heads, dh).permute(0, 2, 1, 3)
    ref = (torch.matmul(q, k.transpose(-1, -2)) * dh ** -0.5).softmax(-1) @ vv
    ref = ref.permute(0, 2, 1, 3).reshape(n, 1, t, heads * dh)
    assert G.report("attention", o, ref)[1] < 3e-3


def test_grouped_linear(G):
    import ctypes as C
    lib = G.lib()
    L = G.L
    torch.manual_seed(6)
    B = 11
    hidden = F.relu(torch.randn(B, 1, 1, 3 * 128, device="cuda")).half()
    specs = [(0, 256), (128, 768), (256, 40)]
    wts = [torch.randn(128, no, device="cuda") * 0.1 for _, no in specs]
    bs = [torch.randn(no, device="cuda") for _, no in specs]
    total = sum(no for _, no in specs)
    groups = (L.LinGroup * len(specs))()
    tiles, off = [], 0
    for i, (ino, no) in enumerate(specs):
        groups[i] = L.LinGroup(wts[i].data_ptr(), bs[i].data_ptr(), ino, 128, off, no)
        tiles += [(i, j) for j in range(0, no, 128)]
        off += no
    gdev = torch.frombuffer(bytearray(bytes(groups)), dtype=torch.uint8).cuda()
    tdev = torch.tensor(tiles, dtype=torch.int32, device="cuda")
    out = torch.zeros(B, total, device="cuda")
    G.ops.op_grouped_linear(lib, hidden, gdev, tdev, len(tiles), out).run()
    ref = torch.cat([hidden.reshape(B, -1)[:, ino:ino + 128].float() @ wts[i] + bs[i] for i, (ino, _) in enumerate(specs)], 1)
    assert G.report("grouped_linear", out, ref)[1] < 1e-5


def test_semantic_windows_bit_exact(G):
    """s2v_semantic_windows (futils/inference_utils.py:73-91 for a whole batch) against the reference's own outputs
    (tests/golden/semantic_golden.npz) and the oracle at clip size: bit-exact, float32 and float64 tables."""
    import os
    from oracle import semantic as osem
    from s2v_b200.futils import inference_utils as iu
    gold = np.load(os.path.join(GOLDEN, "semantic_golden.npz"))
    for tag, dtype, n in (("f32", np.float32, 40), ("f64", np.float64, 17)):
        table = osem.synth_table(n, seed=3, dtype=dtype)
        ratio = iu.find_crop_norm_ratio(table[0:1], table[1:])
        assert np.array_equal(ratio, gold[tag + "_ratio"])
        dev = iu.upload_semantic(table, "cuda")
        frames = gold[tag + "_frames"]
        assert np.array_equal(iu.semantic_windows(dev, frames).cpu().numpy(), gold[tag + "_plain"])
        assert np.array_equal(iu.semantic_windows(dev, frames, ratio).cpu().numpy(), gold[tag + "_scaled"])
        assert np.array_equal(iu.semantic_windows(dev, (3, 1), np.zeros(1, dtype))[0].cpu().numpy(), gold[tag + "_zero_ratio"])
        one = iu.transform_semantic(table, int(frames[4]), ratio)                 # the reference's per-frame signature
        assert one.device.type == "cpu" and one.dtype == torch.float32 and tuple(one.shape) == (73, 26)
        assert np.array_equal(one.numpy(), gold[tag + "_scaled"][4])
    # a 60 s clip's worth of frames as one contiguous range == the oracle per frame
    table = osem.synth_table(1497, seed=5)
    ratio = osem.find_crop_norm_ratio(table[0:1], table)
    got = iu.semantic_windows(iu.upload_semantic(table, "cuda"), (0, 1497), ratio).cpu().numpy()
    for i in (0, 1, 12, 13, 700, 1483, 1484, 1496):
        assert np.array_equal(got[i], osem.transform_semantic(table, i, ratio)), i
    assert iu.semantic_windows(iu.upload_semantic(table, "cuda"), (0, 0)).shape == (0, 73, 26)
    with pytest.raises(Exception):
        iu.semantic_windows(torch.from_numpy(table), (0, 1))                     # CPU tensor: no CPU path


@pytest.mark.parametrize("n,c,h,w,reflect", [(3, 64, 24, 24, 0), (2, 128, 10, 14, 1), (1, 256, 8, 8, 0)])
def test_double_layernorm_add(G, n, c, h, w, reflect):
    """s2v_affine_act2: lrelu(LN2d(x)) + lrelu(LN2d(r)) in one pass (decoder up + jump branches, models/LNet.py:66-72)."""
    lib, L = G.lib(), G.L
    outs = []
    xs = [_norm_inputs(n, c, h, w)[1] for _ in range(2)]
    ab, ref = [], 0.0
    for xh in xs:
        xf = xh.permute(0, 3, 1, 2).float()
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1
        ref = ref + F.leaky_relu(F.layer_norm(xf, xf.shape[1:], gamma[:, None, None].expand(c, h, w), beta[:, None, None].expand(c, h, w)), 0.1)
        chunks = G.ops.stats_chunks(n, h * w, c)
        partial = torch.empty(n, chunks, c, 2, device="cuda")
        a, b = torch.empty(n, c, device="cuda"), torch.empty(n, c, device="cuda")
        G.ops.op_chan_stats(lib, xh, chunks, partial).run()
        G.ops.op_ln2d_finalize(lib, partial, n, chunks, c, h * w, gamma, beta, a, b).run()
        ab.append((a, b))
    yp = torch.zeros(n, h + 2, w + 2, c, dtype=torch.float16, device="cuda")
    y = yp[:, 1:-1, 1:-1, :]
    G.ops.op_affine_act(lib, xs[0], ab[0][0], ab[0][1], y, act=L.ACT_LRELU, act_param=0.1, res=xs[1], res_ab=ab[1], reflect1=reflect).run()
    m, rel = G.report("double LN2d add n%d c%d %dx%d" % (n, c, h, w), G.nchw(y), ref)
    assert rel < 2e-3
    if reflect:
        assert torch.equal(G.nchw(yp), F.pad(G.nchw(y), (1, 1, 1, 1), mode="reflect"))


def test_laplacian_blend_vs_reference_golden_and_oracle(G):
    """s2v_pyrdown_u8 / _f32 / s2v_lap_blend_level (futils/inference_utils.py:181-222) against outputs of the reference's own
    function on the real cv2 (tests/golden/blend_golden.npz) and the oracle: 8-bit pyramids bit-exact, blends within 1e-3
    on the 0..255 scale (float32 summation order)."""
    import os
    from oracle import blend as ob
    from s2v_b200.futils import inference_utils as iu
    lib = G.lib()
    gold = np.load(os.path.join(GOLDEN, "blend_golden.npz"))
    import ctypes as C
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(7):                       # cv2.pyrDown itself, edge shapes (1 x 1, 2 x 1, odd sizes)
        x = torch.from_numpy(gold[f"down_u8_{i}_in"]).cuda()[None].contiguous()
        h, w = x.shape[1:3]
        y = torch.empty(1, (h + 1) // 2, (w + 1) // 2, 3, dtype=torch.uint8, device="cuda")
        G.L.check(lib.s2v_pyrdown_u8(x.data_ptr(), 1, h, w, 3, y.data_ptr(), st))
        assert np.array_equal(y[0].cpu().numpy(), gold[f"down_u8_{i}"]), i
        f = torch.from_numpy(gold[f"down_f32_{i}_in"]).cuda()[None].contiguous()
        g = torch.empty(1, (h + 1) // 2, (w + 1) // 2, device="cuda")
        G.L.check(lib.s2v_pyrdown_f32(f.data_ptr(), 1, h, w, 1, g.data_ptr(), st))
        assert np.abs(g[0].cpu().numpy() - gold[f"down_f32_{i}"]).max() <= 1e-6, i
    for key, hw, seed, levels in (("blend64_l6", (64, 64), 0, 6), ("blend64_l7", (64, 64), 0, 7), ("blend48x80_l4", (48, 80), 2, 4)):
        A, B, m = ob.synth_images(*hw, seed=seed)
        got = iu.Laplacian_Pyramid_Blending_with_mask(A, B, m, levels)              # the reference's signature
        assert got.dtype == np.float32 and got.shape == gold[key].shape
        d = np.abs(got - gold[key]).max()
        G.report("laplacian blend %s" % key, torch.from_numpy(got), torch.from_numpy(gold[key]))
        assert d <= 1e-3, (key, d)
    # the call of inference.py:312 (512 x 512, 10 levels), batched: frame 0 against the golden rows, every frame against the oracle
    A, B, m = ob.synth_images(512, 512, seed=1)
    A2, B2, m2 = ob.synth_images(512, 512, seed=4)
    out = iu.laplacian_blend(torch.from_numpy(np.stack([A, A2])).cuda(), torch.from_numpy(np.stack([B, B2])).cuda(),
                             torch.from_numpy(np.stack([m, m2])).cuda(), 10).cpu().numpy()
    assert np.abs(out[0][::37] - gold["blend512_l10_rows"]).max() <= 1e-3
    assert np.abs(out[1] - ob.laplacian_pyramid_blending_with_mask(A2, B2, m2, 10)).max() <= 1e-3
    # properties at full size: blending an image with itself returns it; mask 1 -> A, mask 0 -> B
    dev = lambda a: torch.from_numpy(a).cuda()[None]
    assert (iu.laplacian_blend(dev(A), dev(A), dev(m), 10)[0].cpu() - torch.from_numpy(A).float()).abs().max().item() <= 2e-3
    assert (iu.laplacian_blend(dev(A), dev(B), torch.ones_like(dev(m)), 10)[0].cpu() - torch.from_numpy(A).float()).abs().max().item() <= 2e-3
    assert (iu.laplacian_blend(dev(A), dev(B), torch.zeros_like(dev(m)), 10)[0].cpu() - torch.from_numpy(B).float()).abs().max().item() <= 2e-3
    assert iu.laplacian_blend(dev(A)[:0], dev(B)[:0], dev(m)[:0], 10).shape == (0, 512, 512, 3)
    with pytest.raises(ValueError):
        iu.laplacian_blend(dev(A)[:, :500], dev(B)[:, :500], dev(m)[:, :500], 10)       # 500 is not a multiple of 512


@pytest.mark.parametrize("n,c,h,w,oh,ow", [(2, 64, 24, 24, 48, 48), (3, 8, 17, 9, 8, 4), (1, 256, 100, 100, 200, 200), (2, 8, 96, 96, 256, 256), (1, 16, 8, 8, 4, 4)])
def test_resize_bilinear(G, n, c, h, w, oh, ow):
    """s2v_resize_bilinear == F.interpolate(mode='bilinear', align_corners=False) on fp16 channels-last tensors, with and
    without the per-(n, c) scale (ENet building block: models/base_blocks.py:42-46,500-503)."""
    import ctypes as C
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(5)
    x = torch.randn(n, h, w, c, device="cuda").half()
    scale = torch.rand(n, c, device="cuda") + 0.5
    for sc in (None, scale):
        y = torch.full((n, oh, ow, c), float("nan"), device="cuda", dtype=torch.float16)
        vx, vy = ops.view(x), ops.view(y)
        L.check(lib.s2v_resize_bilinear(C.byref(vx), C.byref(vy), None if sc is None else sc.data_ptr(), 0, ops.cur_stream()))
        ref = F.interpolate(x.permute(0, 3, 1, 2).float(), size=(oh, ow), mode="bilinear", align_corners=False)
        if sc is not None:
            ref = ref * sc[:, :, None, None]
        m, rel = G.report("resize %dx%d->%dx%d c%d" % (h, w, oh, ow, c), G.nchw(y), ref)
        assert rel < 2e-3


def test_mel_vs_reference_audio_py_goldens(G):
    """CUDA mel against outputs of the UNMODIFIED reference futils/audio.py (oracle/make_golden_mel.py, stub librosa): the seeded
    synthetic wav and 4 s of the reference's own speech sample, whose pauses sit on the -4 clip rail.  Gate: 1e-4 * max(|b|, 1)."""
    import os
    from oracle import resample, synth
    from s2v_b200.futils import audio
    g = np.load(os.path.join(GOLDEN, "mel_ref_golden.npz"))
    for name, wav, ref in (("synthetic 1 s", synth.wav(1.0, seed=0), g["mel_synth_seed0_1s"]),
                           ("speech 4 s", resample.pcm_to_float_mono(g["speech_pcm"]), g["mel_speech"])):
        got = audio.melspectrogram(wav)
        assert got.shape == ref.shape and got.dtype == np.float64
        d = np.abs(got - ref)
        tol = 1e-4 * np.maximum(np.abs(ref), 1.0)
        bad = int((d > tol).sum())
        print("mel vs reference audio.py, %s: max_abs=%.3e, %d of %d bins above tolerance, on the -4 rail: ref %.3f got %.3f"
              % (name, d.max(), bad, d.size, (ref <= -4).mean(), (got <= -4).mean()))
        with open("gpurun_out/parity_report.txt", "a") as f:
            f.write("mel vs reference audio.py goldens (%s): max_abs %.3e, %d / %d bins above 1e-4*max(|b|,1)\n" % (name, d.max(), bad, d.size))
        assert bad == 0


def test_two_devices_in_one_process(G):
    """Kernel attributes (opt-in shared memory > 48 KB) and the SM count are per-device state: the same process must be able to
    run the large-smem kernels on a second GPU (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from oracle import synth, weights
    from s2v_b200.models.LNet import LNet
    sd = weights.make_state_dict("lnet", 0)
    mel, face = synth.lnet_inputs(8, seed=0)
    outs = []
    for d in (0, 1):
        net = LNet().to("cuda:%d" % d).eval()
        net.load_state_dict(sd, strict=True)
        outs.append(net(mel.to("cuda:%d" % d), face.to("cuda:%d" % d)).cpu())      # current device stays cuda:0 throughout
    assert torch.equal(outs[0], outs[1])


def test_load_wav_decode_and_resampler_vs_oracle(G):
    """load_wav (futils/audio.py:9-10) behind the file read: PCM decode + mono mix bit-exact, the kaiser_best resampler against the
    numpy restatement of resampy (oracle/resample.py; float64 accumulation in the same order: bit-exact), librosa's length rule,
    and the reference's own sample read from a real wav file."""
    import os
    from oracle import resample as R
    from s2v_b200.futils import audio
    g = np.load(os.path.join(GOLDEN, "mel_ref_golden.npz"))
    pcm = g["speech_pcm"]                                                  # int16 mono, 16 kHz
    got = audio.load_array_device(pcm, 16000, 16000).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, R.pcm_to_float_mono(pcm))
    rng = np.random.default_rng(0)
    stereo = rng.integers(-32768, 32767, size=(5000, 2), dtype=np.int16)
    assert np.array_equal(audio.load_array_device(stereo, 16000, 16000).cpu().numpy(), R.pcm_to_float_mono(stereo))
    u8 = rng.integers(0, 255, size=3000, dtype=np.uint8)
    assert np.array_equal(audio.load_array_device(u8, 16000, 16000).cpu().numpy(), R.pcm_to_float_mono(u8))
    for sr0 in (44100, 48000, 22050, 8000, 11025):
        x = pcm[: sr0 // 2 + 7]                                           # treat the samples as if recorded at sr0
        ref = R.load_array(x, sr0, 16000)
        got = audio.load_array_device(x, sr0, 16000).cpu().numpy()
        assert got.shape == ref.shape == (int(np.ceil(len(x) * 16000.0 / sr0)),)
        d = np.abs(got - ref).max()
        print("resample %d -> 16000: max_abs %.3e" % (sr0, d))
        assert d <= 1e-7
    with pytest.raises(ValueError):
        audio.resample_device(torch.zeros(1, device="cuda"), 48000, 16000)
    # through a file, like the reference's call
    import tempfile
    from scipy.io import wavfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "a.wav")
        wavfile.write(path, 22050, stereo)
        w = audio.load_wav(path, 16000)
        assert w.dtype == np.float32 and np.abs(w - R.load_array(stereo, 22050, 16000)).max() <= 1e-7


def test_frame_io_vs_reference_lines(G):
    """The per-frame image glue (frame_io.py) against the reference's own lines executed with the real cv2
    (tests/golden/imageops_golden.npz): bit-exact 8-bit resize / paste / batch formation / uint8 conversion; the float resize and
    the blend round trip within one uint8 LSB."""
    import os
    from oracle import imageops as oio
    from s2v_b200 import frame_io as fio
    gold = np.load(os.path.join(GOLDEN, "imageops_golden.npz"))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    i = 0
    while f"rs_u8_{i}_in" in gold:
        ref = gold[f"rs_u8_{i}"]
        oh, ow = ref.shape[:2]
        assert np.array_equal(fio.resize_u8(dev(gold[f"rs_u8_{i}_in"])[None], oh, ow)[0].cpu().numpy(), ref), i
        f = gold[f"rs_f32_{i}_in"]
        assert np.abs(fio.resize_f32(dev(f)[None], oh, ow)[0].cpu().numpy() - gold[f"rs_f32_{i}"]).max() <= 1e-4, i
        assert np.abs(fio.resize_f32(dev(f[:, :, :1])[None], oh, ow)[0, :, :, 0].cpu().numpy() - gold[f"rs_f32c1_{i}"]).max() <= 1e-4, i
        i += 1
    # full-size shapes of the path, against the oracle (itself bit-exact against cv2)
    rng = np.random.default_rng(2)
    big = rng.integers(0, 256, (2, 384, 384, 3), dtype=np.uint8)
    for oh, ow in ((211, 187), (192, 192), (512, 512)):
        got = fio.resize_u8(dev(big), oh, ow).cpu().numpy()
        assert all(np.array_equal(got[k], oio.resize_linear_u8(big[k], oh, ow)) for k in range(2))
    assert np.array_equal(fio.fake_to_bgr_u8(dev(gold["fake_in"]))[0].cpu().numpy(), gold["fake_bgr"])
    of = torch.stack([fio.resize_u8(dev(gold[f"oface_{k}"])[None], 48, 48)[0] for k in range(3)])
    fa = torch.stack([fio.resize_u8(dev(gold[f"face_{k}"])[None], 48, 48)[0] for k in range(3)])
    ib, orig = fio.face_batch(of, fa)
    assert np.array_equal(ib.cpu().numpy(), gold["img_batch"]) and np.array_equal(orig.cpu().numpy(), gold["img_original"])
    pc = fio.compose_pred_u8(dev(gold["pred_in"]), ib, orig)
    assert np.array_equal(pc.cpu().numpy(), gold["pred_u8_composed"])
    assert np.array_equal(fio.compose_pred_u8(dev(gold["pred_in"])).cpu().numpy(), gold["pred_u8_plain"])
    # paste: three frames, three different boxes, one launch
    frame = gold["frame_in"]
    boxes = [tuple(int(v) for v in gold["box"]), (0, 90, 0, 120), (5, 29, 100, 118)]
    ff = fio.paste_faces(pc, dev(np.stack([frame] * 3)), boxes).cpu().numpy()
    assert np.array_equal(ff[0], gold["frame_pasted"])
    for k in (1, 2):
        assert np.array_equal(ff[k], oio.paste_resized(gold["pred_u8_composed"][k], frame, boxes[k]))
    with pytest.raises(ValueError):
        fio.paste_faces(pc, dev(np.stack([frame] * 3)), [(0, 91, 0, 10)] * 3)
    # inference.py:308-313
    pp = fio.blend_paste_back(dev(gold["restored_in"])[None], dev(gold["frame_pasted"])[None], dev(gold["mouse_mask_in"])[None], 10)[0].cpu().numpy()
    d = np.abs(pp.astype(np.int32) - gold["blend_back"].astype(np.int32))
    print("blend_paste_back vs reference lines: max |diff| %d LSB, %.4f %% of the bytes differ" % (d.max(), 100 * (d > 0).mean()))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
