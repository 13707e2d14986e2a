"""-m gpu parity tests of the tcgen05/TMEM/TMA implicit-GEMM convolution (through the C ABI)
against a plain PyTorch fp32 reference of the same op on fp16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def G():
    import gpu_util
    gpu_util.lib()
    return gpu_util


CASES = [
    # plain GEMM (1x1 conv on a flat token matrix), K = 64: the smallest possible use
    dict(name="gemm_k64", n=1, h=1, w=256, cin=64, cout=64, k=1, pad=0),
    dict(name="gemm_k512_n256", n=1, h=1, w=640, cin=512, cout=256, k=1, pad=0, bias=True, act="gelu"),
    dict(name="gemm_n768_tail", n=1, h=1, w=300, cin=384, cout=768, k=1, pad=0, res2=True),
    # 3x3 zero padding by TMA out-of-bounds fill, every box geometry used by LNet
    dict(name="c3_96", n=2, h=96, w=96, cin=64, cout=128, k=3, pad=1, bias=True),
    dict(name="c3_48", n=2, h=48, w=48, cin=128, cout=256, k=3, pad=1, bias=True),
    dict(name="c3_24", n=3, h=24, w=24, cin=256, cout=512, k=3, pad=1, bias=True),
    dict(name="c3_12_box448", n=11, h=12, w=12, cin=128, cout=128, k=3, pad=1),
    # channel counts that are not multiples of 64 (TMA zero fill along C) and small N tiles
    dict(name="c3_cin32_cout96", n=2, h=48, w=48, cin=32, cout=96, k=3, pad=1),
    dict(name="c1_cin96_cout48", n=2, h=48, w=48, cin=96, cout=48, k=1, pad=0, scale=True, bias=True, act="relu"),
    dict(name="c1_cin48_cout96", n=2, h=24, w=24, cin=48, cout=96, k=1, pad=0),
    dict(name="c3_cin192_cout64", n=2, h=24, w=24, cin=192, cout=64, k=3, pad=1),
    # channel-slice views for x, y and residuals (FFC local/global split), pre-padded reflect input
    dict(name="ffc_like_views", n=9, h=12, w=12, cin=256, cout=256, k=3, pad=1, views=True, reflect=True),
    # asymmetric 2x2 phases of nearest-x2 + conv3x3 (sub-pixel decomposition) into strided outputs
    dict(name="up2_phases", n=2, h=12, w=12, cin=128, cout=64, k=3, pad=1, up2=True),
    dict(name="c3_res1_lrelu", n=2, h=24, w=24, cin=64, cout=64, k=3, pad=1, res1=True, act="lrelu", bias=True),
    # strided convs through TMA element strides (audio encoder models/LNet.py:102-120; DNet 4x4 s2)
    dict(name="s31_audio3", n=3, h=80, w=16, cin=32, cout=64, k=3, pad=1, stride=(3, 1), scale=True, bias=True, act="relu"),
    dict(name="s33_audio6", n=5, h=27, w=16, cin=64, cout=128, k=3, pad=1, stride=(3, 3), bias=True),
    dict(name="s32_audio9", n=70, h=9, w=6, cin=128, cout=256, k=3, pad=1, stride=(3, 2), bias=True),
    dict(name="k4s2_dnet", n=2, h=64, w=64, cin=32, cout=64, k=4, pad=1, stride=(2, 2), bias=True),
    dict(name="audio0_cin1", n=3, h=80, w=16, cin=1, cout=32, k=3, pad=1, bias=True, act="relu"),
    dict(name="audio11_3x3_to_1x1", n=9, h=3, w=3, cin=256, cout=512, k=3, pad=0, bias=True),
    dict(name="mlp_rows", n=37, h=1, w=1, cin=512, cout=1280, k=1, pad=0, bias=True, act="relu"),
    # Cout = 3 head with fp32 NCHW output + sigmoid (FinalBlock2d)
    dict(name="head_7x7_cout3", n=2, h=24, w=24, cin=64, cout=3, k=7, pad=3, bias=True, act="sigmoid", f32=True),
]


def _act(name, L):
    return {None: L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU, "gelu": L.ACT_GELU, "sigmoid": L.ACT_SIGMOID}[name]


def _act_ref(name, t):
    if name == "relu":
        return F.relu(t)
    if name == "lrelu":
        return F.leaky_relu(t, 0.1)
    if name == "sigmoid":
        return torch.sigmoid(t)
    if name == "gelu":
        return 0.5 * t * (1 + torch.tanh(0.7978845608028654 * (t + 0.044715 * t ** 3)))
    return t


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_conv_tc(G, case):
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(7)
    n, h, w, cin, cout, k, pad = (case[x] for x in ("n", "h", "w", "cin", "cout", "k", "pad"))
    x = torch.randn(n, cin, h, w, device="cuda")
    wt = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
    bias = torch.randn(cout, device="cuda") if case.get("bias") else None
    scale = torch.rand(cout, device="cuda") + 0.5 if case.get("scale") else None
    stride = case.get("stride", (1, 1))
    if cin % 8:
        xh = torch.zeros(n, h, w, -(-cin // 8) * 8, dtype=torch.float16, device="cuda")
        xh[..., :cin] = G.nhwc(x)
    else:
        xh = G.nhwc(x)
    xr = xh[..., :cin].permute(0, 3, 1, 2).float()
    wr = wt.half().float()

    if case.get("up2"):
        # reference: nearest x2 then 3x3 zero-pad conv; CUDA: four 2x2 phase convs on the low-res input
        ref = F.conv2d(F.interpolate(xr, scale_factor=2), wr, bias, padding=1)
        y = torch.zeros(n, 2 * h, 2 * w, cout, dtype=torch.float16, device="cuda")
        grp = {0: [[0], [1, 2]], 1: [[0, 1], [2]]}
        for p in (0, 1):
            for q in (0, 1):
                w4 = torch.zeros(cout, cin, 2, 2, device="cuda")
                for a in (0, 1):
                    for b in (0, 1):
                        for ky in grp[p][a]:
                            for kx in grp[q][b]:
                                w4[:, :, a, b] += wt[:, :, ky, kx]
                ops.op_conv(lib, xh, ops.pack_w_tc(w4), y[:, p::2, q::2, :], k=(2, 2), pad=(1 - p, 1 - q), bias=bias).run()
        torch.cuda.synchronize()
        m, rel = G.report("conv_tc %s" % case["name"], G.nchw(y), ref)
        assert rel < 6e-3      # phase weights are summed in fp32 then rounded once -> differs from the fp16-rounded reference weights
        return

    if case.get("reflect"):
        ref = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), wr, bias)
    else:
        ref = F.conv2d(xr, wr, bias, stride=stride, padding=pad)
    if scale is not None:
        ref = F.conv2d(xr, wr, None, stride=stride, padding=pad) * scale[None, :, None, None] + (bias[None, :, None, None] if bias is not None else 0)
    oh, ow = ref.shape[2:]
    r1 = r2 = None
    if case.get("res1"):
        r1 = torch.randn(n, oh, ow, cout, device="cuda").half()
        ref = ref + r1.permute(0, 3, 1, 2).float()
    ref = _act_ref(case.get("act"), ref)
    if case.get("res2"):
        r2 = torch.randn(n, oh, ow, cout, device="cuda").half()
        ref = ref + r2.permute(0, 3, 1, 2).float()

    kw = dict(k=(k, k), stride=stride, pad=(pad, pad), scale=scale, bias=bias, res1=r1, res2=r2, act=_act(case.get("act"), L), act_param=0.1)
    if case.get("f32"):
        yf = torch.zeros(n, cout, oh, ow, device="cuda")
        ops.op_conv(lib, xh, ops.pack_w_tc(wt), None, y_f32=yf, out_shape=(n, cout, oh, ow), **kw).run()
        torch.cuda.synchronize()
        m, rel = G.report("conv_tc %s" % case["name"], yf, ref)
        assert rel < 3e-3
        return
    if case.get("views"):
        # x lives in channels [64, 64+cin) of a reflect-padded buffer; y in channels [8, 8+cout) of a wider one
        big = torch.randn(n, h + 2, w + 2, cin + 128, device="cuda").half()
        big[:, 1:-1, 1:-1, 64:64 + cin] = xh
        ops.op_reflect_border(lib, big[:, 1:-1, 1:-1, :]).run()
        ybig = torch.full((n, h, w, cout + 64), 3.0, dtype=torch.float16, device="cuda")
        yv = ybig[..., 8:8 + cout]
        kw.update(pad=(0, 0))
        ops.op_conv(lib, big[..., 64:64 + cin], ops.pack_w_tc(wt), yv, **kw).run()
        torch.cuda.synchronize()
        assert (ybig[..., :8] == 3).all() and (ybig[..., 8 + cout:] == 3).all()
        got = G.nchw(yv)
    else:
        y = torch.zeros(n, oh, ow, cout, dtype=torch.float16, device="cuda")
        ops.op_conv(lib, xh, ops.pack_w_tc(wt), y, **kw).run()
        torch.cuda.synchronize()
        got = G.nchw(y)
    m, rel = G.report("conv_tc %s" % case["name"], got, ref)
    assert rel < 3e-3


@pytest.mark.parametrize("cin,cout,k,hw", [(3, 64, 7, 96), (6, 64, 7, 64), (3, 32, 7, 40)])
def test_stem_conv_overlapping_view(G, cin, cout, k, hw):
    """Tiny-Cin stems (FirstBlock2d / ADAINEncoder.input_layer): one 64-wide K chunk = k x 8 consecutive
    values of a row of the padded 8-channel buffer, read through an overlapping TMA view."""
    from s2v_b200.models._engine import EngineBase
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(9)
    n = 3
    src = torch.randn(n, cin + 2, hw, hw, device="cuda")[:, 1:1 + cin]        # a channel window, like LNet's face halves
    wt = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
    bias = torch.randn(cout, device="cuda")
    eng = EngineBase(torch.device("cuda", torch.cuda.current_device()))
    eng.pack_conv("stem", wt, bias, cin_pad=8, rowtaps=True)
    plan, ws = ops.Plan(), {}
    y = torch.zeros(n, hw, hw, cout, dtype=torch.float16, device="cuda")
    eng.stem_conv(plan, ws, "stem", src, y, k=k, cin_true=cin)
    plan.run()
    torch.cuda.synchronize()
    ref = F.conv2d(src.half().float(), wt.half().float(), bias, padding=k // 2)
    m, rel = G.report("stem conv cin%d k%d %d" % (cin, k, hw), G.nchw(y), ref)
    assert rel < 3e-3


def test_two_segment_gemm(G):
    """y = conv3x3_reflect(x_l) + conv1x1(s2): both K segments accumulate in one TMEM tile (FFC out_g).  The 24 x 24 and 48 x 48
    geometries run in halo mode with resident weights and few A-patch slots: the segment-2 patches get a ring of their own
    (conv_tc.cu: a_slots2), with and without a per-channel scale / bias / activation epilogue (table space changes the slot plan)."""
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(11)
    for (n, s, cl, cg, ch, epi) in ((9, 12, 64, 192, 96, False), (3, 24, 64, 192, 96, False), (3, 24, 64, 192, 96, True),
                                    (2, 48, 32, 96, 48, False), (5, 48, 32, 96, 48, True)):
        xp = torch.randn(n, s + 2, s + 2, cl + cg, device="cuda").half()
        ops.op_reflect_border(lib, xp[:, 1:-1, 1:-1, :]).run()
        s2 = torch.randn(n, s, s, ch, device="cuda").half()
        w3 = torch.randn(cg, cl, 3, 3, device="cuda") / (cl * 9) ** 0.5
        w1 = torch.randn(cg, ch, 1, 1, device="cuda") / ch ** 0.5
        R = torch.full((n, s, s, cl + cg), 5.0, dtype=torch.float16, device="cuda")
        wcat = torch.cat([ops.pack_w_tc(w3), ops.pack_w_tc(w1)], 1).contiguous()
        kw = dict(scale=torch.rand(cg, device="cuda") + 0.5, bias=torch.randn(cg, device="cuda"), act=L.ACT_LRELU, act_param=0.1) if epi else {}
        ops.op_conv(lib, xp[..., :cl], wcat, R[..., cl:], k=(3, 3), x2=s2, **kw).run()
        torch.cuda.synchronize()
        xl = xp[:, 1:-1, 1:-1, :cl].permute(0, 3, 1, 2).float()
        ref = F.conv2d(F.pad(xl, (1, 1, 1, 1), mode="reflect"), w3.half().float()) + F.conv2d(s2.permute(0, 3, 1, 2).float(), w1.half().float())
        if epi:
            ref = F.leaky_relu(ref * kw["scale"][None, :, None, None] + kw["bias"][None, :, None, None], 0.1)
        m, rel = G.report("conv_tc two-segment (3x3 + 1x1) %dx%d%s" % (s, s, " + epilogue" if epi else ""), G.nchw(R[..., cl:]), ref)
        assert rel < 3e-3 and (R[..., :cl] == 5).all(), (n, s, epi, rel)


def test_fused_epilogue_statistics(G):
    """conv_tc's optional fused per-(image, tile, channel) sums feed the same finalize kernels as chan_stats."""
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(12)
    for (n, s, cin, cout) in ((9, 12, 128, 256), (3, 48, 64, 32), (2, 24, 64, 192), (3, 24, 128, 64), (2, 96, 64, 128), (10, 12, 64, 768)):
        x = torch.randn(n, s, s, cin, device="cuda").half()
        wt = torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
        y = torch.zeros(n, s, s, cout, dtype=torch.float16, device="cuda")
        assert ops.stats_fusable(lib, cout)
        tiles = ops.box_tiles(s, s, n, (3, 3))
        partial = torch.full((n, tiles, cout, 2), float("nan"), device="cuda")       # every entry must be written
        ops.op_conv(lib, x, ops.pack_w_tc(wt), y, k=(3, 3), pad=(1, 1), stats=(partial, 0, 0)).run()
        torch.cuda.synchronize()
        yf = y.float()
        ref_s, ref_q = yf.sum((1, 2)), (yf * yf).sum((1, 2))
        got = partial.sum(1)
        assert (got[..., 0] - ref_s).abs().max().item() < 2e-2 * max(1.0, ref_s.abs().max().item())
        assert (got[..., 1] - ref_q).abs().max().item() < 1e-3 * ref_q.abs().max().item()


def test_narrow_zero_block_hint(G):
    """s2v_conv.narrow_*: K chunks whose weights are zero outside the first `narrow_cout` rows run as narrower MMAs
    (merged FFC GEMM at 48x48) - same result as the plain launch, with and without fused statistics."""
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(21)
    n, s, c, cl, ch = 5, 48, 128, 32, 48
    xp = torch.randn(n, s + 2, s + 2, c, device="cuda").half()
    s2 = torch.randn(n, s, s, ch, device="cuda").half()
    w3 = torch.randn(c, c, 3, 3, device="cuda") / (c * 9) ** 0.5
    w3[cl:, cl:] = 0                       # global inputs reach only the local outputs
    w1 = torch.randn(c, ch, 1, 1, device="cuda") / ch ** 0.5
    w1[:cl] = 0
    wcat = torch.cat([ops.pack_w_tc(w3), ops.pack_w_tc(w1)], 1).contiguous()
    outs = []
    for narrow in (None, (64, cl)):
        R = torch.zeros(n, s, s, c, dtype=torch.float16, device="cuda")
        tiles = ops.box_tiles(s, s, n, (3, 3))
        partial = torch.full((n, tiles, c, 2), float("nan"), device="cuda")
        ops.op_conv(lib, xp, wcat, R, k=(3, 3), x2=s2, narrow=narrow, stats=(partial, 0, 0)).run()
        torch.cuda.synchronize()
        outs.append((R, partial))
    ref = F.conv2d(xp.permute(0, 3, 1, 2).float(), w3.half().float()) + F.conv2d(s2.permute(0, 3, 1, 2).float(), w1.half().float())
    m, rel = G.report("conv_tc merged FFC with narrow hint", G.nchw(outs[1][0]), ref)
    assert rel < 3e-3
    assert (outs[0][0].float() - outs[1][0].float()).abs().max().item() < 2e-3 * ref.abs().max().item()
    yf = outs[1][0].float()
    assert (outs[1][1].sum(1)[..., 0] - yf.sum((1, 2))).abs().max().item() < 2e-2 * max(1.0, yf.sum((1, 2)).abs().max().item())


def test_conv_head_7x7_folded_taps(G):
    """s2v_conv_head: 7x7 heads with Cout <= 8, kx taps folded into N (FinalBlock2d of LNet / DNet, flow_out)."""
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(31)
    for (n, h, w_, cin, cout, act) in ((3, 96, 96, 64, 3, L.ACT_SIGMOID), (2, 64, 64, 256, 2, L.ACT_NONE),
                                       (2, 50, 37, 64, 3, L.ACT_TANH), (1, 8, 8, 128, 8, L.ACT_NONE), (2, 33, 70, 64, 1, L.ACT_RELU),
                                       (1, 40, 27, 128, 5, L.ACT_LRELU), (2, 64, 64, 64, 4, L.ACT_NONE)):
        x = torch.randn(n, h, w_, cin, device="cuda").half()
        wt = torch.randn(cout, cin, 7, 7, device="cuda") / (cin * 49) ** 0.5
        b = torch.randn(cout, device="cuda")
        y = torch.full((n, cout, h, w_), float("nan"), device="cuda")
        ops.op_conv_head(lib, x, ops.pack_w_head(wt), y, bias=b, act=act, act_param=0.1).run()
        torch.cuda.synchronize()
        ref = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.half().float(), b, padding=3)
        ref = {L.ACT_SIGMOID: torch.sigmoid, L.ACT_TANH: torch.tanh, L.ACT_RELU: torch.relu,
               L.ACT_LRELU: lambda t: F.leaky_relu(t, 0.1), L.ACT_NONE: lambda t: t}[act](ref)
        m, rel = G.report("conv_head %dx%d cin%d cout%d" % (h, w_, cin, cout), y, ref)
        assert rel < 2e-3, (h, w_, cin, cout, rel)


def test_fused_layernorm_totals(G):
    """conv_tc stats_gmax = 0: per-(image, tile, 32-row group) totals over ALL channels from the fp32 accumulators
    (LayerNorm2d consumers), finalized by s2v_ln2d_finalize_totals; against torch on the stored fp16 output."""
    lib, L, ops = G.lib(), G.L, G.ops
    torch.manual_seed(13)
    for (n, h, w_, cin, cout, k) in ((9, 12, 12, 128, 256, 3), (3, 48, 48, 64, 32, 3), (3, 24, 24, 128, 64, 3), (2, 96, 96, 64, 128, 3),
                                      (2, 50, 37, 64, 64, 3), (5, 12, 12, 64, 128, 1), (1, 64, 64, 64, 64, 3)):
        x = torch.randn(n, h, w_, cin, device="cuda").half()
        wt = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
        bias = torch.randn(cout, device="cuda") * 0.2
        y = torch.zeros(n, h, w_, cout, dtype=torch.float16, device="cuda")
        tiles = ops.box_tiles(h, w_, n, (k, k))
        partial = torch.zeros(n, tiles, 4, 2, device="cuda")
        ops.op_conv(lib, x, ops.pack_w_tc(wt), y, k=(k, k), pad=(k // 2, k // 2), bias=bias, stats=(partial, 0, 0, "totals")).run()
        torch.cuda.synchronize()
        ref = F.conv2d(x.permute(0, 3, 1, 2).float(), wt.half().float(), bias, padding=k // 2)
        assert (G.nchw(y) - ref).abs().max().item() < 3e-3 * ref.abs().max().item()
        yf = y.float()
        ref_s, ref_q = yf.sum((1, 2, 3)), (yf * yf).sum((1, 2, 3))
        got = partial.sum((1, 2))
        assert (got[:, 0] - ref_s).abs().max().item() < 2e-2 * max(1.0, ref_s.abs().max().item()) + 2e-4 * yf[0].numel() ** 0.5, (h, cout)
        assert (got[:, 1] - ref_q).abs().max().item() < 1e-3 * ref_q.abs().max().item(), (h, cout)
        # finalize -> the LayerNorm2d scale / shift
        gamma, beta = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.1
        a, b = torch.empty(n, cout, device="cuda"), torch.empty(n, cout, device="cuda")
        ops.op_ln2d_finalize_totals(lib, partial, n, tiles, cout, h * w_, gamma, beta, a, b).run()
        torch.cuda.synchronize()
        mean = yf.mean((1, 2, 3))
        rstd = 1.0 / torch.sqrt(yf.var((1, 2, 3), unbiased=False) + 1e-5)
        assert (a - rstd[:, None] * gamma[None]).abs().max().item() < 2e-3 * a.abs().max().item()
        assert (b - (beta[None] - mean[:, None] * rstd[:, None] * gamma[None])).abs().max().item() < 2e-3 * max(1.0, b.abs().max().item())
