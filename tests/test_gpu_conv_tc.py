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
]


def _act(name, L):
    return {None: L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU, "gelu": L.ACT_GELU}[name]


def _act_ref(name, t):
    if name == "relu":
        return F.relu(t)
    if name == "lrelu":
        return F.leaky_relu(t, 0.1)
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
    xh = G.nhwc(x)
    xr = xh.permute(0, 3, 1, 2).float()
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
        ref = F.conv2d(xr, wr, bias, padding=pad)
    if scale is not None:
        ref = F.conv2d(xr, wr, None, padding=pad) * scale[None, :, None, None] + (bias[None, :, None, None] if bias is not None else 0)
    r1 = r2 = None
    if case.get("res1"):
        r1 = torch.randn(n, h, w, cout, device="cuda").half()
        ref = ref + r1.permute(0, 3, 1, 2).float()
    ref = _act_ref(case.get("act"), ref)
    if case.get("res2"):
        r2 = torch.randn(n, h, w, cout, device="cuda").half()
        ref = ref + r2.permute(0, 3, 1, 2).float()

    kw = dict(k=(k, k), pad=(pad, pad), scale=scale, bias=bias, res1=r1, res2=r2, act=_act(case.get("act"), L), act_param=0.1)
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
        y = torch.zeros(n, h, w, cout, dtype=torch.float16, device="cuda")
        ops.op_conv(lib, xh, ops.pack_w_tc(wt), y, **kw).run()
        torch.cuda.synchronize()
        got = G.nchw(y)
    m, rel = G.report("conv_tc %s" % case["name"], got, ref)
    assert rel < 3e-3
