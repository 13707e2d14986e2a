"""state_dict schema (key -> shape/kind) of the reference's LNet and DNet, enumerated from the
architecture constants so that ``load_state_dict`` accepts reference checkpoints unchanged.

Reference constructors: models/LNet.py:10-120, models/DNet.py:30-118, models/base_blocks.py:12-457,
models/ffc.py:62-233, models/transformer.py:24-100.  Key order follows PyTorch's module traversal of
those constructors (parameters, then buffers, then children in registration order), so
``state_dict()`` of the mirrors lists the same keys in the same order as the reference.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

P, B, I = "param", "buffer", "int_buffer"


def _conv(p, cin, cout, k, bias=True, spect=False, k2=None):
    shape = (cout, cin, k, k2 if k2 is not None else k)
    if spect:
        out = [(p + ".bias", (cout,), P)] if bias else []
        return out + [(p + ".weight_orig", shape, P), (p + ".weight_u", (cout,), B),
                      (p + ".weight_v", (cin * shape[2] * shape[3],), B)]
    return [(p + ".weight", shape, P)] + ([(p + ".bias", (cout,), P)] if bias else [])


def _ln2d(p, c):
    return [(p + ".weight", (c, 1, 1), P), (p + ".bias", (c, 1, 1), P)]


def _bn(p, c):
    return [(p + ".weight", (c,), P), (p + ".bias", (c,), P), (p + ".running_mean", (c,), B),
            (p + ".running_var", (c,), B), (p + ".num_batches_tracked", (), I)]


def _linear(p, cin, cout, bias=True):
    return [(p + ".weight", (cout, cin), P)] + ([(p + ".bias", (cout,), P)] if bias else [])


def _adain(p, c, f):
    return _linear(p + ".mlp_shared.0", f, 128) + _linear(p + ".mlp_gamma", 128, c) + _linear(p + ".mlp_beta", 128, c)


def _block(p, cin, cout, k, spect):          # First/Down/Up/Jump: conv + LayerNorm2d
    return _conv(p + ".model.0", cin, cout, k, True, spect) + _ln2d(p + ".model.1", cout)


def _ffc(p, c):
    cg = int(c * 0.75)
    cl = c - cg
    q = p + ".convg2g"
    return (_conv(p + ".convl2l", cl, cl, 3, False) + _conv(p + ".convl2g", cl, cg, 3, False) +
            _conv(p + ".convg2l", cg, cl, 3, False) +
            _conv(q + ".conv1.0", cg, cg // 2, 1, False) + _bn(q + ".conv1.1", cg // 2) +
            _conv(q + ".fu.conv_layer", cg, cg, 1, False) + _bn(q + ".fu.bn", cg) +
            _conv(q + ".conv2", cg // 2, cg, 1, False))


def _lama(p, c, f):
    cg = int(c * 0.75)
    return _ffc(p + ".ffc", c) + _adain(p + ".bn_l", c - cg, f) + _adain(p + ".bn_g", cg, f)


def _transformer(p, dim, depth, heads, dim_head, mlp_dim):
    inner = heads * dim_head
    out = []
    for l in range(depth):
        a, f = f"{p}.layers.{l}.0", f"{p}.layers.{l}.1"
        out += [(a + ".normx.weight", (dim,), P), (a + ".normx.bias", (dim,), P),
                (a + ".normy.weight", (dim,), P), (a + ".normy.bias", (dim,), P)]
        out += _linear(a + ".fn.to_q", dim, inner, False) + _linear(a + ".fn.to_k", dim, inner, False)
        out += _linear(a + ".fn.to_v", dim, inner, False) + _linear(a + ".fn.to_out.0", inner, dim, True)
        out += [(f + ".norm.weight", (dim,), P), (f + ".norm.bias", (dim,), P)]
        out += _linear(f + ".fn.net.0", dim, mlp_dim) + _linear(f + ".fn.net.3", mlp_dim, dim)
    return out


AUDIO_CFG = [  # (cin, cout, k, stride, pad, residual)   models/LNet.py:102-120
    (1, 32, 3, (1, 1), 1, False), (32, 32, 3, (1, 1), 1, True), (32, 32, 3, (1, 1), 1, True),
    (32, 64, 3, (3, 1), 1, False), (64, 64, 3, (1, 1), 1, True), (64, 64, 3, (1, 1), 1, True),
    (64, 128, 3, (3, 3), 1, False), (128, 128, 3, (1, 1), 1, True), (128, 128, 3, (1, 1), 1, True),
    (128, 256, 3, (3, 2), 1, False), (256, 256, 3, (1, 1), 1, True),
    (256, 512, 3, (1, 1), 0, False), (512, 512, 1, (1, 1), 0, False),
]


def lnet_spec(image_nc=3, descriptor_nc=512, layer=3, base_nc=64, max_nc=512, num_res_blocks=9, use_spect=True):
    s = []
    e = "encoder"
    s += _block(e + ".first_inp", image_nc, base_nc, 7, use_spect) + _block(e + ".first_ref", image_nc, base_nc, 7, use_spect)
    for i in range(layer):
        cin, cout = min(base_nc * 2 ** i, max_nc), min(base_nc * 2 ** (i + 1), max_nc)
        if i >= 2:
            s += _transformer(f"{e}.ca{i}", 2 ** (i + 1) * base_nc, 2, 4, base_nc, base_nc * 4)
        s += _block(f"{e}.ref_down{i}", cin, cout, 3, use_spect) + _block(f"{e}.inp_down{i}", cin, cout, 3, use_spect)
    d = "decoder"
    for i in range(layer)[::-1]:
        cin = base_nc * 2 ** (i + 1) * 2 if i == layer - 1 else min(base_nc * 2 ** (i + 1), max_nc)
        cout = min(base_nc * 2 ** i, max_nc)
        s += _block(f"{d}.up{i}", cin, cout, 3, use_spect)
        for b in range(num_res_blocks):
            s += _lama(f"{d}.res{i}.res{b}.conv1", cin, descriptor_nc) + _lama(f"{d}.res{i}.res{b}.conv2", cin, descriptor_nc)
        s += _block(f"{d}.jump{i}", cout, cout, 3, use_spect)
    s += _conv(d + ".final.model.0", base_nc, image_nc, 7, True, use_spect)
    for i, (cin, cout, k, _, _, _) in enumerate(AUDIO_CFG):
        cout = descriptor_nc if i == len(AUDIO_CFG) - 1 else cout
        s += _conv(f"audio_encoder.{i}.conv_block.0", cin, cout, k) + _bn(f"audio_encoder.{i}.conv_block.1", cout)
    return s


def dnet_spec():
    s = []
    s += [("mapping_net.first.0.weight", (256, 73, 7), P), ("mapping_net.first.0.bias", (256,), P)]
    for i in range(3):
        s += [(f"mapping_net.encoder{i}.1.weight", (256, 256, 3), P), (f"mapping_net.encoder{i}.1.bias", (256,), P)]
    h = "warpping_net.hourglass"
    ngf, img_f, f = 32, 256, 256
    s += _conv(h + ".encoder.input_layer", 3, ngf, 7)
    for i in range(5):
        cin, cout = min(ngf * 2 ** i, img_f), min(ngf * 2 ** (i + 1), img_f)
        p = f"{h}.encoder.encoder{i}"
        s += _conv(p + ".conv_0", cin, cout, 4) + _conv(p + ".conv_1", cout, cout, 3)
        s += _adain(p + ".norm_0", cin, f) + _adain(p + ".norm_1", cout, f)
    for i in range(2, 5)[::-1]:
        cin = min(ngf * 2 ** (i + 1), img_f)
        cin = cin * 2 if i != 4 else cin
        cout = min(ngf * 2 ** i, img_f)
        p = f"{h}.decoder.decoder{i}"
        s += _conv(p + ".conv_0", cin, cout, 3)
        # ConvTranspose2d weights are [Cin, Cout, k, k]
        s += [(p + ".conv_1.weight", (cout, cout, 3, 3), P), (p + ".conv_1.bias", (cout,), P)]
        s += [(p + ".conv_s.weight", (cin, cout, 3, 3), P), (p + ".conv_s.bias", (cout,), P)]
        s += _adain(p + ".norm_0", cin, f) + _adain(p + ".norm_1", cout, f) + _adain(p + ".norm_s", cin, f)
    s += _ln2d("warpping_net.flow_out.0", 256) + _conv("warpping_net.flow_out.2", 256, 2, 7)
    e = "editing_net"
    ngf = 64
    s += _block(e + ".encoder.first", 6, ngf, 7, False)
    for i in range(3):
        s += _block(f"{e}.encoder.down{i}", min(ngf * 2 ** i, img_f), min(ngf * 2 ** (i + 1), img_f), 3, False)
    for i in range(3)[::-1]:
        cin, cout = min(ngf * 2 ** (i + 1), img_f), min(ngf * 2 ** i, img_f)
        s += _block(f"{e}.decoder.up{i}", cin, cout, 3, False)
        for b in range(2):
            p = f"{e}.decoder.res{i}.res{b}"
            s += _conv(p + ".conv1", cin, cin, 3) + _conv(p + ".conv2", cin, cin, 3)
            s += _adain(p + ".norm1", cin, f) + _adain(p + ".norm2", cin, f)
        s += _block(f"{e}.decoder.jump{i}", cout, cout, 3, False)
    s += _conv(e + ".decoder.final.model.0", ngf, 3, 7)
    return s


def enet_spec(num_style_feat=512):
    """ENet's own 64 tensors (models/ENet.py:8-80; ``low_res.*`` = the wrapped LNet, registered first)."""
    ch = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128}
    s = _conv("conv_body_first", 3, ch[128], 1)
    cin = ch[128]
    for j, i in enumerate(range(8, 2, -1)):
        cout = ch[2 ** (i - 1)]
        p = f"conv_body_down.{j}"
        s += _conv(p + ".conv1", cin, cin, 3) + _conv(p + ".conv2", cin, cout, 3) + _conv(p + ".skip", cin, cout, 1, bias=False)
        cin = cout
    s += _linear("final_linear", ch[4] * 16, num_style_feat) + _conv("final_conv", cin, ch[4], 3)
    cin, convs = 3, []
    for i in (7, 8):
        cout = ch[2 ** i]
        convs += [(cin, cout), (cout, cout)]
        cin = cout
    for j, (ci, co) in enumerate(convs):
        p = f"style_convs.{j}"
        s += [(p + ".weight", (1,), P), (p + ".bias", (1, co, 1, 1), P), (p + ".modulated_conv.weight", (1, co, ci, 3, 3), P)]
        s += _linear(p + ".modulated_conv.modulation", num_style_feat, ci)
    for j, ci in enumerate((ch[128], ch[256])):
        p = f"to_rgbs.{j}"
        s += [(p + ".bias", (1, 3, 1, 1), P), (p + ".modulated_conv.weight", (1, 3, ci, 1, 1), P)]
        s += _linear(p + ".modulated_conv.modulation", num_style_feat, ci)
    return s


def build_param_tree(root: nn.Module, spec, seed: int = 0) -> None:
    """Registers every tensor of ``spec`` under nested container modules of ``root`` and
    gives it a PyTorch-default-like init (spectral-norm u/v from a short power iteration,
    i.e. the state a trained module holds)."""
    g = torch.Generator().manual_seed(seed)
    shapes = {n: s for n, s, _ in spec}
    for name, shape, kind in spec:
        parts = name.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        leaf = parts[-1]
        if kind == I:
            mod.register_buffer(leaf, torch.tensor(0, dtype=torch.long))
            continue
        if leaf in ("weight", "weight_orig") and len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = (torch.rand(shape, generator=g) * 2 - 1) * fan_in ** -0.5
        elif leaf == "bias":
            wshape = shapes.get(name[:-4] + "weight") or shapes.get(name[:-4] + "weight_orig")
            if wshape is not None and len(wshape) >= 2:
                fan_in = 1
                for d in wshape[1:]:
                    fan_in *= d
                t = (torch.rand(shape, generator=g) * 2 - 1) * fan_in ** -0.5
            else:
                t = torch.zeros(shape)
        elif leaf in ("weight", "running_var"):
            t = torch.ones(shape)
        elif leaf == "running_mean":
            t = torch.zeros(shape)
        elif leaf in ("weight_u", "weight_v"):
            t = torch.zeros(shape)            # filled below
        else:
            raise KeyError(name)
        if kind == P:
            mod.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
        else:
            mod.register_buffer(leaf, t)
    sd = dict(root.state_dict())
    for name, shape, _ in spec:
        if name.endswith(".weight_orig"):
            p = name[: -len(".weight_orig")]
            w = sd[name].flatten(1)
            u = F.normalize(torch.randn(w.shape[0], generator=g), dim=0)
            for _ in range(10):
                v = F.normalize(torch.mv(w.t(), u), dim=0)
                u = F.normalize(torch.mv(w, v), dim=0)
            sd[p + ".weight_u"].copy_(u)
            sd[p + ".weight_v"].copy_(v)
