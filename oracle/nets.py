"""Plain-PyTorch functional restatement of the reference nets (TEST INFRASTRUCTURE).

Runs LNet / DNet / flow_util from a reference-schema ``state_dict`` with stock
``torch.nn.functional`` calls only, on whatever device / dtype the state_dict
and inputs live on.  It is the checker the CUDA path is compared with on the
GPU box (where /root/reference does not exist); in the build container it is
pinned against the unmodified reference (tests/test_oracle_vs_reference.py).

Reference citations (relative to /root/reference):
  LNet.forward             models/LNet.py:122-139
  Visual_Encoder           models/LNet.py:10-43
  Decoder                  models/LNet.py:46-77
  audio encoder / Conv2d   models/LNet.py:102-120, models/base_blocks.py:12-26
  Transformer / Attention  models/transformer.py:54-112
  FFC / SpectralTransform / FourierUnit   models/ffc.py:62-233
  LayerNorm2d, First/Down/Up/Jump/Final, ADAIN, FFCResnetBlock,
  FineADAINResBlock2d, ADAINEncoder/Decoder(Block)   models/base_blocks.py:52-457
  DNet / MappingNet / WarpingNet / EditingNet        models/DNet.py:12-118
  convert_flow_to_deformation / warp_image           futils/flow_util.py:3-56
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# shared blocks
# --------------------------------------------------------------------------


def sn_weight(sd, p):
    """Eval-mode spectral norm (torch.nn.utils.spectral_norm): W / (u^T W_mat v)
    with the STORED u, v (no power iteration in eval) - base_blocks.py:72-76."""
    if p + ".weight" in sd:
        return sd[p + ".weight"]
    w = sd[p + ".weight_orig"]
    sigma = torch.dot(sd[p + ".weight_u"], torch.mv(w.flatten(1), sd[p + ".weight_v"]))
    return w / sigma


def layernorm2d(x, sd, p):
    """F.layer_norm over (C,H,W), per-channel affine broadcast (base_blocks.py:52-69)."""
    shp = x.shape[1:]
    return F.layer_norm(x, shp, sd[p + ".weight"].expand(shp), sd[p + ".bias"].expand(shp))


def bn_eval(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def adain(x, z, sd, p):
    """InstanceNorm2d(affine=False) * (1+gamma) + beta (base_blocks.py:127-157)."""
    normalized = F.instance_norm(x, eps=1e-5)
    f = z.reshape(z.size(0), -1)
    actv = F.relu(F.linear(f, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"]))
    gamma = F.linear(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"])
    beta = F.linear(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"])
    return normalized * (1 + gamma[:, :, None, None]) + beta[:, :, None, None]


def first_block(x, sd, p):       # base_blocks.py:79-92
    x = F.conv2d(x, sn_weight(sd, p + ".model.0"), sd[p + ".model.0.bias"], padding=3)
    return F.leaky_relu(layernorm2d(x, sd, p + ".model.1"), 0.1)


def down_block(x, sd, p):        # base_blocks.py:95-109
    x = F.conv2d(x, sn_weight(sd, p + ".model.0"), sd[p + ".model.0.bias"], padding=1)
    return F.avg_pool2d(F.leaky_relu(layernorm2d(x, sd, p + ".model.1"), 0.1), 2)


def up_block(x, sd, p):          # base_blocks.py:112-124 (nearest x2)
    x = F.interpolate(x, scale_factor=2)
    x = F.conv2d(x, sn_weight(sd, p + ".model.0"), sd[p + ".model.0.bias"], padding=1)
    return F.leaky_relu(layernorm2d(x, sd, p + ".model.1"), 0.1)


def jump_block(x, sd, p):        # base_blocks.py:429-441
    x = F.conv2d(x, sn_weight(sd, p + ".model.0"), sd[p + ".model.0.bias"], padding=1)
    return F.leaky_relu(layernorm2d(x, sd, p + ".model.1"), 0.1)


def final_block(x, sd, p, act):  # base_blocks.py:444-457
    x = F.conv2d(x, sn_weight(sd, p + ".model.0"), sd[p + ".model.0.bias"], padding=3)
    return torch.sigmoid(x) if act == "sigmoid" else torch.tanh(x)


# --------------------------------------------------------------------------
# LNet
# --------------------------------------------------------------------------


def _gelu_tanh(x):               # transformer.py:11-15
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * torch.pow(x, 3))))


def transformer(x, y, sd, p, depth=2, heads=4):
    """transformer.py:89-112; q,k from x, v from y (Attention :77-79)."""
    bs, c, h, w = x.shape
    x = x.reshape(bs, c, -1).permute(0, 2, 1)
    y = y.reshape(bs, c, -1).permute(0, 2, 1)
    for l in range(depth):
        a = f"{p}.layers.{l}.0"
        xn = F.layer_norm(x, (c,), sd[a + ".normx.weight"], sd[a + ".normx.bias"])
        yn = F.layer_norm(y, (c,), sd[a + ".normy.weight"], sd[a + ".normy.bias"])
        q = F.linear(xn, sd[a + ".fn.to_q.weight"])
        k = F.linear(xn, sd[a + ".fn.to_k.weight"])
        v = F.linear(yn, sd[a + ".fn.to_v.weight"])
        n = q.shape[1]
        dh = q.shape[2] // heads
        q, k, v = (t.reshape(bs, n, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
        dots = torch.matmul(q, k.transpose(-1, -2)) * (dh ** -0.5)
        out = torch.matmul(dots.softmax(dim=-1), v)
        out = out.permute(0, 2, 1, 3).reshape(bs, n, heads * dh)
        x = F.linear(out, sd[a + ".fn.to_out.0.weight"], sd[a + ".fn.to_out.0.bias"]) + x
        f = f"{p}.layers.{l}.1"
        xn = F.layer_norm(x, (c,), sd[f + ".norm.weight"], sd[f + ".norm.bias"])
        hdn = _gelu_tanh(F.linear(xn, sd[f + ".fn.net.0.weight"], sd[f + ".fn.net.0.bias"]))
        x = F.linear(hdn, sd[f + ".fn.net.3.weight"], sd[f + ".fn.net.3.bias"]) + x
    # the reference re-views [bs, hw, c] as [bs,h,w,c] then permutes (transformer.py:111)
    return x.reshape(bs, h, w, c).permute(0, 3, 1, 2)


_AUDIO_CFG = [  # (stride, padding, residual)  models/LNet.py:102-120
    (1, 1, False), (1, 1, True), (1, 1, True),
    ((3, 1), 1, False), (1, 1, True), (1, 1, True),
    (3, 1, False), (1, 1, True), (1, 1, True),
    ((3, 2), 1, False), (1, 1, True),
    (1, 0, False), (1, 0, False),
]


def audio_encoder(a, sd, p="audio_encoder"):
    x = a
    for i, (stride, pad, res) in enumerate(_AUDIO_CFG):
        q = f"{p}.{i}.conv_block"
        out = bn_eval(F.conv2d(x, sd[q + ".0.weight"], sd[q + ".0.bias"], stride=stride, padding=pad), sd, q + ".1")
        if res:
            out = out + x
        x = F.relu(out)
    return x


def fourier_unit(x, sd, p):      # ffc.py:89-126
    b, c, h, w = x.shape
    ff = torch.fft.rfftn(x, dim=(-2, -1), norm="ortho")
    ff = torch.stack((ff.real, ff.imag), dim=-1).permute(0, 1, 4, 2, 3).reshape(b, 2 * c, h, w // 2 + 1)
    ff = F.relu(bn_eval(F.conv2d(ff, sd[p + ".conv_layer.weight"]), sd, p + ".bn"))
    ff = ff.reshape(b, -1, 2, h, w // 2 + 1).permute(0, 1, 3, 4, 2)
    ff = torch.complex(ff[..., 0].contiguous(), ff[..., 1].contiguous())
    return torch.fft.irfftn(ff, s=(h, w), dim=(-2, -1), norm="ortho")


def spectral_transform(x, sd, p):  # ffc.py:129-173 (stride 1, enable_lfu=False)
    x = F.relu(bn_eval(F.conv2d(x, sd[p + ".conv1.0.weight"]), sd, p + ".conv1.1"))
    out = fourier_unit(x, sd, p + ".fu")
    return F.conv2d(x + out, sd[p + ".conv2.weight"])


def _conv3_reflect(x, w):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)


def ffc(x_l, x_g, sd, p):        # ffc.py:176-233 (not gated, bias False, reflect pad)
    out_l = _conv3_reflect(x_l, sd[p + ".convl2l.weight"]) + _conv3_reflect(x_g, sd[p + ".convg2l.weight"])
    out_g = _conv3_reflect(x_l, sd[p + ".convl2g.weight"]) + spectral_transform(x_g, sd, p + ".convg2g")
    return out_l, out_g


def fine_adain_lama(x_l, x_g, z, sd, p):   # base_blocks.py:368-386; LeakyReLU() default slope 0.01 (quirk C.1)
    x_l, x_g = ffc(x_l, x_g, sd, p + ".ffc")
    x_l = F.leaky_relu(adain(x_l, z, sd, p + ".bn_l"), 0.01)
    x_g = F.leaky_relu(adain(x_g, z, sd, p + ".bn_g"), 0.01)
    return x_l, x_g


def ffc_resnet_block(x, z, sd, p):         # base_blocks.py:389-411
    c = x.shape[1]
    cg = int(c * 0.75)
    x_l, x_g = x[:, :-cg], x[:, -cg:]
    id_l, id_g = x_l, x_g
    x_l, x_g = fine_adain_lama(x_l, x_g, z, sd, p + ".conv1")
    x_l, x_g = fine_adain_lama(x_l, x_g, z, sd, p + ".conv2")
    return torch.cat([id_l + x_l, id_g + x_g], dim=1)


def lnet_encoder(cropped, ref, sd, p="encoder", layers=3):
    x, r = first_block(cropped, sd, p + ".first_inp"), first_block(ref, sd, p + ".first_ref")
    out = [x]
    for i in range(layers):
        x, r = down_block(x, sd, f"{p}.inp_down{i}"), down_block(r, sd, f"{p}.ref_down{i}")
        if i >= 2:
            x = transformer(x, r, sd, f"{p}.ca{i}")
        out.append(x if i < layers - 1 else torch.cat([x, r], dim=1))
    return out


def lnet_decoder(feats, z, sd, p="decoder", layers=3, num_block=9):
    feats = list(feats)
    out = feats.pop()
    for i in range(layers)[::-1]:
        for b in range(num_block):
            out = ffc_resnet_block(out, z, sd, f"{p}.res{i}.res{b}")
        out = up_block(out, sd, f"{p}.up{i}")
        out = jump_block(feats.pop(), sd, f"{p}.jump{i}") + out
    return final_block(out, sd, p + ".final", "sigmoid")


def lnet_forward(sd, audio, face):
    """models/LNet.py:122-139, incl. the 5-D training form."""
    b = audio.size(0)
    five_d = face.dim() > 4
    if five_d:
        audio = torch.cat([audio[:, i] for i in range(audio.size(1))], dim=0)
        face = torch.cat([face[:, :, i] for i in range(face.size(2))], dim=0)
    cropped, ref = torch.split(face, 3, dim=1)
    feats = lnet_encoder(cropped, ref, sd)
    z = audio_encoder(audio, sd)
    out = lnet_decoder(feats, z, sd)
    if five_d:
        out = torch.stack(torch.split(out, b, dim=0), dim=2)
    return out


# --------------------------------------------------------------------------
# flow_util
# --------------------------------------------------------------------------


def convert_flow_to_deformation(flow):     # futils/flow_util.py:3-38
    b, c, h, w = flow.shape
    flow_norm = 2 * torch.cat([flow[:, :1] / (w - 1), flow[:, 1:] / (h - 1)], 1)
    x = 2 * (torch.arange(w).to(flow) / (w - 1)) - 1
    y = 2 * (torch.arange(h).to(flow) / (h - 1)) - 1
    grid = torch.stack([x.view(1, -1).expand(h, w), y.view(-1, 1).expand(h, w)], 2)
    return grid.unsqueeze(0) + flow_norm.permute(0, 2, 3, 1)


def warp_image(source, deformation):       # futils/flow_util.py:41-56
    _, h_old, w_old, _ = deformation.shape
    _, _, h, w = source.shape
    if h_old != h or w_old != w:
        deformation = F.interpolate(deformation.permute(0, 3, 1, 2), size=(h, w), mode="bilinear",
                                    align_corners=False).permute(0, 2, 3, 1)
    return F.grid_sample(source, deformation, mode="bilinear", padding_mode="zeros", align_corners=False)


def warp_closed_form(source, flow):
    """SURVEY A.4 closed form (independent derivation; used to cross-check warp_image)."""
    b, c, H, W = source.shape
    _, _, h, w = flow.shape
    dt = source.dtype
    jx = torch.arange(w, dtype=dt, device=flow.device)
    iy = torch.arange(h, dtype=dt, device=flow.device)
    dx = (2 * jx / (w - 1) - 1)[None, None, :] + 2 * flow[:, 0] / (w - 1)
    dy = (2 * iy / (h - 1) - 1)[None, :, None] + 2 * flow[:, 1] / (h - 1)

    def up(d):
        sy = ((torch.arange(H, dtype=dt, device=d.device) + 0.5) * h / H - 0.5).clamp(min=0)
        sx = ((torch.arange(W, dtype=dt, device=d.device) + 0.5) * w / W - 0.5).clamp(min=0)
        y0 = sy.floor().long(); x0 = sx.floor().long()
        y1 = (y0 + 1).clamp(max=h - 1); x1 = (x0 + 1).clamp(max=w - 1)
        ly = (sy - y0)[None, :, None]; lx = (sx - x0)[None, None, :]
        g = lambda yy, xx: d[:, yy][:, :, xx]
        return (1 - ly) * ((1 - lx) * g(y0, x0) + lx * g(y0, x1)) + ly * ((1 - lx) * g(y1, x0) + lx * g(y1, x1))

    gx, gy = up(dx), up(dy)
    ix = ((gx + 1) * W - 1) / 2
    iy_ = ((gy + 1) * H - 1) / 2
    x0 = ix.floor(); y0 = iy_.floor()
    out = torch.zeros_like(source)
    flat = source.reshape(b, c, H * W)
    for oy, ox in ((0, 0), (0, 1), (1, 0), (1, 1)):
        xx = x0 + ox; yy = y0 + oy
        wgt = (1 - (ix - xx).abs()) * (1 - (iy_ - yy).abs())
        ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().reshape(b, 1, H * W).expand(b, c, H * W)
        out += (torch.gather(flat, 2, idx) * (wgt * ok).reshape(b, 1, H * W)).reshape(b, c, H, W)
    return out


# --------------------------------------------------------------------------
# DNet
# --------------------------------------------------------------------------


def mapping_net(coeff, sd, p="mapping_net", layer=3):   # models/DNet.py:30-54
    out = F.conv1d(coeff, sd[p + ".first.0.weight"], sd[p + ".first.0.bias"])
    for i in range(layer):
        q = f"{p}.encoder{i}.1"
        out = F.conv1d(F.leaky_relu(out, 0.1), sd[q + ".weight"], sd[q + ".bias"], dilation=3) + out[:, :, 3:-3]
    return out.mean(dim=2, keepdim=True)


def _adain_enc_block(x, z, sd, p):         # base_blocks.py:195-212
    x = F.conv2d(F.leaky_relu(adain(x, z, sd, p + ".norm_0"), 0.1), sd[p + ".conv_0.weight"], sd[p + ".conv_0.bias"],
                 stride=2, padding=1)
    x = F.conv2d(F.leaky_relu(adain(x, z, sd, p + ".norm_1"), 0.1), sd[p + ".conv_1.weight"], sd[p + ".conv_1.bias"],
                 padding=1)
    return x


def _adain_dec_block(x, z, sd, p):         # base_blocks.py:215-252 (use_transpose=True)
    ct = dict(stride=2, padding=1, output_padding=1)
    x_s = F.conv_transpose2d(F.leaky_relu(adain(x, z, sd, p + ".norm_s"), 0.1), sd[p + ".conv_s.weight"],
                             sd[p + ".conv_s.bias"], **ct)
    dx = F.conv2d(F.leaky_relu(adain(x, z, sd, p + ".norm_0"), 0.1), sd[p + ".conv_0.weight"], sd[p + ".conv_0.bias"],
                  padding=1)
    dx = F.conv_transpose2d(F.leaky_relu(adain(dx, z, sd, p + ".norm_1"), 0.1), sd[p + ".conv_1.weight"],
                            sd[p + ".conv_1.bias"], **ct)
    return x_s + dx


def warping_net(img, z, sd, p="warpping_net", enc_layers=5, dec_layers=3):   # models/DNet.py:56-90
    h = p + ".hourglass"
    out = F.conv2d(img, sd[h + ".encoder.input_layer.weight"], sd[h + ".encoder.input_layer.bias"], padding=3)
    feats = [out]
    for i in range(enc_layers):
        out = _adain_enc_block(out, z, sd, f"{h}.encoder.encoder{i}")
        feats.append(out)
    out = feats.pop()
    for i in range(enc_layers - dec_layers, enc_layers)[::-1]:
        out = _adain_dec_block(out, z, sd, f"{h}.decoder.decoder{i}")
        out = torch.cat([out, feats.pop()], 1)
    out = F.leaky_relu(layernorm2d(out, sd, p + ".flow_out.0"), 0.1)
    flow = F.conv2d(out, sd[p + ".flow_out.2.weight"], sd[p + ".flow_out.2.bias"], padding=3)
    return {"flow_field": flow, "warp_image": warp_image(img, convert_flow_to_deformation(flow))}


def _fine_adain_resblock(x, z, sd, p):     # base_blocks.py:160-177 (conv1/norm1 result is dead: quirk C.3)
    dx = adain(F.conv2d(x, sn_weight(sd, p + ".conv2"), sd[p + ".conv2.bias"], padding=1), z, sd, p + ".norm2")
    return dx + x


def editing_net(img, warp, z, sd, p="editing_net", layers=3, num_block=2):   # models/DNet.py:93-118
    x = first_block(torch.cat([img, warp], 1), sd, p + ".encoder.first")
    feats = [x]
    for i in range(layers):
        x = down_block(x, sd, f"{p}.encoder.down{i}")
        feats.append(x)
    out = feats.pop()
    for i in range(layers)[::-1]:
        for b in range(num_block):
            out = _fine_adain_resblock(out, z, sd, f"{p}.decoder.res{i}.res{b}")
        out = up_block(out, sd, f"{p}.decoder.up{i}")
        out = jump_block(feats.pop(), sd, f"{p}.decoder.jump{i}") + out
    return final_block(out, sd, p + ".decoder.final", "tanh")


def dnet_forward(sd, input_image, driving_source, stage=None):   # models/DNet.py:20-28
    z = mapping_net(driving_source, sd)
    out = warping_net(input_image, z, sd)
    if stage != "warp":
        out["fake_image"] = editing_net(input_image, out["warp_image"], z, sd)
    return out


def glue_dnet_to_lnet(fake_image):
    """Synthetic DNet->LNet glue of SURVEY 8(d) config 4 (harness convention, applied
    identically in oracle and CUDA path): ref = bilinear((clamp(fake,-1,1)+1)/2 -> 96x96),
    inp = ref with rows 48: zeroed, face = cat(inp, ref)."""
    ref = F.interpolate((fake_image.clamp(-1, 1) + 1) / 2, size=(96, 96), mode="bilinear", align_corners=False)
    inp = ref.clone()
    inp[:, :, 48:] = 0
    return torch.cat([inp, ref], 1)
