"""Drop-in for the reference's models/DNet.py: ``DNet().forward(input_image, driving_source, stage=None)``
returns the same dict (``flow_field`` [B,2,64,64], ``warp_image`` [B,3,256,256] and, unless
``stage == 'warp'``, ``fake_image`` [B,3,256,256]) with the reference's 304-tensor state_dict schema,
executed by libs2v's sm_100a kernels.

Layer mapping (reference file:line -> kernel):
  MappingNet        models/DNet.py:30-54      Conv1d k7 / k3 dil3 as 1xk tcgen05 convs on [B,1,T,C]; mean_over_w
  ADAINHourglass    base_blocks.py:195-365    AdaIN = chan_stats + adain_finalize + affine_act(LReLU 0.1);
                                              4x4 s2 convs via TMA element strides; ConvTranspose2d k3 s2 p1 op1
                                              as 4 sub-pixel phase convs writing strided output views
  flow_out          models/DNet.py:77-79      LayerNorm2d + LReLU apply, 7x7 256->2 conv with fp32 NCHW output
  warp              futils/flow_util.py       ONE fused kernel (s2v_flow_warp_f32), fp16 NHWC side output
  EditingNet        models/DNet.py:93-118     stem 6->64 (overlapping-view trick), DownBlock2d x3,
                                              FineADAINResBlock2d (only conv2/norm2: conv1/norm1 are dead code in the
                                              reference, base_blocks.py:174-176), sub-pixel up convs, Jump, 7x7 + tanh
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from . import _schema
from ._engine import EngineBase, sn_fold, up2_phase_weights


def convT_phase_weights(w):
    """ConvTranspose2d(k3, s2, p1, output_padding 1) weight [Cin,Cout,3,3] -> {(p,q): conv weight [Cout,Cin,kh,kw]}
    for output parity (p,q): p=0 uses ky=1 at input row a; p=1 uses ky=2 at row a and ky=0 at row a+1."""
    taps = {0: (1,), 1: (2, 0)}
    out = {}
    for p in (0, 1):
        for q in (0, 1):
            ky, kx = taps[p], taps[q]
            w4 = torch.zeros(w.shape[1], w.shape[0], len(ky), len(kx), dtype=w.dtype, device=w.device)
            for a, y in enumerate(ky):
                for b, x in enumerate(kx):
                    w4[:, :, a, b] = w[:, :, y, x].t()
            out[(p, q)] = w4
    return out


_IO_OF = {"flow_field": "flow", "warp_image": "warp", "fake_image": "fake"}


def s2d_weights(w, p):
    """4x4 stride-2 pad-1 conv [Cout,Cin,4,4] as a stride-1 conv on the row-phase-p view of its input.

    The channels-last input [H,W,Cin] IS, without moving a byte, two tensors Z_p[H/2, W/2, 2*Cin] (row phase p = 0 / 1: base
    offset p*W*Cin, row stride 2*W*Cin, pixel stride 2*Cin, channels = (column phase q, c)).  Input row 2Y-1+ky is row
    Y + (ky - 1 - p) / 2 of Z_p for the two ky of that phase, input column 2X-1+kx is column X + dX of column phase q with
    (dX, q) = (-1, 1), (0, 0), (0, 1), (+1, 0) for kx = 0..3.  So   out = conv_{2x3}(Z_0; pad (0,1)) + conv_{2x3}(Z_1; pad (1,1))
    with the weights below (zero where a (dX, q) pair does not occur): two unit-stride K segments that the conv kernel serves from
    ONE halo patch each instead of 16 strided tap loads per tile (TMA element strides fetch every skipped pixel as well)."""
    co, ci = w.shape[:2]
    out = torch.zeros(co, 2 * ci, 2, 3, dtype=w.dtype, device=w.device)
    ky_of = {0: (1, 3), 1: (0, 2)}[p]                     # tap t = 0, 1 of the phase
    kx_of = {(0, 1): 0, (1, 0): 1, (1, 1): 2, (2, 0): 3}  # (column tap u = dX + 1, column phase q) -> kx
    for t, ky in enumerate(ky_of):
        for (u, q), kx in kx_of.items():
            out[:, q * ci:(q + 1) * ci, t, u] = w[:, :, ky, kx]
    return out


class DNetEngine(EngineBase):
    def __init__(self, sd, device, conv_impl="tc", use_graph=True):
        super().__init__(device, conv_impl, use_graph)
        assert conv_impl == "tc", "DNet engine is built on the tcgen05 conv path"
        self.s2d = {int(v) for v in os.environ.get("S2V_S2D", "0").split(",") if v != ""}   # encoder levels on the s2d-view conv
        sd = {k: v.detach().to(self.fold_dev) for k, v in sd.items()}
        self._pack(sd)
        self.finish_pack()

    def _pack(self, sd):
        f32 = lambda t: t.float().contiguous()
        self.P = {}
        # mapping net: Conv1d weights [Cout,Cin,k] -> 1xk convs
        self.pack_conv("map.first", sd["mapping_net.first.0.weight"].float()[:, :, None, :], sd["mapping_net.first.0.bias"])
        for i in range(3):
            self.pack_conv(f"map.enc{i}", sd[f"mapping_net.encoder{i}.1.weight"].float()[:, :, None, :], sd[f"mapping_net.encoder{i}.1.bias"])
        self.adain_list = []                       # (prefix, C) in MLP-table order

        def reg_adain(p, c):
            self.adain_list.append((p, c))

        h = "warpping_net.hourglass"
        self.pack_conv(h + ".encoder.input_layer", sd[h + ".encoder.input_layer.weight"].float(), sd[h + ".encoder.input_layer.bias"],
                       cin_pad=8, rowtaps=True)
        ngf, img_f = 32, 256
        for i in range(5):
            cin, cout = min(ngf * 2 ** i, img_f), min(ngf * 2 ** (i + 1), img_f)
            p = f"{h}.encoder.encoder{i}"
            self.pack_conv(p + ".conv_0", sd[p + ".conv_0.weight"].float(), sd[p + ".conv_0.bias"])
            if i in self.s2d:
                e = self.pack_conv(p + ".conv_0.s2d", s2d_weights(sd[p + ".conv_0.weight"].float(), 0), sd[p + ".conv_0.bias"])
                e["w"] = torch.cat([e["w"], ops.pack_w_tc(s2d_weights(sd[p + ".conv_0.weight"].float(), 1))], 1).contiguous()
            self.pack_conv(p + ".conv_1", sd[p + ".conv_1.weight"].float(), sd[p + ".conv_1.bias"])
            reg_adain(p + ".norm_0", cin)
            reg_adain(p + ".norm_1", cout)
        for i in (4, 3, 2):
            cin = min(ngf * 2 ** (i + 1), img_f) * (1 if i == 4 else 2)
            cout = min(ngf * 2 ** i, img_f)
            p = f"{h}.decoder.decoder{i}"
            self.pack_conv(p + ".conv_0", sd[p + ".conv_0.weight"].float(), sd[p + ".conv_0.bias"])
            for nm in ("conv_1", "conv_s"):
                for (ph, qh), w4 in convT_phase_weights(sd[f"{p}.{nm}.weight"].float()).items():
                    self.pack_conv(f"{p}.{nm}.ph{ph}{qh}", w4, sd[f"{p}.{nm}.bias"])
            reg_adain(p + ".norm_0", cin)
            reg_adain(p + ".norm_1", cout)
            reg_adain(p + ".norm_s", cin)
        p = "warpping_net.flow_out"
        self.P[p + ".g"], self.P[p + ".b"] = f32(sd[p + ".0.weight"].flatten()), f32(sd[p + ".0.bias"].flatten())
        self.pack_conv(p + ".2", sd[p + ".2.weight"].float(), sd[p + ".2.bias"])
        e = "editing_net"
        p = e + ".encoder.first.model"
        self.pack_conv(p, sn_fold(sd, p + ".0"), sd[p + ".0.bias"], cin_pad=8, rowtaps=True)
        self.P[p + ".g"], self.P[p + ".b"] = f32(sd[p + ".1.weight"].flatten()), f32(sd[p + ".1.bias"].flatten())
        for i in range(3):
            p = f"{e}.encoder.down{i}.model"
            self.pack_conv(p, sn_fold(sd, p + ".0"), sd[p + ".0.bias"])
            self.P[p + ".g"], self.P[p + ".b"] = f32(sd[p + ".1.weight"].flatten()), f32(sd[p + ".1.bias"].flatten())
        for i in (2, 1, 0):
            cin = min(64 * 2 ** (i + 1), 256)
            for b in range(2):
                p = f"{e}.decoder.res{i}.res{b}"
                self.pack_conv(p + ".conv2", sn_fold(sd, p + ".conv2"), sd[p + ".conv2.bias"])      # conv1/norm1: dead in the reference
                reg_adain(p + ".norm2", cin)
            p = f"{e}.decoder.up{i}.model"
            for (ph, qh), w4 in up2_phase_weights(sn_fold(sd, p + ".0")).items():
                self.pack_conv(f"{p}.ph{ph}{qh}", w4, sd[p + ".0.bias"])
            self.P[p + ".g"], self.P[p + ".b"] = f32(sd[p + ".1.weight"].flatten()), f32(sd[p + ".1.bias"].flatten())
            p = f"{e}.decoder.jump{i}.model"
            self.pack_conv(p, sn_fold(sd, p + ".0"), sd[p + ".0.bias"])
            self.P[p + ".g"], self.P[p + ".b"] = f32(sd[p + ".1.weight"].flatten()), f32(sd[p + ".1.bias"].flatten())
        p = e + ".decoder.final.model.0"
        self.pack_conv(p, sn_fold(sd, p), sd[p + ".bias"])
        # AdaIN MLP tables
        shared_w, shared_b, groups, off = [], [], [], 0
        self.gb_off = {}
        for inst, (p, c) in enumerate(self.adain_list):
            shared_w.append(sd[p + ".mlp_shared.0.weight"].float())
            shared_b.append(sd[p + ".mlp_shared.0.bias"].float())
            groups.append((sd[p + ".mlp_gamma.weight"].float().t(), sd[p + ".mlp_gamma.bias"], inst * 128, off))
            groups.append((sd[p + ".mlp_beta.weight"].float().t(), sd[p + ".mlp_beta.bias"], inst * 128, off + c))
            self.gb_off[p] = (off, c)
            off += 2 * c
        self.gb_total, self.n_inst = off, len(self.adain_list)
        self.pack_conv("adain.shared", torch.cat(shared_w, 0)[:, :, None, None], torch.cat(shared_b, 0))
        self.pack_lin_groups("adain.heads", groups)

    # ------------------------------------------------------------------ plan
    def _build(self, B, T, stage):
        def builder(plan, ws):
            lib, buf = self.lib, lambda *a, **k: self.buf(ws, *a, **k)
            img = buf("in.img", (B, 3, 256, 256), torch.float32)
            coeff = buf("in.coeff", (B, 73, 1, T), torch.float32)
            flow = buf("out.flow", (B, 2, 64, 64), torch.float32)
            warp = buf("out.warp", (B, 3, 256, 256), torch.float32)
            ones = buf("const.ones", (B, 256), torch.float32)
            zeros = buf("const.zeros", (B, 256), torch.float32, zero=True)
            ones.fill_(1.0)
            ones._s2v_init = ("fill", 1.0)

            # ---- MappingNet -> z [B,1,1,256] --------------------------------------------------------
            c80 = buf("map.in", (B, 1, T, 80))
            plan.add(ops.op_pack(lib, coeff, c80, 0, 80))
            x = buf("map.x0", (B, 1, T - 6, 256))
            self.conv(plan, "map.first", c80, x, cin_true=73)
            for i in range(3):
                xa = buf(f"map.a{i}", tuple(x.shape))
                plan.add(ops.op_affine_act(lib, x, ones, zeros, xa, act=L.ACT_LRELU, act_param=0.1))
                y = buf(f"map.x{i + 1}", (B, 1, x.shape[2] - 6, 256))
                self.conv(plan, f"map.enc{i}", xa, y, dil=(1, 3), res2=x[:, :, 3:-3, :])
                x = y
            z = buf("map.z", (B, 1, 1, 256))
            plan.add(ops.op_mean_over_w(lib, x, z))
            hidden = buf("adain.hidden", (B, 1, 1, self.n_inst * 128))
            self.conv(plan, "adain.shared", z, hidden, act=L.ACT_RELU)
            gb = buf("adain.gb", (B, self.gb_total), torch.float32)
            hd = self.W["adain.heads"]
            plan.add(ops.op_grouped_linear(lib, hidden, hd["groups"], hd["tiles"], hd["n_tiles"], gb))

            def adain_fin(tag):
                off, c = self.gb_off[tag]
                return ("adain", gb[:, off:off + c], gb[:, off + c:off + 2 * c], gb.stride(0))

            def adain(tag, x, y, act=L.ACT_LRELU, res=None, stats=None):
                off, c = self.gb_off[tag]
                return self.adain(plan, ws, tag, x, gb[:, off:off + c], gb[:, off + c:off + 2 * c], gb.stride(0), y,
                                  act=act, slope=0.1, res=res, stats=stats)

            # ---- WarpingNet: AdaIN hourglass ------------------------------------------------------------
            h = "warpping_net.hourglass"
            # skip tensors live in the upper channel half of the decoder's concat buffers
            cat3 = buf("wd.cat3", (B, 16, 16, 512))      # [decoder4 out | e3]
            cat2 = buf("wd.cat2", (B, 32, 32, 512))      # [decoder3 out | e2]
            cat1 = buf("wd.cat1", (B, 64, 64, 256))      # [decoder2 out | e1]
            x = buf("we.out0", (B, 256, 256, 32))
            # the statistics of every encoder tensor are emitted by the conv that produces it (no chan_stats pass)
            # (the 32-channel stem keeps a separate chan_stats pass: its per-channel epilogue statistics cost +100 us on a 256 us
            # launch - 32-column passes leave 32 lanes per channel chunk, a 5-stage butterfly per tile - vs a 45 us pass)
            st_x = self.stem_conv(plan, ws, h + ".encoder.input_layer", img, x)
            enc_out = {1: cat1[..., 128:], 2: cat2[..., 256:], 3: cat3[..., 256:]}
            ngf, img_f = 32, 256
            for i in range(5):
                cin, cout = min(ngf * 2 ** i, img_f), min(ngf * 2 ** (i + 1), img_f)
                s = 256 >> i
                p = f"{h}.encoder.encoder{i}"
                xa = buf(p + ".a0", (B, s, s, cin))
                adain(p + ".norm_0", x, xa, stats=st_x)
                y0 = buf(p + ".y0", (B, s // 2, s // 2, cout))
                if i in self.s2d:
                    # 4x4 stride-2 conv as two unit-stride 2x3 K segments on the row-phase views of xa (see s2d_weights)
                    z = [torch.as_strided(xa, (B, s // 2, s // 2, 2 * cin), (xa.stride(0), 2 * s * cin, 2 * cin, 1),
                                          xa.storage_offset() + ph * s * cin) for ph in (0, 1)]
                    st = self.conv_stats(plan, ws, p + ".conv_0.s2d", z[0], y0, tag=p + ".conv_0", k=(2, 3), pad=(0, 1), x2=z[1], k2=(2, 3),
                                         pad2=(1, 1), fin=adain_fin(p + ".norm_1"),
                                         alg_flops=2.0 * B * (s // 2) * (s // 2) * cout * 16 * cin)
                else:
                    st = self.conv_stats(plan, ws, p + ".conv_0", xa, y0, stride=(2, 2), pad=(1, 1), fin=adain_fin(p + ".norm_1"))
                ya = buf(p + ".a1", (B, s // 2, s // 2, cout))
                adain(p + ".norm_1", y0, ya, stats=st)
                y1 = enc_out.get(i) if i in enc_out else buf(p + ".y1", (B, s // 2, s // 2, cout))
                # (encoder4's output feeds TWO AdaINs of decoder4 with different gamma / beta: partials only, no in-kernel finalize)
                st_x = self.conv_stats(plan, ws, p + ".conv_1", ya, y1, pad=(1, 1),
                                       fin=adain_fin(f"{h}.encoder.encoder{i + 1}.norm_0") if i < 4 else None)
                x = y1
            for i, dst in ((4, cat3[..., :256]), (3, cat2[..., :256]), (2, cat1[..., :128])):
                cin, s = x.shape[3], x.shape[1]
                cout = dst.shape[3]
                p = f"{h}.decoder.decoder{i}"
                xs_a, x0_a = buf(p + ".as", (B, s, s, cin)), buf(p + ".a0", (B, s, s, cin))
                st = adain(p + ".norm_s", x, xs_a, stats=st_x if i == 4 else None)
                adain(p + ".norm_0", x, x0_a, stats=st)
                for ph in (0, 1):                                # shortcut: ConvTranspose2d phases
                    for qh in (0, 1):
                        self.conv(plan, f"{p}.conv_s.ph{ph}{qh}", xs_a, dst[:, ph::2, qh::2, :])
                d0 = buf(p + ".d0", (B, s, s, cout))
                st = self.conv_stats(plan, ws, p + ".conv_0", x0_a, d0, pad=(1, 1), fin=adain_fin(p + ".norm_1"))
                d0a = buf(p + ".d0a", (B, s, s, cout))
                adain(p + ".norm_1", d0, d0a, stats=st)
                for ph in (0, 1):                                # main branch, accumulated onto the shortcut
                    for qh in (0, 1):
                        v = dst[:, ph::2, qh::2, :]
                        self.conv(plan, f"{p}.conv_1.ph{ph}{qh}", d0a, v, res2=v)
                x = {4: cat3, 3: cat2, 2: cat1}[i]
            p = "warpping_net.flow_out"
            fa = buf("wd.flow_in", (B, 64, 64, 256))
            self.layernorm2d(plan, ws, p, x, self.P[p + ".g"], self.P[p + ".b"], fa)
            self.head_conv(plan, p + ".2", fa, flow)
            io = dict(img=img, coeff=coeff, flow=flow, warp=warp)
            if stage == "warp":
                plan.add(ops.op_flow_warp(lib, img, flow, warp))
                return io

            # ---- EditingNet --------------------------------------------------------------------------
            e = "editing_net"
            fake = buf("out.fake", (B, 3, 256, 256), torch.float32)
            io["fake"] = fake
            xp = buf("ed.in8p", (B, 262, 264, 8), zero=True)          # stem input, padded by 3: [img | warp | 0 0]
            inter = xp[:, 3:259, 3:259, :]
            # torch.cat([input_image, warp_image], 1) (DNet.py:104) as ONE 16-byte texel store per pixel from the warp kernel itself
            plan.add(ops.op_flow_warp(lib, img, flow, warp, inter, 3, pack_src=True))
            win = torch.as_strided(xp, (B, 262, 256, 64), (xp.stride(0), xp.stride(1), 8, 1))
            p = e + ".encoder.first.model"
            raw = buf("ed.raw0", (B, 256, 256, 64))
            st = self.conv_stats(plan, ws, p, win, raw, pad=(0, 0), cin_true=6 * 7, fin=("ln", self.P[p + ".g"], self.P[p + ".b"]))
            f0 = buf("ed.f0", (B, 256, 256, 64))
            self.layernorm2d(plan, ws, p, raw, self.P[p + ".g"], self.P[p + ".b"], f0, stats=st)
            feats, x = [f0], f0
            for i in range(3):
                p = f"{e}.encoder.down{i}.model"
                s, co = 256 >> i, min(128 << i, 256)
                raw = buf(f"ed.raw{i + 1}", (B, s, s, co))
                st = self.conv_stats(plan, ws, p, x, raw, pad=(1, 1), fin=("ln", self.P[p + ".g"], self.P[p + ".b"]))
                y = buf(f"ed.f{i + 1}", (B, s // 2, s // 2, co))
                self.layernorm2d(plan, ws, p, raw, self.P[p + ".g"], self.P[p + ".b"], y, pool2=1, stats=st)
                feats.append(y)
                x = y
            out = feats.pop()
            for i in (2, 1, 0):
                s, c = out.shape[1], out.shape[3]
                for b in range(2):
                    p = f"{e}.decoder.res{i}.res{b}"
                    raw = buf(f"ed.res{i}.raw", (B, s, s, c))
                    st = self.conv_stats(plan, ws, p + ".conv2", out, raw, pad=(1, 1), fin=adain_fin(p + ".norm2"))
                    y = buf(f"ed.res{i}.o{b}", (B, s, s, c))
                    adain(p + ".norm2", raw, y, act=L.ACT_NONE, res=out, stats=st)          # dx + x, no activation (quirk C.3)
                    out = y
                co = min(64 << i, 256)
                p = f"{e}.decoder.up{i}.model"
                uraw = buf(f"ed.up{i}.raw", (B, 2 * s, 2 * s, co))
                for ph in (0, 1):
                    for qh in (0, 1):
                        st = self.conv_stats(plan, ws, f"{p}.ph{ph}{qh}", out, uraw[:, ph::2, qh::2, :], tag=p, phase=2 * ph + qh, phases=4,
                                             pad=(1 - ph, 1 - qh), alg_scale=9.0 / 4.0, fin=("ln", self.P[p + ".g"], self.P[p + ".b"]))
                # up-branch LN + LReLU is applied inside the jump-branch pass (s2v_affine_act2): uact is never materialised
                uab = self.ln2d_scale_shift(plan, ws, p, uraw, self.P[p + ".g"], self.P[p + ".b"], stats=st)
                p = f"{e}.decoder.jump{i}.model"
                jraw = buf(f"ed.jump{i}.raw", (B, 2 * s, 2 * s, co))
                st = self.conv_stats(plan, ws, p, feats.pop(), jraw, pad=(1, 1), fin=("ln", self.P[p + ".g"], self.P[p + ".b"]))
                nxt = buf(f"ed.dec{i}.out", (B, 2 * s, 2 * s, co))
                self.layernorm2d(plan, ws, p, jraw, self.P[p + ".g"], self.P[p + ".b"], nxt, res=uraw, res_ab=uab, stats=st)
                out = nxt
            self.head_conv(plan, e + ".decoder.final.model.0", out, fake, act=L.ACT_TANH)
            return io

        return builder

    def forward(self, img, coeff, stage=None, only=None):
        """``only``: return just that output as a VIEW of the plan's I/O buffer (valid until this engine's next forward; the
        caller consumes it on the same stream) instead of fresh copies of all outputs."""
        B, T = img.shape[0], coeff.shape[2]
        if B == 0:                                  # empty batch: empty outputs with the reference's shapes
            out = {"flow_field": img.new_empty(0, 2, 64, 64), "warp_image": img.new_empty(0, 3, 256, 256)}
            if stage != "warp":
                out["fake_image"] = img.new_empty(0, 3, 256, 256)
            return out
        # Batches that are not a multiple of 8 run on the next multiple's plan (frames are independent, the zeroed padding rows
        # cannot influence the first B outputs): a bounded set of plan sizes for clips of any length, as in LNetEngine.forward.
        Bp = B if (B < 8 or B % 8 == 0) else (B + 7) // 8 * 8
        key = (Bp, T, "warp" if stage == "warp" else "full")
        with self._lock, torch.cuda.device(self.dev):
            self.begin_forward()
            ent = self._get_plan(key, self._build(Bp, T, key[2]))
            io = ent["io"]
            io["img"][:B].copy_(img, non_blocking=True)
            io["coeff"][:B].copy_(coeff.reshape(B, 73, 1, T), non_blocking=True)
            if Bp != B:
                io["img"][B:].zero_()
                io["coeff"][B:].zero_()
            self._run(ent)
            if only is not None:                    # the pipeline needs a single output and no copies of the others
                out = {only: io[_IO_OF[only]][:B]}
            else:
                out = {"flow_field": io["flow"][:B].clone(), "warp_image": io["warp"][:B].clone()}
                if "fake" in io:
                    out["fake_image"] = io["fake"][:B].clone()
            self.end_forward()
        return out


class DNet(nn.Module):
    def __init__(self, conv_impl="tc", use_graph=True):
        super().__init__()
        self._conv_impl, self._use_graph = conv_impl, use_graph
        _schema.build_param_tree(self, _schema.dnet_spec())
        self._engine, self._engine_key = None, None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> DNetEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise L.S2VError("DNet runs on CUDA only (sm_100a kernels, no CPU fallback); call .cuda() first")
        if self._engine is None or self._engine_key != dev:
            from .. import custom_ops
            if getattr(self, "_handle", None):
                custom_ops.release_engine(self._handle)
            self._engine = DNetEngine(self.state_dict(), dev, conv_impl=self._conv_impl, use_graph=self._use_graph)
            self._engine_key = dev
            self._handle = custom_ops.register_engine(self._engine)
        return self._engine

    @torch.no_grad()
    def forward(self, input_image, driving_source, stage=None):
        if self.training:
            raise L.S2VError("this DNet is an inference engine (eval-mode semantics); call .eval()")
        if driving_source.shape[2] < 25:
            raise RuntimeError("driving_source needs at least 25 frames (MappingNet crops 24, models/DNet.py:48-53)")
        self.engine()
        res = torch.ops.s2v.dnet_forward(input_image.float().contiguous(), driving_source.float().contiguous(),
                                         stage == "warp", self._handle)
        out = {"flow_field": res[0], "warp_image": res[1]}
        if stage != "warp":
            out["fake_image"] = res[2]
        return out
