"""Drop-in for the reference's models/ENet.py - the 96 -> 384 upsampler that wraps LNet and is the object inference.py:266
really calls: ``ENet(lnet=LNet()).forward(audio_sequences, face_sequences, gt_sequences) -> (outputs, low_res_img)`` with
the reference's state_dict schema (``low_res.*`` = LNet's 1 721 tensors, then ENet's own 64), executed by libs2v's sm_100a
kernels.

Layer mapping (reference file:line -> kernel):
  style encoder     models/ENet.py:93-101, base_blocks.py:29-49   ref -> 256 x 256 (s2v_resize_planes_f32), 1x1 + 6 ResBlocks
                                                                 ('down': conv3x3 + LReLU -> bilinear x0.5 -> conv3x3 + LReLU, + 1x1
                                                                 skip of the halved input) + final conv / linear on s2v_conv_tc;
                                                                 bias + LeakyReLU(0.2) + the skip add live in the conv epilogue
  LNet call         models/ENet.py:103-113                        inputs resized to 96 x 96 in float32, then the LNet plan itself
                                                                 (the same op list LNetEngine builds), in the SAME CUDA graph
  StyleConv         base_blocks.py:487-536                        per-sample modulated 3x3 conv as ONE shared-weight tcgen05 GEMM:
                                                                 conv(x, W*s*d) == d * conv(x*s, W); s = modulation Linear (all six
                                                                 in one s2v_grouped_linear), x*s fused into the bilinear x2 pass or
                                                                 the previous epilogue, d*sqrt2 + noise + bias + LReLU in
                                                                 s2v_style_epilogue
  ToRGB             base_blocks.py:539-553                        s2v_to_rgb: per-pixel 3 x C dot product with the per-sample folded
                                                                 weights + bilinear x2 of the running RGB skip, final crop fused
The StyleConv noise (base_blocks.py:528-530) is drawn from torch's global RNG in the reference's order unless explicit
``noises`` are given (the hook the parity tests use).  Precision: as LNet - fp16 operands / fp32 accumulation, fp16 activations.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops
from . import _schema
from ._engine import EngineBase
from .LNet import LNet, LNetEngine


class ENetEngine(EngineBase):
    def __init__(self, sd, device, lnet_engine: LNetEngine, use_graph=True):
        super().__init__(device, "tc", use_graph)
        self.lnet = lnet_engine
        sd = {k: v.detach().to(self.fold_dev) for k, v in sd.items()}
        self._pack(sd)
        self.finish_pack()

    # ------------------------------------------------------------------ weights
    def _pack(self, sd):
        f32 = lambda t: t.float().contiguous()
        self.P = {}
        self.pack_conv("conv_body_first", sd["conv_body_first.weight"].float(), sd["conv_body_first.bias"])
        self.body = []                              # (cin, cout) per ResBlock
        for i in range(6):
            p = f"conv_body_down.{i}"
            w1, w2 = sd[p + ".conv1.weight"].float(), sd[p + ".conv2.weight"].float()
            self.pack_conv(p + ".conv1", w1, sd[p + ".conv1.bias"])
            self.pack_conv(p + ".conv2", w2, sd[p + ".conv2.bias"])
            self.pack_conv(p + ".skip", sd[p + ".skip.weight"].float())
            self.body.append((w1.shape[1], w2.shape[0]))
        self.pack_conv("final_conv", sd["final_conv.weight"].float(), sd["final_conv.bias"])
        # final_linear consumes feat.reshape(B, -1) of an NCHW [B,512,4,4] tensor (column c*16 + y*4 + x); the activation here is
        # channels-last [B,4,4,512] (column (y*4 + x)*512 + c): permute the weight columns once
        wl = sd["final_linear.weight"].float()
        nf, cfeat = wl.shape[0], wl.shape[1] // 16
        wl = wl.reshape(nf, cfeat, 16).permute(0, 2, 1).reshape(nf, 16 * cfeat)
        self.pack_conv("final_linear", wl[:, :, None, None], sd["final_linear.bias"])
        self.nf = nf
        # modulated convs: shared weights W, w2 = sum over taps of W^2 (demodulation), noise strength, bias
        groups, off = [], 0
        self.mod_off = {}

        def reg_mod(p, cin):
            nonlocal off
            q = p + ".modulated_conv.modulation"
            groups.append((sd[q + ".weight"].float().t(), sd[q + ".bias"], 0, off))
            self.mod_off[p] = (off, cin)
            off += -(-cin // 8) * 8                  # every window starts 32-byte aligned and is >= 8 wide (the 3-channel input is padded to 8)

        self.style = []
        for j in range(4):
            p = f"style_convs.{j}"
            w = sd[p + ".modulated_conv.weight"].float()[0]
            self.pack_conv(p, w)
            self.P[p + ".w2"] = f32((w * w).sum((2, 3)))
            self.P[p + ".nw"] = f32(sd[p + ".weight"].flatten())
            self.P[p + ".bias"] = f32(sd[p + ".bias"].flatten())
            reg_mod(p, w.shape[1])
            self.style.append((w.shape[1], w.shape[0]))
        for j in range(2):
            p = f"to_rgbs.{j}"
            w = sd[p + ".modulated_conv.weight"].float()[0, :, :, 0, 0]
            self.P[p + ".w"] = f32(w)
            self.P[p + ".bias"] = f32(sd[p + ".bias"].flatten())
            reg_mod(p, w.shape[1])
        self.mod_total = off
        self.pack_lin_groups("modulation", groups)

    # ------------------------------------------------------------------ plan
    def _build(self, B, hf, wf, hg, wg):
        def builder(plan, ws):
            lib, buf = self.lib, lambda *a, **k: self.buf(ws, *a, **k)
            f16, f32 = torch.float16, torch.float32
            face_in = buf("enet.in.face", (B, 6, hf, wf), f32)
            gt_in = buf("enet.in.gt", (B, 3, hg, wg), f32)
            noise = [buf(f"enet.in.noise{j}", (B, 1, s, s), f32) for j, s in enumerate((200, 200, 400, 400))]
            out = buf("enet.out", (B, 3, 384, 384), f32)
            # ---- LNet input (models/ENet.py:103-104): cat(inp, gt) -> bilinear 96 x 96, written straight into LNet's input buffer --
            lface = buf("in.face", (B, 6, 96, 96), f32)
            plan.add(ops.op_resize_planes(lib, face_in[:, :3], lface[:, :3]))
            plan.add(ops.op_resize_planes(lib, gt_in, lface[:, 3:]))
            # ---- style encoder on the reference frame (models/ENet.py:93-101) -------------------------------------------------------
            ref256 = buf("enet.ref256", (B, 3, 256, 256), f32)
            plan.add(ops.op_resize_planes(lib, face_in[:, 3:], ref256))
            r8 = buf("enet.ref8", (B, 256, 256, 8))
            plan.add(ops.op_pack(lib, ref256, r8, 0, 8))
            c0 = self.W["conv_body_first"]["cout"]
            f = buf("enet.body.in", (B, 256, 256, c0))
            self.conv(plan, "conv_body_first", r8, f, act=L.ACT_LRELU, act_param=0.2, cin_true=3)
            S = 256
            for i, (cin, cout) in enumerate(self.body):
                p = f"conv_body_down.{i}"
                t = buf(f"enet.body{i}.t", (B, S, S, cin))
                self.conv(plan, p + ".conv1", f, t, pad=(1, 1), act=L.ACT_LRELU, act_param=0.2)
                th, fh = buf(f"enet.body{i}.th", (B, S // 2, S // 2, cin)), buf(f"enet.body{i}.fh", (B, S // 2, S // 2, cin))
                plan.add(ops.op_resize(lib, t, th))                  # bilinear x0.5, align_corners=False == 2x2 mean
                plan.add(ops.op_resize(lib, f, fh))
                sk = buf(f"enet.body{i}.skip", (B, S // 2, S // 2, cout))
                self.conv(plan, p + ".skip", fh, sk)
                nxt = buf(f"enet.body{i}.out", (B, S // 2, S // 2, cout))
                self.conv(plan, p + ".conv2", th, nxt, pad=(1, 1), act=L.ACT_LRELU, act_param=0.2, res2=sk)
                f, S = nxt, S // 2
            feat = buf("enet.feat", (B, 4, 4, f.shape[3]))
            self.conv(plan, "final_conv", f, feat, pad=(1, 1), act=L.ACT_LRELU, act_param=0.2)
            style = buf("enet.style", (B, 1, 1, self.nf))
            self.conv(plan, "final_linear", feat.reshape(B, 1, 1, -1), style)
            mod = buf("enet.mod", (B, self.mod_total), f32, zero=True)
            hd = self.W["modulation"]
            plan.add(ops.op_grouped_linear(lib, style, hd["groups"], hd["tiles"], hd["n_tiles"], mod))
            s_of = lambda p: mod[:, self.mod_off[p][0]:self.mod_off[p][0] + -(-self.mod_off[p][1] // 8) * 8]
            # ---- LNet itself: the op list LNetEngine builds, appended to THIS plan (one graph for the whole ENet forward) --------------
            io_l = self.lnet._build(B)(plan, ws)
            low = io_l["out"]
            # ---- upsampler (models/ENet.py:118-131) -------------------------------------------------------------------------------------
            skip0 = buf("enet.skip0", (B, 3, 100, 100), f32)
            plan.add(ops.op_reflect_pad_nchw(lib, low, 2, skip0))
            x = buf("enet.up.x0", (B, 100, 100, 8))
            plan.add(ops.op_pack(lib, skip0, x, 0, 8))
            rgb, S = skip0, 100
            for lvl in range(2):
                S *= 2
                pa, pb, pr = f"style_convs.{2 * lvl}", f"style_convs.{2 * lvl + 1}", f"to_rgbs.{lvl}"
                (ca, co) = self.style[2 * lvl]
                # StyleConv a: modulate + bilinear x2 in one pass, conv, demodulate + noise + bias + LReLU (and the NEXT conv's modulation)
                u = buf(f"enet.up{lvl}.u", (B, S, S, x.shape[3]))
                plan.add(ops.op_resize(lib, x, u, s_of(pa)))
                ra = buf(f"enet.up{lvl}.a", (B, S, S, co))
                self.conv(plan, pa, u, ra, pad=(1, 1), cin_true=ca)
                da = buf(f"enet.up{lvl}.da", (B, co), f32)
                plan.add(ops.op_style_demod(lib, self.P[pa + ".w2"], s_of(pa), ca, co, math.sqrt(2.0), da))
                plan.add(ops.op_style_epilogue(lib, ra, ra, a=da, bias=self.P[pa + ".bias"], noise=noise[2 * lvl], noise_w=self.P[pa + ".nw"],
                                               post=s_of(pb)))
                # StyleConv b
                rb = buf(f"enet.up{lvl}.b", (B, S, S, co))
                self.conv(plan, pb, ra, rb, pad=(1, 1))
                db = buf(f"enet.up{lvl}.db", (B, co), f32)
                plan.add(ops.op_style_demod(lib, self.P[pb + ".w2"], s_of(pb), co, co, math.sqrt(2.0), db))
                plan.add(ops.op_style_epilogue(lib, rb, rb, a=db, bias=self.P[pb + ".bias"], noise=noise[2 * lvl + 1], noise_w=self.P[pb + ".nw"]))
                # ToRGB + bilinear x2 of the running skip (+ the final crop)
                last = lvl == 1
                nrgb = out if last else buf(f"enet.up{lvl}.rgb", (B, 3, S, S), f32)
                plan.add(ops.op_to_rgb(lib, rb, self.P[pr + ".w"], s_of(pr), self.P[pr + ".bias"], rgb, nrgb, crop=8 if last else 0))
                rgb, x = nrgb, rb
            return dict(mel=io_l["mel"], face=face_in, gt=gt_in, noise=noise, out=out, low=low)

        return builder

    def forward(self, audio, face, gt, noises=None):
        B = audio.shape[0]
        dev = audio.device
        if B == 0:
            return torch.empty(0, 3, 384, 384, device=dev), torch.empty(0, 3, 96, 96, device=dev)
        Bp = B if (B < 8 or B % 8 == 0) else (B + 7) // 8 * 8
        key = (Bp,) + tuple(face.shape[2:]) + tuple(gt.shape[2:])
        with self._lock, torch.cuda.device(self.dev):
            if noises is None:          # the reference's draws (base_blocks.py:528-530), same order, torch's global RNG
                noises = [torch.empty(B, 1, s, s, device=dev).normal_() for s in (200, 200, 400, 400)]
            self.begin_forward()
            ent = self._get_plan(key, self._build(*key))
            io = ent["io"]
            pairs = [(io["mel"], audio), (io["face"], face), (io["gt"], gt)] + list(zip(io["noise"], noises))
            for dst, src in pairs:
                dst[:B].copy_(src, non_blocking=True)
                if Bp != B:
                    dst[B:].zero_()
            self._run(ent)
            res = io["out"][:B].clone(), io["low"][:B].clone()
            self.end_forward()
        return res


class ENet(nn.Module):
    """Same constructor arguments as the reference (models/ENet.py:8-14).  ``concat=True`` (LNet returning features as well) is
    not used by the reference's own loaders and is not supported."""

    def __init__(self, num_style_feat=512, lnet=None, concat=False):
        super().__init__()
        if concat:
            raise L.S2VError("ENet(concat=True) is not supported (the reference never builds it: models/__init__.py:29-35)")
        if lnet is None or not isinstance(lnet, LNet):
            raise TypeError("ENet needs the LNet module it wraps: ENet(lnet=LNet())")
        self.low_res = lnet
        for p in self.low_res.parameters():
            p.requires_grad = False
        self.num_style_feat, self.concat = num_style_feat, concat
        _schema.build_param_tree(self, _schema.enet_spec(num_style_feat), seed=1)
        self._engine, self._engine_key = None, None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    def _invalidate(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def engine(self) -> ENetEngine:
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise L.S2VError("ENet runs on CUDA only (sm_100a kernels, no CPU fallback); call .cuda() first")
        leng = self.low_res.engine()
        if self._engine is None or self._engine_key != dev or self._engine.lnet is not leng:
            own = {k: v for k, v in self.state_dict().items() if not k.startswith("low_res.")}
            self._engine = ENetEngine(own, dev, leng)
            self._engine_key = dev
        return self._engine

    @torch.no_grad()
    def forward(self, audio_sequences, face_sequences, gt_sequences, noises=None):
        if self.training:
            raise L.S2VError("this ENet is an inference engine (eval-mode semantics); call .eval()")
        B = audio_sequences.size(0)
        five_d = face_sequences.dim() > 4
        if five_d:            # time-major flatten, models/ENet.py:87-91
            audio_sequences = torch.cat([audio_sequences[:, i] for i in range(audio_sequences.size(1))], dim=0)
            face_sequences = torch.cat([face_sequences[:, :, i] for i in range(face_sequences.size(2))], dim=0)
            gt_sequences = torch.cat([gt_sequences[:, :, i] for i in range(gt_sequences.size(2))], dim=0)
        out, low = self.engine().forward(audio_sequences.float().contiguous(), face_sequences.float().contiguous(),
                                         gt_sequences.float().contiguous(), noises)
        if five_d:            # models/ENet.py:133-138 (the low-res image is stretched with F.interpolate's default: nearest)
            out = torch.stack(torch.split(out, B, dim=0), dim=2)
            low = torch.nn.functional.interpolate(low, out.shape[3:])
            low = torch.stack(torch.split(low, B, dim=0), dim=2)
        return out, low
