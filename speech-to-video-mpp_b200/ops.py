"""Thin Python layer over the C ABI: tensor -> s2v_view marshalling and op builders.

Every builder returns an ``Op`` (C function + prepared ctypes arguments + keep-alive
references).  Engines collect Ops into a ``Plan`` once per batch size and replay the
plan (directly, or captured into a CUDA graph); the eager helpers at the bottom run a
single Op on the current stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib as L


def view(t: torch.Tensor) -> L.View:
    """fp16 channels-last [N,H,W,C] tensor (any strides, channel stride 1) -> s2v_view."""
    assert t.dtype == torch.float16 and t.dim() == 4 and (t.stride(3) == 1 or t.shape[3] == 1), (t.dtype, t.shape, t.stride())
    n, h, w, c = t.shape
    # PyTorch leaves arbitrary strides on size-1 dims; give them the canonical (dense) value
    sw = t.stride(2) if w > 1 else c
    sh = t.stride(1) if h > 1 else sw * w
    sn = t.stride(0) if n > 1 else sh * h
    return L.View(t.data_ptr(), n, h, w, c, sn, sh, sw)


def null_view() -> L.View:
    return L.View(None, 0, 0, 0, 0, 0, 0, 0)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class Op:
    name: str
    fn: object
    args: tuple
    keep: tuple = field(default_factory=tuple, repr=False)
    io_bytes: float = 0.0       # tensor-core ops: minimum HBM bytes (operands once + output once), for the per-layer roofline
    alg_bytes: float = 0.0      # algorithmic HBM bytes of a memory-bound op: every input read once + every output written once

    def run(self, stream=None):
        L.check(self.fn(*self.args, stream if stream is not None else cur_stream()), self.name)


class Plan:
    """Ordered list of Ops replayed on one stream, plus an optional SIDE branch: ops added inside ``with plan.side():``
    run on a second stream, forked at the start of the plan and joined where ``plan.join()`` was called.  (LNet: the
    audio encoder and the AdaIN MLPs are a chain of small-grid kernels that only the decoder needs - they overlap the
    visual encoder instead of serialising ~16 launches in front of it.)"""

    def __init__(self):
        self.main_ops: list[Op] = []
        self.side_ops: list[Op] = []
        self._target = self.main_ops
        self.join_at = None
        self._side_stream = None

    @property
    def ops(self) -> list[Op]:
        return self.side_ops + self.main_ops

    def add(self, op: Op) -> Op:
        self._target.append(op)
        return op

    def side(self):
        plan = self

        class _Side:
            def __enter__(self_inner):
                plan._target = plan.side_ops

            def __exit__(self_inner, *a):
                plan._target = plan.main_ops
        return _Side()

    def join(self):
        """Everything added to the main branch from here on may consume the side branch's results."""
        if self.join_at is None:
            self.join_at = len(self.main_ops)

    def run(self, stream=None):
        if not self.side_ops or stream is not None:
            s = stream if stream is not None else cur_stream()
            for op in self.side_ops + self.main_ops:
                rc = op.fn(*op.args, s)
                if rc != 0:
                    L.check(rc, op.name)
            return
        main = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=main.device)
        side = self._side_stream
        fork, joined = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        ss, ms = C.c_void_p(side.cuda_stream), C.c_void_p(main.cuda_stream)
        for op in self.side_ops:
            rc = op.fn(*op.args, ss)
            if rc != 0:
                L.check(rc, op.name)
        joined.record(side)
        join_at = len(self.main_ops) if self.join_at is None else self.join_at
        for i, op in enumerate(self.main_ops):
            if i == join_at:
                main.wait_event(joined)
            rc = op.fn(*op.args, ms)
            if rc != 0:
                L.check(rc, op.name)
        if join_at >= len(self.main_ops):
            main.wait_event(joined)

    def __len__(self):
        return len(self.side_ops) + len(self.main_ops)


def conv_box(h: int, w: int, k=(1, 1), stride=(1, 1), dil=(1, 1)) -> tuple[int, int, int]:
    """Box for a conv_tc launch.  k > 1 unit-stride convs prefer the 8 x 16 single-image box: it enables the
    kernel's halo mode (one A patch per channel chunk serves every tap) unless it wastes > 34 % of the tile."""
    if k[0] * k[1] > 1 and tuple(stride) == (1, 1) and tuple(dil) == (1, 1):
        tiles = -(-w // 8) * -(-h // 16)
        if tiles * 128 / float(h * w) <= 1.34:
            return (8, 16, 1)
    return choose_box(h, w, 1)


def choose_box(h: int, w: int, n: int) -> tuple[int, int, int]:
    """Box of 128 output pixels (box_w, box_h, box_n) with the least padding waste.  The choice is a
    function of the image size ONLY (evaluated for a large batch): the spatial tiling also defines the
    reduction tree of the fused statistics, which must not change with the batch a frame is computed in."""
    n = 4096
    best, best_key = None, None
    for bw in (128, 64, 32, 16, 8, 4, 2, 1):
        for bh in (128, 64, 32, 16, 8, 4, 2, 1):
            if bw * bh > 128:
                continue
            bn = 128 // (bw * bh)
            tiles = -(-w // bw) * -(-h // bh) * -(-n // bn)
            waste = tiles * 128 / float(n * h * w)
            key = (round(waste, 6), bn, -bw)
            if best_key is None or key < best_key:
                best, best_key = (bw, bh, bn), key
    return best


def op_conv(lib, x, w, y, *, k=(1, 1), stride=(1, 1), pad=(0, 0), dil=(1, 1), pad_mode=L.PAD_ZERO, up2=0,
            scale=None, bias=None, res1=None, res2=None, act=L.ACT_NONE, act_param=0.0, y_f32=None,
            out_shape=None, impl="tc", box=None, name="conv", cin_true=None, alg_scale=1.0,
            x2=None, k2=(1, 1), pad2=(0, 0), stats=None, alg_flops=None, narrow=None) -> Op:
    """x: fp16 NHWC tensor; y: fp16 NHWC tensor (or None with y_f32 [N,Cout,OH,OW] float32)."""
    d = L.Conv()
    d.x = view(x)
    cin_true = x.shape[3] if cin_true is None else cin_true
    if y is not None:
        d.y = view(y)
        d.out_mode = L.OUT_F16_NHWC
    else:
        n, co, oh, ow = out_shape
        d.y = L.View(None, n, oh, ow, co, 0, 0, 0)
        d.out_mode = L.OUT_F32_NCHW
        d.y_f32 = y_f32.data_ptr()
    d.w = w.data_ptr()
    d.scale = None if scale is None else scale.data_ptr()
    d.bias = None if bias is None else bias.data_ptr()
    d.res1 = view(res1) if res1 is not None else null_view()
    d.res2 = view(res2) if res2 is not None else null_view()
    d.kh, d.kw = k
    d.stride_h, d.stride_w = stride
    d.pad_h, d.pad_w = pad
    d.dil_h, d.dil_w = dil
    d.pad_mode, d.up2, d.act, d.act_param = pad_mode, up2, act, float(act_param)
    d.x2 = view(x2) if x2 is not None else null_view()
    d.k2h, d.k2w = k2
    d.pad2_h, d.pad2_w = pad2
    if stats is not None:       # (partial [N,chunks,C,2] float32, c_off, chunk_off[, "totals"]): fused output statistics
        partial, c_off, chunk_off = stats[:3]
        d.stats_groups = d.stats_gmax = 1
        if len(stats) > 3 and stats[3] == "totals":      # LayerNorm2d consumer: [N,chunks,4,2] totals over all channels
            d.stats_groups, d.stats_gmax = 4, 0
            assert partial.shape[2] == 4 and c_off == 0
        assert impl == "tc" and partial.dtype == torch.float32 and partial.dim() == 4 and partial.shape[0] == d.y.n
        d.stats_partial = partial.data_ptr()
        d.stats_c_off, d.stats_c_total = c_off, partial.shape[2]
        d.stats_chunk_off, d.stats_chunks_total = chunk_off, partial.shape[1]
    if narrow is not None:      # (cin_from, cout): input channels >= cin_from only feed the first `cout` outputs (zero weight block)
        d.narrow_cin_from, d.narrow_cout = narrow
        assert impl == "tc" and narrow[0] % 64 == 0 and narrow[1] % 32 == 0
    keep = (d, x, w, y, scale, bias, res1, res2, y_f32, x2, stats)
    cout, taps = d.y.c, k[0] * k[1]
    for t in (scale, bias):
        assert t is None or (t.dtype == torch.float32 and t.numel() == cout), (name, cout)
    for r in (res1, res2):
        assert r is None or tuple(r.shape) == (d.y.n, d.y.h, d.y.w, cout), (name, r.shape)
    if impl == "simt":
        co_pad = -(-cout // 16) * 16 if cout >= 16 else -(-cout // 4) * 4
        assert w.dtype == torch.float32 and tuple(w.shape) == (taps, x.shape[3], co_pad), (name, w.shape, x.shape, cout)
    else:
        kcols = taps * (-(-x.shape[3] // 64) * 64) + (k2[0] * k2[1] * (-(-x2.shape[3] // 64) * 64) if x2 is not None else 0)
        assert w.dtype == torch.float16 and tuple(w.shape) == (-(-cout // 8) * 8, kcols), (name, w.shape, x.shape, cout)
        assert x2 is None or impl == "tc"
    if impl == "simt":
        op = Op(name + "[simt]", lib.s2v_conv_simt, (C.byref(d),), keep)
    else:
        if box is None:
            box = conv_box(d.y.h, d.y.w, k, stride, dil)
        op = Op(name + "[tc]", lib.s2v_conv_tc, (C.byref(d), box[0], box[1], box[2]), keep)
    # algorithmic FLOPs of the reference op this launch replaces (2*MACs on true channel counts);
    # alg_scale lets the sub-pixel phases of nearest-x2 + conv3x3 report the reference's 3x3 work
    op.alg_flops = 2.0 * d.y.n * d.y.h * d.y.w * cout * taps * cin_true * alg_scale
    if x2 is not None:
        op.alg_flops += 2.0 * d.y.n * d.y.h * d.y.w * cout * k2[0] * k2[1] * x2.shape[3]
    if alg_flops is not None:           # launches that execute structurally-zero weight blocks report the reference's work
        op.alg_flops = float(alg_flops)
    # minimum HBM bytes of this launch (every operand read once, the output written once): with alg_flops it gives the
    # per-layer roofline time max(flops / tensor peak, bytes / HBM peak) that bench.py sums over the step
    out_b = (4.0 if y is None else 2.0) * d.y.n * d.y.h * d.y.w * cout
    op.io_bytes = (2.0 * x.numel() + w.numel() * w.element_size() + out_b
                   + sum(2.0 * t.numel() for t in (res1, res2, x2) if t is not None))
    return op


def box_tiles(h: int, w: int, n: int, k=(1, 1), stride=(1, 1), dil=(1, 1)) -> int:
    """Spatial tiles per image of a conv_tc launch with the default box (= statistics chunks it emits)."""
    bw, bh, _ = conv_box(h, w, k, stride, dil)
    return -(-w // bw) * -(-h // bh)


def stats_fusable(lib, cout: int) -> bool:
    """Whether the conv_tc epilogue can emit the output statistics of a Cout-wide conv (its <=128-column passes must
    be 32, 64 or 128 columns wide, see s2v_conv_tc)."""
    return lib.s2v_conv_tc_tile_n(cout) in (32, 64, 128, 192, 256)


def stats_chunks(n: int, hw: int, c: int = 64) -> int:
    """Pixel chunks per image for chan_stats.  A function of the image geometry ONLY: the reduction
    tree must not depend on the batch size, so that a frame's result is bit-identical whichever
    batch / GPU shard it is computed in.  Sized so that every thread streams ~8 pixels (two batches of
    four independent 16-byte loads): a block covers min(32, 256/(C/8)) pixel lanes."""
    pg = min(32, max(1, 256 // max(c // 8, 1)))
    return max(1, min(hw // (8 * pg), 64))


def op_chan_stats(lib, x, chunks, partial) -> Op:
    v = view(x)
    op = Op("chan_stats", lib.s2v_chan_stats, (C.byref(v), chunks, _ptr(partial)), (v, x, partial))
    op.alg_bytes = 2.0 * x.numel()
    return op


def op_ln2d_finalize(lib, partial, n, chunks, c, count, gamma, beta, a, b, eps=1e-5) -> Op:
    return Op("ln2d_finalize", lib.s2v_ln2d_finalize,
              (_ptr(partial), n, chunks, c, count, _ptr(gamma), _ptr(beta), eps, _ptr(a), _ptr(b)),
              (partial, gamma, beta, a, b))


def op_ln2d_finalize_totals(lib, partial, n, chunks, c, count, gamma, beta, a, b, eps=1e-5) -> Op:
    """partial [n, chunks, 4, 2]: LayerNorm2d totals written by s2v_conv_tc (stats_gmax = 0)."""
    assert tuple(partial.shape) == (n, chunks, 4, 2)
    return Op("ln2d_finalize", lib.s2v_ln2d_finalize_totals,
              (_ptr(partial), n, chunks, c, count, _ptr(gamma), _ptr(beta), eps, _ptr(a), _ptr(b)),
              (partial, gamma, beta, a, b))


def op_adain_finalize(lib, partial, n, chunks, c, count, gamma, beta, gb_stride, a, b, eps=1e-5) -> Op:
    return Op("adain_finalize", lib.s2v_adain_finalize,
              (_ptr(partial), n, chunks, c, count, _ptr(gamma), _ptr(beta), gb_stride, eps, _ptr(a), _ptr(b)),
              (partial, gamma, beta, a, b))


def op_affine_act(lib, x, a, b, y, *, act=L.ACT_NONE, act_param=0.0, pool2=0, res=None, reflect1=0, res_ab=None) -> Op:
    vx, vy = view(x), view(y)
    vr = view(res) if res is not None else null_view()
    if res_ab is not None:          # y = act(x*a+b) + act(res*ra+rb): res is a raw conv output with its own affine
        assert res is not None and not pool2
        ra, rb = res_ab
        op = Op("affine_act", lib.s2v_affine_act2,
                (C.byref(vx), _ptr(a), _ptr(b), act, float(act_param), C.byref(vr), _ptr(ra), _ptr(rb), C.byref(vy), reflect1),
                (vx, vy, vr, x, y, res, a, b, ra, rb))
        pad = (y.shape[1] + 2) * (y.shape[2] + 2) / float(y.shape[1] * y.shape[2]) if reflect1 else 1.0
        op.alg_bytes = 2.0 * (x.numel() + res.numel() + y.numel() * pad)
        return op
    op = Op("affine_act", lib.s2v_affine_act,
            (C.byref(vx), _ptr(a), _ptr(b), act, float(act_param), pool2, C.byref(vr), C.byref(vy), reflect1),
            (vx, vy, vr, x, y, res, a, b))
    pad = (y.shape[1] + 2) * (y.shape[2] + 2) / float(y.shape[1] * y.shape[2]) if reflect1 else 1.0
    op.alg_bytes = 2.0 * (x.numel() + (res.numel() if res is not None else 0) + y.numel() * pad)
    return op


def op_adain_fused(lib, x, gamma, beta, gb_stride, y, *, act=L.ACT_NONE, act_param=0.0, res=None, reflect1=0, eps=1e-5) -> Op:
    vx, vy = view(x), view(y)
    vr = view(res) if res is not None else null_view()
    return Op("adain_fused", lib.s2v_adain_fused,
              (C.byref(vx), _ptr(gamma), _ptr(beta), gb_stride, eps, act, float(act_param), C.byref(vr), C.byref(vy), reflect1),
              (vx, vy, vr, x, y, res, gamma, beta))


def op_token_ln(lib, x, gamma, beta, y, eps=1e-5) -> Op:
    vx, vy = view(x), view(y)
    op = Op("token_layernorm", lib.s2v_token_layernorm, (C.byref(vx), _ptr(gamma), _ptr(beta), eps, C.byref(vy)),
            (vx, vy, x, y, gamma, beta))
    op.alg_bytes = 2.0 * (x.numel() + y.numel())
    return op


def op_add(lib, a, b, y) -> Op:
    va, vb, vy = view(a), view(b), view(y)
    return Op("add", lib.s2v_add, (C.byref(va), C.byref(vb), C.byref(vy)), (va, vb, vy, a, b, y))


def op_reflect_border(lib, interior) -> Op:
    v = view(interior)
    return Op("reflect_border", lib.s2v_reflect_border, (C.byref(v),), (v, interior))


def op_rfft2(lib, x, spec) -> Op:
    vx, vs = view(x), view(spec)
    op = Op("rfft2", lib.s2v_rfft2, (C.byref(vx), C.byref(vs)), (vx, vs, x, spec))
    op.alg_bytes = 2.0 * (x.numel() + spec.numel())
    return op


def op_irfft2(lib, spec, add, y) -> Op:
    vs, vy = view(spec), view(y)
    va = view(add) if add is not None else null_view()
    op = Op("irfft2", lib.s2v_irfft2, (C.byref(vs), C.byref(va), C.byref(vy)), (vs, va, vy, spec, add, y))
    op.alg_bytes = 2.0 * (spec.numel() + (add.numel() if add is not None else 0) + y.numel())
    return op


def op_attention(lib, q, k, v, o, heads, scale) -> Op:
    vq, vk, vv, vo = view(q), view(k), view(v), view(o)
    return Op("attention", lib.s2v_attention, (C.byref(vq), C.byref(vk), C.byref(vv), C.byref(vo), heads, float(scale)),
              (vq, vk, vv, vo, q, k, v, o))


def op_pack(lib, src, dst, c_off=0, c_fill=None, scale=1.0, shift=0.0) -> Op:
    """src: float32 [N,C,H,W], contiguous or a channel window (narrow on dim 1) of a contiguous tensor."""
    n, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.stride(3) == 1 and src.stride(2) == w and src.stride(1) == h * w
    vd = view(dst)
    c_fill = c if c_fill is None else c_fill
    return Op("pack_nchw", lib.s2v_pack_nchw_f32,
              (_ptr(src), n, c, h, w, src.stride(0) if n > 1 else c * h * w, C.byref(vd), c_off, c_fill, float(scale), float(shift)),
              (vd, src, dst))


def op_unpack(lib, src, c_off, c, dst) -> Op:
    vs = view(src)
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    return Op("unpack_nchw", lib.s2v_unpack_to_nchw_f32, (C.byref(vs), c_off, c, _ptr(dst)), (vs, src, dst))


def op_grouped_linear(lib, hidden, groups_dev, tile2group_dev, n_tiles, out) -> Op:
    b = hidden.shape[0]
    h2 = hidden.reshape(b, -1)
    assert h2.stride(1) == 1 and out.dim() == 2 and out.stride(1) == 1
    return Op("grouped_linear", lib.s2v_grouped_linear,
              (_ptr(h2), h2.stride(0), b, _ptr(groups_dev), _ptr(tile2group_dev), n_tiles, _ptr(out), out.stride(0)),
              (hidden, h2, groups_dev, tile2group_dev, out))


def op_resize(lib, x, y, chan_scale=None) -> Op:
    """bilinear resize (align_corners=False) of an fp16 NHWC tensor, optional per-(n, c) float32 scale [N, C]."""
    vx, vy = view(x), view(y)
    assert chan_scale is None or (chan_scale.dtype == torch.float32 and tuple(chan_scale.shape) == (x.shape[0], x.shape[3]) and chan_scale.stride(1) == 1)
    op = Op("resize_bilinear", lib.s2v_resize_bilinear,
            (C.byref(vx), C.byref(vy), _ptr(chan_scale), 0 if chan_scale is None else chan_scale.stride(0)), (vx, vy, x, y, chan_scale))
    op.alg_bytes = 2.0 * (x.numel() + y.numel())
    return op


def op_resize_planes(lib, src, dst) -> Op:
    """src [N,P,H,W] float32 (a channel window of a contiguous NCHW tensor) -> dst [N,P,OH,OW], bilinear, align_corners=False."""
    n, p, h, w = src.shape
    assert src.dtype == dst.dtype == torch.float32 and src.stride(3) == 1 and src.stride(2) == w and dst.stride(3) == 1 and dst.stride(2) == dst.shape[3]
    assert tuple(dst.shape[:2]) == (n, p)
    return Op("resize_planes", lib.s2v_resize_planes_f32,
              (_ptr(src), src.stride(0), src.stride(1), n, p, h, w, _ptr(dst), dst.stride(0), dst.stride(1), dst.shape[2], dst.shape[3]), (src, dst))


def op_style_demod(lib, w2, s, cin, cout, gain, out, eps=1e-8) -> Op:
    """s: float32 rows (a column window of the modulation table) [N, >= cin]; out float32 [N, cout]."""
    assert w2.dtype == torch.float32 and tuple(w2.shape) == (cout, cin) and s.stride(1) == 1 and out.is_contiguous()
    return Op("style_demod", lib.s2v_style_demod, (_ptr(w2), _ptr(s), s.stride(0), s.shape[0], cin, cout, float(eps), float(gain), _ptr(out)),
              (w2, s, out))


def op_style_epilogue(lib, x, y, *, a=None, bias=None, noise=None, noise_w=None, slope=0.2, post=None) -> Op:
    vx, vy = view(x), view(y)
    assert post is None or post.stride(1) == 1
    op = Op("style_epilogue", lib.s2v_style_epilogue,
            (C.byref(vx), _ptr(a), _ptr(bias), _ptr(noise), _ptr(noise_w), float(slope), _ptr(post), 0 if post is None else post.stride(0), C.byref(vy)),
            (vx, vy, x, y, a, bias, noise, noise_w, post))
    op.alg_bytes = 2.0 * (x.numel() + y.numel()) + (4.0 * noise.numel() if noise is not None else 0.0)
    return op


def op_to_rgb(lib, x, w, s, bias, skip, out, crop=0) -> Op:
    vx = view(x)
    assert w.dtype == torch.float32 and tuple(w.shape) == (3, x.shape[3]) and s.stride(1) == 1 and out.is_contiguous()
    op = Op("to_rgb", lib.s2v_to_rgb, (C.byref(vx), _ptr(w), _ptr(s), s.stride(0), _ptr(bias), _ptr(skip), _ptr(out), int(crop)),
            (vx, x, w, s, bias, skip, out))
    op.alg_bytes = 2.0 * x.numel() + 4.0 * out.numel() + (4.0 * skip.numel() if skip is not None else 0.0)
    return op


def op_reflect_pad_nchw(lib, src, pad, dst) -> Op:
    n, c, h, w = src.shape
    assert src.is_contiguous() and dst.is_contiguous() and tuple(dst.shape) == (n, c, h + 2 * pad, w + 2 * pad)
    return Op("reflect_pad_nchw", lib.s2v_reflect_pad_nchw_f32, (_ptr(src), n * c, h, w, pad, _ptr(dst)), (src, dst))


def op_mean_over_w(lib, x, y) -> Op:
    vx, vy = view(x), view(y)
    return Op("mean_over_w", lib.s2v_mean_over_w, (C.byref(vx), C.byref(vy)), (vx, vy, x, y))


WARP_PACK_SRC = 0x100     # include/s2v.h: S2V_WARP_PACK_SRC


def op_flow_warp(lib, src, flow, out, out16=None, c_off=0, pack_src=False) -> Op:
    """pack_src: the launch writes the whole 8-channel fp16 texel [src | warp | 0] of out16 (c_off == C), not only the warp channels."""
    b, c, h, w = src.shape
    fh, fw = flow.shape[2], flow.shape[3]
    v16 = view(out16) if out16 is not None else null_view()
    if pack_src:
        assert out16 is not None and c_off == c and src.is_contiguous()
    return Op("flow_warp", lib.s2v_flow_warp_f32,
              (_ptr(src), _ptr(flow), _ptr(out), b, c, h, w, fh, fw, C.byref(v16), c_off | (WARP_PACK_SRC if pack_src else 0)),
              (src, flow, out, out16, v16))


# ------------------------------------------------------------------ weight packing


def pack_w_tc(w: torch.Tensor) -> torch.Tensor:
    """[Cout,Cin,kh,kw] float -> fp16 [Cout][Cin64/64][kh*kw][64]: K-major, CHUNK-major (all taps of a
    64-channel chunk are adjacent, so one A patch serves them), zero-filled channel pad."""
    co, ci, kh, kw = w.shape
    ci64 = -(-ci // 64) * 64
    co8 = -(-co // 8) * 8                       # rows padded to 8 (Cout = 3 heads)
    out = torch.zeros(co8, kh * kw, ci64, dtype=torch.float16, device=w.device)
    out[:co, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(torch.float16)
    out = out.reshape(co8, kh * kw, ci64 // 64, 64).permute(0, 2, 1, 3)
    return out.reshape(co8, kh * kw * ci64).contiguous()


def head_co_pad(cout: int) -> int:
    """Channel pad of the folded-tap head GEMM (csrc/conv_head.cu): the smallest of 2 / 4 / 8 that holds Cout."""
    assert 0 < cout <= 8
    return 2 if cout <= 2 else 4 if cout <= 4 else 8


def pack_w_head(w: torch.Tensor) -> torch.Tensor:
    """7x7 head conv [Cout<=8, Cin, 7, 7] -> fp16 [NB][chunks*7*64] for s2v_conv_head: row kx*CP + co, column
    (chunk*7 + ky)*64 + ci (the kx taps are folded into the GEMM's N dimension; CP = head_co_pad(Cout), NB = 16 / 32 / 64 rows)."""
    co, ci, kh, kw = w.shape
    assert kh == 7 and kw == 7 and co <= 8 and ci % 64 == 0
    chunks, cp = ci // 64, head_co_pad(co)
    nb = {2: 16, 4: 32, 8: 64}[cp]
    out = torch.zeros(nb // cp, cp, chunks, 7, 64, dtype=torch.float16, device=w.device)      # [kx][co][chunk][ky][ci]
    out[:7, :co] = w.reshape(co, chunks, 64, 7, 7).permute(4, 0, 1, 3, 2).to(torch.float16)   # (kx, co, chunk, ky, ci)
    return out.reshape(nb, chunks * 7 * 64).contiguous()


def op_conv_head(lib, x, w, y_f32, *, bias=None, act=L.ACT_NONE, act_param=0.0, name="head") -> Op:
    """x fp16 NHWC [N,H,W,Cin]; y_f32 float32 [N,Cout,H,W]; w from pack_w_head."""
    d = L.Conv()
    d.x = view(x)
    n, co, oh, ow = y_f32.shape
    d.y = L.View(None, n, oh, ow, co, 0, 0, 0)
    d.out_mode = L.OUT_F32_NCHW
    d.y_f32 = y_f32.data_ptr()
    d.w = w.data_ptr()
    d.bias = None if bias is None else bias.data_ptr()
    d.res1, d.res2, d.x2 = null_view(), null_view(), null_view()
    d.kh = d.kw = 7
    d.stride_h = d.stride_w = d.dil_h = d.dil_w = 1
    d.pad_h = d.pad_w = 3
    d.act, d.act_param = act, float(act_param)
    nb = {2: 16, 4: 32, 8: 64}[head_co_pad(co)]
    assert w.dtype == torch.float16 and tuple(w.shape) == (nb, x.shape[3] // 64 * 7 * 64) and x.shape[3] % 64 == 0
    op = Op(name + "[head]", lib.s2v_conv_head, (C.byref(d),), (d, x, w, y_f32, bias))
    op.alg_flops = 2.0 * n * oh * ow * co * 49 * x.shape[3]
    op.io_bytes = 2.0 * x.numel() + 2.0 * w.numel() + 4.0 * y_f32.numel()
    return op


def pack_w_tc_rowtaps(w: torch.Tensor, cpad: int = 8) -> torch.Tensor:
    """Stem convs with tiny Cin (3 or 6): [Cout,Cin,kh,kw] -> fp16 [Cout][kh][64] where the 64-wide K
    chunk of row tap ky holds (kx, c) = kx*cpad + c, i.e. kw consecutive pixels of a channels-last
    buffer with cpad channels (read by an OVERLAPPING view: pixel stride cpad, 64 'channels')."""
    co, ci, kh, kw = w.shape
    assert kw * cpad <= 64 and ci <= cpad
    out = torch.zeros(co, kh, 64 // cpad, cpad, dtype=torch.float16, device=w.device)
    out[:, :, :kw, :ci] = w.permute(0, 2, 3, 1).to(torch.float16)
    return out.reshape(co, kh * 64).contiguous()


def pack_w_simt(w: torch.Tensor, cin_pad: int | None = None) -> torch.Tensor:
    """[Cout,Cin,kh,kw] float -> float32 [kh*kw][Cin_pad][Cout_pad]."""
    co, ci, kh, kw = w.shape
    cin_pad = cin_pad or -(-ci // 8) * 8
    co_pad = -(-co // 16) * 16 if co >= 16 else -(-co // 4) * 4
    out = torch.zeros(kh * kw, cin_pad, co_pad, dtype=torch.float32, device=w.device)
    out[:, :ci, :co] = w.permute(2, 3, 1, 0).reshape(kh * kw, ci, co).float()
    return out.contiguous()
