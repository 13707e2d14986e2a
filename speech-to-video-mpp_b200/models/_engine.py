"""Shared machinery of the LNet / DNet CUDA engines: weight folding, workspace arena, norm
helpers, CUDA-graph capture.  Everything here only *orders* C-ABI calls; all arithmetic on the
path happens in libs2v's kernels."""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading

import torch

from .. import _lib as L
from .. import ops


def sn_fold(sd, p):
    """Eval-mode spectral norm: W / (u^T W_mat v) with the STORED u, v (reference quirk C.4,
    models/base_blocks.py:72-76); plain ``weight`` when the conv has no spectral norm."""
    if p + ".weight" in sd:
        return sd[p + ".weight"].float()
    w = sd[p + ".weight_orig"].float()
    sigma = torch.dot(sd[p + ".weight_u"].float(), torch.mv(w.flatten(1), sd[p + ".weight_v"].float()))
    return w / sigma


def bn_fold(sd, p, conv_bias=None, eps=1e-5):
    """Eval BatchNorm after a conv -> per-channel (scale, shift) for the conv epilogue."""
    s = sd[p + ".weight"].float() / torch.sqrt(sd[p + ".running_var"].float() + eps)
    b = sd[p + ".bias"].float() - sd[p + ".running_mean"].float() * s
    if conv_bias is not None:
        b = b + conv_bias.float() * s
    return s.contiguous(), b.contiguous()


def up2_phase_weights(w):
    """nearest-x2 upsample followed by a 3x3 zero-padded conv == four 2x2 convs on the low-res
    input, one per output parity (p,q), with top/left padding (1-p, 1-q).  Returns {(p,q): w2x2}."""
    grp = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = {}
    for p in (0, 1):
        for q in (0, 1):
            w4 = torch.zeros(w.shape[0], w.shape[1], 2, 2, dtype=w.dtype, device=w.device)
            for a in (0, 1):
                for b in (0, 1):
                    for ky in grp[p][a]:
                        for kx in grp[q][b]:
                            w4[:, :, a, b] += w[:, :, ky, kx]
            out[(p, q)] = w4
    return out


class EngineBase:
    """Holds packed weights + per-batch-size plans (workspace + op list + optional CUDA graph)."""

    def __init__(self, device: torch.device, conv_impl: str = "tc", use_graph: bool = True):
        self.dev = device
        # a CPU "device" only builds plans (shape/ABI validation in the -m "not gpu" tests); nothing can run there
        self.lib = L.require_device(device.index) if device.type == "cuda" else L.load_library()
        self.impl = conv_impl
        self.use_graph = use_graph
        # Weight folding / packing (spectral norm, BatchNorm, phase weights, K-major fp16 packing) runs once per engine on the
        # HOST by default - a few hundred MB of one-time numpy-style work instead of ~1000 tiny ATen launches in front of the
        # first forward; the packed tensors are uploaded by finish_pack().  S2V_FOLD_DEVICE=1 folds on the GPU instead.
        self.fold_dev = device if os.environ.get("S2V_FOLD_DEVICE", "0") == "1" else torch.device("cpu")
        self.W = {}
        # per-batch-size plans, least recently used first.  Every plan owns a private workspace (dozens of activation
        # buffers) and a captured CUDA graph, so the cache is bounded: at most S2V_MAX_PLANS plans and S2V_PLAN_CACHE_GB of
        # workspace; the least recently used plan is dropped first (its buffers go back to torch's caching allocator).
        self._plans = collections.OrderedDict()
        self.max_plans = int(os.environ.get("S2V_MAX_PLANS", "8"))
        self.max_plan_bytes = int(float(os.environ.get("S2V_PLAN_CACHE_GB", "96")) * (1 << 30))
        # One forward at a time per engine: a plan's I/O and workspace buffers are shared by every call that uses it.
        # The lock serialises host-side issue; _last_done orders a forward issued on a DIFFERENT stream behind the previous one.
        self._lock = threading.RLock()
        self._last_done = None
        self._last_stream = None

    # ---- workspace ---------------------------------------------------------------------
    def buf(self, ws, name, shape, dtype=torch.float16, zero=False):
        if name in ws:
            t = ws[name]
            assert tuple(t.shape) == tuple(shape) and t.dtype == dtype, name
            return t
        t = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=self.dev)
        if zero:
            t._s2v_init = ("zero", 0.0)          # plan_export: a region the plan expects initialised (padding borders, never-written rows)
        ws[name] = t
        return t

    def finish_pack(self):
        """Uploads everything _pack produced (packed weights, epilogue vectors, norm parameters) to the engine's device."""
        mv = lambda t: t.to(self.dev) if isinstance(t, torch.Tensor) else t
        for ent in self.W.values():
            for k in list(ent):
                if k != "w_f32":                    # kept on the folding device for the lazy head packing
                    ent[k] = mv(ent[k])
        if hasattr(self, "P"):
            self.P = {k: mv(v) for k, v in self.P.items()}

    # ---- conv helpers ------------------------------------------------------------------
    def pack_conv(self, name, w, bias=None, scale=None, impl=None, cin_pad=None, rowtaps=False):
        """w: folded fp32 [Cout,Cin,kh,kw].  rowtaps (tc only): tiny-Cin stem packed for the
        overlapping-view trick (ops.pack_w_tc_rowtaps); the conv then runs as a kh x 1 conv."""
        impl = impl or self.impl
        ent = dict(impl=impl, k=(w.shape[2], w.shape[3]), cout=w.shape[0],
                   bias=None if bias is None else bias.float().contiguous(),
                   scale=None if scale is None else scale.float().contiguous())
        if impl == "tc" and tuple(w.shape[2:]) == (7, 7) and w.shape[0] <= 8 and w.shape[1] % 64 == 0:
            ent["w_f32"] = w.float()              # kept for the folded-tap head packing (ops.pack_w_head)
        if impl == "tc" and rowtaps:
            ent["w"], ent["k"] = ops.pack_w_tc_rowtaps(w, cin_pad or 8), (w.shape[2], 1)
        else:
            ent["w"] = ops.pack_w_tc(w) if impl == "tc" else ops.pack_w_simt(w, cin_pad)
        self.W[name] = ent
        return ent

    def conv(self, plan, name, x, y, **kw):
        e = self.W[name]
        kw.setdefault("k", e["k"])
        kw.setdefault("bias", e["bias"])
        kw.setdefault("scale", e["scale"])
        return plan.add(ops.op_conv(self.lib, x, e["w"], y, impl=e["impl"], name=name, **kw))

    def head_conv(self, plan, name, x, y_f32, *, act=L.ACT_NONE, act_param=0.0):
        """7x7 pad-3 head conv with Cout <= 8 to a float32 NCHW tensor (FinalBlock2d, flow_out): s2v_conv_head (kx taps folded
        into N) on the tc path when the geometry allows, else the generic conv."""
        e = self.W[name]
        n, co, oh, ow = y_f32.shape
        if e["impl"] == "tc" and e["k"] == (7, 7) and co <= 8 and x.shape[3] % 64 == 0 and os.environ.get("S2V_HEAD", "1") == "1":
            if "w_head" not in e:
                e["w_head"] = ops.pack_w_head(e["w_f32"]).to(self.dev)
            return plan.add(ops.op_conv_head(self.lib, x, e["w_head"], y_f32, bias=e["bias"], act=act, act_param=act_param, name=name))
        return self.conv(plan, name, x, None, pad=(3, 3), act=act, act_param=act_param, y_f32=y_f32, out_shape=(n, co, oh, ow))

    def stem_conv(self, plan, ws, name, src_nchw, y, *, k=7, cin_true=3, stats=False, fin=None):
        """k x k zero-padded stem conv on a tiny-Cin NCHW float input (FirstBlock2d / input_layer,
        models/base_blocks.py:79-92,312).  tc path: the input is packed to 8 channels into a buffer
        padded by k//2; one row tap's k x 8 = 56 (+8 zero-weight) consecutive values are exactly one
        64-wide K chunk, exposed to TMA as an OVERLAPPING view (pixel stride 8, 64 'channels')."""
        n, c, h, w = src_nchw.shape
        pad = k // 2
        if self.W[name]["impl"] != "tc":
            x8 = self.buf(ws, name + ".in8", (n, h, w, 8))
            plan.add(ops.op_pack(self.lib, src_nchw, x8, 0, 8))
            self.conv(plan, name, x8, y, pad=(pad, pad), cin_true=cin_true)
            return None
        wp = w + 2 * pad + 2                        # +2: the 8th (zero-weight) pixel of the last window stays in the row
        xp = self.buf(ws, name + ".in8p", (n, h + 2 * pad, wp, 8), zero=True)
        plan.add(ops.op_pack(self.lib, src_nchw, xp[:, pad:pad + h, pad:pad + w, :], 0, 8))
        win = torch.as_strided(xp, (n, h + 2 * pad, w, 64), (xp.stride(0), xp.stride(1), 8, 1))
        if stats:
            return self.conv_stats(plan, ws, name, win, y, pad=(0, 0), cin_true=cin_true * k, fin=fin)
        self.conv(plan, name, win, y, pad=(0, 0), cin_true=cin_true * k)
        return None

    def conv_stats(self, plan, ws, name, x, y, *, tag=None, c_total=None, c_off=0, phase=0, phases=1, fuse=True, fin=None, **kw):
        """conv whose epilogue also emits the per-(image, spatial tile, channel) sum / sum of squares of its fp16 output
        (tc path) - this replaces the separate full-tensor chan_stats pass.  ``fin`` names the consumer of the statistics:
        ("ln", gamma, beta) = a LayerNorm2d over (C,H,W), for which the conv only emits per-tile TOTALS from its accumulator
        registers; ("adain", ...) or None = per-channel partials.
        Returns a dict for layernorm2d/adain(stats=...), or None when the statistics cannot be fused (simt path, odd tile
        widths, S2V_FUSED_STATS=0): the caller then falls back to a chan_stats pass."""
        n, h, w, c = y.shape
        ct = c_total or c
        if self.W[name]["impl"] != "tc" or os.environ.get("S2V_FUSED_STATS", "1") != "1" or not fuse or not ops.stats_fusable(self.lib, c):
            self.conv(plan, name, x, y, **kw)
            return None
        k = kw.get("k", self.W[name]["k"])
        tiles = ops.box_tiles(h, w, n, k, kw.get("stride", (1, 1)), kw.get("dil", (1, 1)))
        t = tag or name
        # LayerNorm2d consumer (fin = ("ln", ...)): the conv only emits totals over all channels, from its accumulator
        # registers - no shared-memory statistics pass in the epilogue and a 4-entry-per-tile partial list to finalize
        totals = (fin is not None and fin[0] == "ln" and c_off == 0 and ct == c and c <= 256 and self.lib.s2v_conv_tc_tile_n(c) >= c
                  and os.environ.get("S2V_LN_TOTALS", "1") == "1")
        if totals:
            partial = self.buf(ws, t + ".epi_totals", (n, tiles * phases, 4, 2), torch.float32, zero=True)
            self.conv(plan, name, x, y, stats=(partial, 0, phase * tiles, "totals"), **kw)
            return dict(partial=partial, chunks=tiles * phases, done=False, totals=True)
        partial = self.buf(ws, t + ".epi_partial", (n, tiles * phases, ct, 2), torch.float32, zero=True)
        self.conv(plan, name, x, y, stats=(partial, c_off, phase * tiles), **kw)
        return dict(partial=partial, chunks=tiles * phases, done=False)

    # ---- norm helpers ------------------------------------------------------------------
    def _stats(self, plan, ws, tag, x):
        n, h, w, c = x.shape
        chunks = ops.stats_chunks(n, h * w, c)
        partial = self.buf(ws, tag + ".partial", (n, chunks, c, 2), torch.float32)
        plan.add(ops.op_chan_stats(self.lib, x, chunks, partial))
        return dict(partial=partial, chunks=chunks, done=False)

    def _ab(self, ws, tag, st, n, c):
        if st.get("done"):
            return st["a"], st["b"]
        return self.buf(ws, tag + ".a", (n, c), torch.float32), self.buf(ws, tag + ".b", (n, c), torch.float32)

    def ln2d_scale_shift(self, plan, ws, tag, x, gamma, beta, *, stats=None):
        """The per-(n,c) scale / shift of LayerNorm2d(x) without applying it (for layernorm2d(res_ab=...))."""
        n, h, w, c = x.shape
        st = stats if stats is not None else self._stats(plan, ws, tag, x)
        a, b = self._ab(ws, tag, st, n, c)
        if st.get("totals"):
            plan.add(ops.op_ln2d_finalize_totals(self.lib, st["partial"], n, st["chunks"], c, h * w, gamma, beta, a, b))
        elif not st.get("done"):
            plan.add(ops.op_ln2d_finalize(self.lib, st["partial"], n, st["chunks"], c, h * w, gamma, beta, a, b))
        return a, b

    def layernorm2d(self, plan, ws, tag, x, gamma, beta, y, *, slope=0.1, pool2=0, res=None, reflect1=0, stats=None, res_ab=None):
        """LayerNorm2d over (C,H,W) + LeakyReLU(slope) [+ AvgPool2] [+ res] (base_blocks.py:52-69,79-124).
        ``stats``: what conv_stats returned for the producer of x (partials, possibly already finalized).
        ``res_ab``: res is a RAW tensor whose own LayerNorm2d scale / shift (ln2d_scale_shift) and the same LeakyReLU
        are applied on the fly: y = lrelu(LN(x)) + lrelu(LN(res)) in one pass."""
        a, b = self.ln2d_scale_shift(plan, ws, tag, x, gamma, beta, stats=stats)
        plan.add(ops.op_affine_act(self.lib, x, a, b, y, act=L.ACT_LRELU, act_param=slope, pool2=pool2, res=res,
                                   reflect1=reflect1, res_ab=res_ab))

    def adain(self, plan, ws, tag, x, gamma, beta, gb_stride, y, *, act=L.ACT_LRELU, slope=0.01, res=None, reflect1=0,
              stats=None):
        """InstanceNorm2d*(1+gamma)+beta + activation [+ res] (base_blocks.py:127-157).
        ``stats``: what conv_stats returned for the producer of x, or the return value of an earlier adain() on the same x
        (two AdaINs of one tensor share the partials, not the finalized scale / shift)."""
        n, h, w, c = x.shape
        if stats is None and act in (L.ACT_NONE, L.ACT_LRELU, L.ACT_RELU) and self.lib.s2v_adain_fused_fits(h, w, c) > 0 \
                and os.environ.get("S2V_ADAIN_FUSED", "0") == "1":
            plan.add(ops.op_adain_fused(self.lib, x, gamma, beta, gb_stride, y, act=act, act_param=slope, res=res,
                                        reflect1=reflect1))
            return None
        st = stats if stats is not None else self._stats(plan, ws, tag, x)
        a, b = self._ab(ws, tag, st, n, c)
        if not st.get("done"):
            plan.add(ops.op_adain_finalize(self.lib, st["partial"], n, st["chunks"], c, h * w, gamma, beta, gb_stride, a, b))
        plan.add(ops.op_affine_act(self.lib, x, a, b, y, act=act, act_param=slope, res=res, reflect1=reflect1))
        return dict(partial=st["partial"], chunks=st["chunks"], done=False)

    # ---- grouped AdaIN heads -----------------------------------------------------------
    def pack_lin_groups(self, name, groups):
        """groups: list of (wt [K,nout] fp32, bias [nout], in_off, out_off)."""
        arr = (L.LinGroup * len(groups))()
        tiles, keep = [], []
        for i, (wt, bias, in_off, out_off) in enumerate(groups):
            wt, bias = wt.float().contiguous().to(self.dev), bias.float().contiguous().to(self.dev)
            if wt.shape[0] > 512:               # kLinMaxK of csrc/linear.cu: the hidden vector is staged in a fixed smem buffer
                raise ValueError("s2v_grouped_linear supports hidden widths up to 512, got %d" % wt.shape[0])
            keep += [wt, bias]
            arr[i] = L.LinGroup(wt.data_ptr(), bias.data_ptr(), in_off, wt.shape[0], out_off, wt.shape[1])
            tiles += [(i, j) for j in range(0, wt.shape[1], 128)]
        gdev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
        tdev = torch.tensor(tiles, dtype=torch.int32, device=self.dev)
        self.W[name] = dict(groups=gdev, tiles=tdev, n_tiles=len(tiles), keep=keep)

    # ---- plan execution ----------------------------------------------------------------
    # Frames are independent, so a batch can run as `self.streams` independent sub-batches on parallel streams (own
    # workspace each, shared weights): while one sub-batch sits in a kernel's prologue / tail / launch gap the other's
    # CTAs fill the idle SMs.  The I/O buffers stay whole-batch tensors; each part works on its slice.
    streams = int(os.environ.get("S2V_STREAMS", "1"))

    def _get_plan(self, key, builder, builder_of=None, batch=None, io_spec=None):
        """builder(plan, ws) -> io dict.  With builder_of(B) / io_spec(B) -> {io name: (workspace name, shape, dtype)} given,
        a batch of `batch` frames is split over `self.streams` parallel sub-plans."""
        if key in self._plans:
            self._plans.move_to_end(key)
        else:
            self._evict(reserve=1)
            B = batch or 0
            parts_n = self.streams if (builder_of and io_spec and self.streams > 1 and B % self.streams == 0 and B // self.streams >= 8) else 1
            if parts_n == 1:
                ws, plan = {}, ops.Plan()
                io = builder(plan, ws)
                self._plans[key] = dict(plan=plan, ws=ws, io=io, graph=None, warm=0, parts=None)
            else:
                sub = B // parts_n
                spec = io_spec(B)
                io = {name: torch.empty(shape, dtype=dt, device=self.dev) for name, (_, shape, dt) in spec.items()}
                parts, allp = [], ops.Plan()
                for i in range(parts_n):
                    ws, plan = {}, ops.Plan()
                    for name, (ws_name, _, _) in spec.items():      # part i works on rows [i*sub, (i+1)*sub) of the whole-batch I/O
                        ws[ws_name] = io[name][i * sub:(i + 1) * sub]
                    builder_of(sub)(plan, ws)
                    parts.append(dict(plan=plan, ws=ws, stream=torch.cuda.Stream(device=self.dev)))
                    allp.main_ops.extend(plan.ops)
                self._plans[key] = dict(plan=allp, ws=None, io=io, graph=None, warm=0, parts=parts)
            self._plans[key]["bytes"] = self._plan_bytes(self._plans[key])
            self._evict()
        return self._plans[key]

    @staticmethod
    def _plan_bytes(ent):
        seen, total = set(), 0
        wss = [ent["ws"]] if ent["ws"] is not None else [pt["ws"] for pt in ent["parts"]]
        for ws in wss + [ent["io"]]:
            flat = []
            for t in ws.values():
                flat += list(t) if isinstance(t, (list, tuple)) else [t]
            for t in flat:
                st = t.untyped_storage()
                if st.data_ptr() not in seen:
                    seen.add(st.data_ptr())
                    total += st.nbytes()
        return total

    def _evict(self, reserve=0):
        """Drops least-recently-used plans until the cache holds at most max_plans - reserve plans and max_plan_bytes of
        workspace (the newest plan always stays).  The dropped plan's graph and buffers are released once the work already
        queued on them has run (stream-ordered free of torch's caching allocator)."""
        while len(self._plans) > 1 and (len(self._plans) > self.max_plans - reserve or
                                        sum(e.get("bytes", 0) for e in self._plans.values()) > self.max_plan_bytes):
            _, old = self._plans.popitem(last=False)
            if self._last_done is not None:
                self._last_done.synchronize()          # a captured graph must not be destroyed while a replay is in flight
            old.clear()

    def plan_cache_info(self):
        return {"plans": len(self._plans), "bytes": sum(e.get("bytes", 0) for e in self._plans.values()), "keys": list(self._plans)}

    def _run_plan(self, ent):
        if not ent["parts"]:
            ent["plan"].run()
            return
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        for pt in ent["parts"]:
            pt["stream"].wait_event(fork)
            pt["plan"].run(C.c_void_p(pt["stream"].cuda_stream))
            done = torch.cuda.Event()
            done.record(pt["stream"])
            main.wait_event(done)

    def begin_forward(self):
        """Called (under self._lock, inside torch.cuda.device(self.dev)) before a forward touches a plan's buffers: if the
        previous forward of this engine was issued on another stream, the current stream waits for it."""
        cur = torch.cuda.current_stream(self.dev)
        if self._last_done is not None and self._last_stream != cur.cuda_stream:
            cur.wait_event(self._last_done)

    def end_forward(self):
        cur = torch.cuda.current_stream(self.dev)
        if self._last_done is None:
            self._last_done = torch.cuda.Event()
        self._last_done.record(cur)
        self._last_stream = cur.cuda_stream

    def _run(self, ent):
        """Runs the plan on the current stream; the first call runs the op list eagerly (one-time kernel attribute
        setup happens there) and then captures it into a CUDA graph (launch-bound: ~500 small kernels per LNet forward);
        later calls replay the graph.  Capturing on the first call matters for the pipeline's tail batches, whose
        plans run once per clip."""
        if not self.use_graph:
            self._run_plan(ent)
            return
        if ent["graph"] is None:
            self._run_plan(ent)
            ent["warm"] += 1
            if ent["warm"] >= 1 and not torch.cuda.is_current_stream_capturing():
                torch.cuda.synchronize(self.dev)
                g = torch.cuda.CUDAGraph()
                # an explicit capture stream on THIS engine's device: torch's default capture stream is created once, on
                # whichever device was current first, and launching device-1 kernels into a device-0 stream fails
                if getattr(self, "_capture_stream", None) is None:
                    self._capture_stream = torch.cuda.Stream(device=self.dev)
                with torch.cuda.graph(g, stream=self._capture_stream):
                    self._run_plan(ent)
                ent["graph"] = g
            return
        ent["graph"].replay()

    def launches_per_forward(self, key):
        return len(self._plans[key]["plan"]) if key in self._plans else 0
