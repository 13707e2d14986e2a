"""Full per-frame path on one GPU for a contiguous frame range: mel -> windows -> DNet -> glue -> LNet
(BASELINE.json configs[3]/[4]).  Mirrors the order of the reference's inference.py (:204-216 mel +
windows, preprocessing/facing.py:176-194 DNet per frame, inference.py:259-267 LNet batches) with the
per-frame CPU<->GPU round trips removed: everything stays in HBM between stages.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib as L
from . import parallel
from .futils import audio, inference_utils


def glue_fake_to_face(fake: torch.Tensor, size: int = 96, out: torch.Tensor | None = None) -> torch.Tensor:
    """fake [B,3,H,W] in [-1,1] -> LNet face input [B,6,size,size] (synthetic glue); ``out``: a
    contiguous float32 [B,6,size,size] destination (e.g. a row range of the clip's face tensor)."""
    fake = fake.contiguous().float()
    b, c, h, w = fake.shape
    lib = L.require_device(fake.device.index)
    if out is None:
        out = torch.empty(b, 2 * c, size, size, dtype=torch.float32, device=fake.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (b, 2 * c, size, size)
    with torch.cuda.device(fake.device):
        L.check(lib.s2v_glue_fake_to_face_f32(fake.data_ptr(), out.data_ptr(), b, c, h, w, size, size, size // 2,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s2v_glue_fake_to_face_f32")
    return out


def balanced_batches(n: int, cap: int) -> list:
    """n frames -> batch sizes: full batches of ``cap`` and one tail.  (Equal-size batches - 188 -> 94 + 94 - were measured on
    B200 and lose: the engines run batches on multiple-of-8 plan buckets, so 1 497 frames as 12 x 125 pad to 12 x 128 = 2.6 %
    more work than 11 x 128 + 89 -> 96, and the per-batch fixed cost is the same either way: 3 611 vs 3 746 frames/s.)  A clip
    of any length touches at most two plan sizes per network."""
    if n <= 0:
        return []
    full, tail = divmod(n, cap)
    return [cap] * full + ([tail] if tail else [])


_PIPE_STREAMS = {}


def _pipe_streams(dev):
    """(DNet stream, LNet stream, copy-in stream, copy-out stream) of a device, created once."""
    key = (dev.type, dev.index)
    if key not in _PIPE_STREAMS:
        _PIPE_STREAMS[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev),
                              torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev, priority=-1))
    return _PIPE_STREAMS[key]


class LipSyncPipeline:
    """mel -> windows -> DNet -> glue -> LNet for one rank's contiguous frame range.

    The two networks run on TWO streams: while LNet consumes the faces of batch k, DNet already produces batch k+1 (the
    reference runs them strictly one after the other, frame by frame, preprocessing/facing.py:176-194 then inference.py:259-267).
    Each network is a chain of ~300-500 dependent launches with a prologue / tail bubble per launch; two independent chains
    fill each other's bubbles.  ``overlap=False`` (or S2V_PIPE_OVERLAP=0) keeps everything on the caller's stream.
    Inputs may live in HBM or in PINNED host memory (``sources`` / ``coeffs`` on the CPU): host inputs are staged batch by
    batch on a copy stream, and ``out_host`` (pinned) receives the frames batch by batch on another - the form bench.py's
    end-to-end number uses.
    Batch caps (measured on B200, 600 s clip on 8 GPUs, gather timed): LNet 64 / 128 / 256 / 512 -> 27.5 / 29.4 / 29.8 / 29.7 k
    frames/s; a rank of the 60 s clip at N = 8 owns 188 frames: one batch of 192 instead of 128 + 64 is 29.3 k vs 28.0 k frames/s.
    DNet (1.3 ms fixed + 125 us per frame): caps 64 / 128 / 192 -> 3 790 / 3 851 / 3 876 frames/s on the 60 s clip at N = 1
    (plan workspaces 9.7 / 19.5 / 29 GB); 192 also makes a rank's 188 frames at N = 8 a single DNet batch."""

    def __init__(self, lnet, dnet, lnet_batch: int = 256, dnet_batch: int = 192, fps: float = 25.0, overlap: bool | None = None):
        self.lnet, self.dnet, self.lb, self.db, self.fps = lnet, dnet, lnet_batch, dnet_batch, fps
        self.overlap = (os.environ.get("S2V_PIPE_OVERLAP", "1") == "1") if overlap is None else overlap
        self.host_first = int(os.environ.get("S2V_HOST_FIRST", "48"))      # frames of the first DNet batch when the inputs are on the host
        self._stage = {}

    def n_frames(self, n_samples: int) -> int:
        return audio.mel_window_count(1 + n_samples // 200, self.fps)

    def _staging(self, dev, slot, b, t):
        key = (slot, b, t)
        if key not in self._stage:
            self._stage[key] = (torch.empty(b, 3, 256, 256, dtype=torch.float32, device=dev),
                                torch.empty(b, 73, t, dtype=torch.float32, device=dev), torch.cuda.Event(), torch.cuda.Event())
        return self._stage[key]

    @torch.no_grad()
    def run(self, wav: torch.Tensor, sources: torch.Tensor, coeffs: torch.Tensor | None, rank: int = 0, world: int = 1,
            semantic: torch.Tensor | None = None, crop_norm_ratio=None, out_host: torch.Tensor | None = None):
        """wav: float32 [n_samples] (CUDA, or pinned host); sources [N,3,256,256], coeffs [N,73,26] for THIS rank's frame
        range (or all N frames when world == 1), CUDA or pinned host.  Instead of ``coeffs``, ``semantic`` may hold the clip's
        whole 3DMM coefficient table [T,262] on the device (+ ``crop_norm_ratio``): the [73,26] windows of this rank's frames are
        then built per DNet batch by s2v_semantic_windows (futils/inference_utils.py:78-91 of the reference, which runs
        it per frame on the CPU at preprocessing/facing.py:184).  Returns this rank's generated frames [n_r,3,96,96] on the
        device; with ``out_host`` (pinned [n_r,3,96,96]) they are also copied there, complete when the call returns."""
        dev = next(self.lnet.parameters()).device
        main = torch.cuda.current_stream(dev)
        if not wav.is_cuda:
            wav = wav.to(dev, non_blocking=True)
        mel = audio.melspectrogram_device(wav)
        total = audio.mel_window_count(mel.shape[1], self.fps)
        lo, hi = parallel.shard_range(total, rank, world)
        n = hi - lo
        assert sources.shape[0] >= n and (coeffs is None) != (semantic is None)
        assert coeffs is None or coeffs.shape[0] >= n
        windows = audio.mel_windows(mel, self.fps, lo, n)
        faces = torch.empty(n, 6, 96, 96, dtype=torch.float32, device=dev)
        frames = torch.empty(n, 3, 96, 96, dtype=torch.float32, device=dev)
        if n == 0:
            return frames
        host_in = not sources.is_cuda
        if self.overlap:
            s_d, s_l, s_in, s_out = _pipe_streams(dev)
            for s in (s_d, s_l, s_in, s_out):
                s.wait_stream(main)
        else:
            s_d = s_l = s_in = s_out = main
        deng, leng = self.dnet.engine(), self.lnet.engine()
        # ---- DNet batches (stream s_d), host inputs staged one batch ahead on s_in -----------------------------------------
        d_sizes, l_sizes = balanced_batches(n, self.db), balanced_batches(n, self.lb)
        if host_in and n > 2 * self.host_first:
            # host inputs: nothing can run before the first batch's sources have crossed PCIe (0.79 MB per frame), so the first
            # DNet batch is small; every later batch's copy hides behind the previous batch's forward
            d_sizes = [self.host_first] + balanced_batches(n - self.host_first, self.db)
        d_done = []                                     # (last frame + 1, event) per DNet batch
        pos = 0
        staged = None
        bounds = []
        for b in d_sizes:
            bounds.append((pos, pos + b))
            pos += b

        def stage(j):
            s, e = bounds[j]
            t = coeffs.shape[2] if coeffs is not None else 26
            img_d, co_d, ready, freed = self._staging(dev, j & 1, max(d_sizes), t)
            with torch.cuda.stream(s_in):
                if j >= 2:
                    s_in.wait_event(freed)              # the DNet forward that read this slot two batches ago has copied it
                img_d[:e - s].copy_(sources[s:e], non_blocking=True)
                if coeffs is not None:
                    co_d[:e - s].copy_(coeffs[s:e], non_blocking=True)
                ready.record(s_in)
            return img_d[:e - s], (co_d[:e - s] if coeffs is not None else None), ready, freed

        if host_in:
            staged = stage(0)
        l_bounds, pos = [], 0
        for b in l_sizes:
            l_bounds.append((pos, pos + b))
            pos += b
        li = 0
        out_events = []
        for j, (s, e) in enumerate(bounds):
            if host_in:
                img, co, ready, freed = staged
                if j + 1 < len(bounds):
                    staged = stage(j + 1)
            else:
                img, co, ready, freed = sources[s:e], (coeffs[s:e] if coeffs is not None else None), None, None
            with torch.cuda.stream(s_d):
                if ready is not None:
                    s_d.wait_event(ready)
                drive = co if co is not None else inference_utils.semantic_windows(semantic, (lo + s, e - s), crop_norm_ratio)
                fake = deng.forward(img, drive, only="fake_image")["fake_image"]
                if freed is not None:
                    freed.record(s_d)                   # the engine has copied the staged inputs into its own buffers
                glue_fake_to_face(fake, out=faces[s:e])
                ev = torch.cuda.Event()
                ev.record(s_d)
            d_done.append((e, ev))
            # ---- every LNet batch whose faces are now complete (stream s_l) ------------------------------------------------
            while li < len(l_bounds) and l_bounds[li][1] <= e:
                ls, le = l_bounds[li]
                with torch.cuda.stream(s_l):
                    if s_l is not s_d:
                        s_l.wait_event(ev)
                    leng.forward(windows[ls:le], faces[ls:le], out=frames[ls:le])
                    if out_host is not None:
                        lev = torch.cuda.Event()
                        lev.record(s_l)
                        with torch.cuda.stream(s_out):
                            s_out.wait_event(lev)
                            out_host[ls:le].copy_(frames[ls:le], non_blocking=True)
                li += 1
        assert li == len(l_bounds)
        if self.overlap:
            for s in (s_d, s_l, s_in, s_out):
                main.wait_stream(s)
        if out_host is not None:
            (s_out if self.overlap else main).synchronize()
        return frames


_COPY_STREAMS = {}


def _copy_streams(dev):
    """One (copy-in, copy-out) stream pair per device, created once, high priority.  Taking fresh streams from torch's
    round-robin pool on every call made the loop bimodal on B200 (13.9 or 18 ms/step for LNet B=128, depending on which
    pool streams came back - streams that share a hardware queue with the compute stream serialise the copies' event
    waits in front of the forward's kernels); a fixed high-priority pair comes from a different pool than the compute
    stream's neighbours."""
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = (torch.cuda.Stream(device=dev, priority=-1), torch.cuda.Stream(device=dev, priority=-1))
    return _COPY_STREAMS[key]


@torch.no_grad()
def stream_batches(net, batches, depth: int = 3):
    """Host-to-host batch loop with the copies overlapped with the forward passes.

    The reference's generation loop (inference.py:259-267, 288) handles one batch at a time: host arrays -> ``.to(device)``
    -> ``model(mel, img)`` -> ``.cpu()``.  This is the same loop with the three stages on three streams: while batch i runs
    on the compute stream, batch i+1's pinned inputs are already crossing PCIe on the copy-in stream and batch i-1's frames
    are going back on the copy-out stream.

    ``net``: a callable on device tensors (``LNet`` module: ``net(mel, face)``); ``batches``: an iterable of
    ``(inputs, out_host)`` where ``inputs`` is a tuple of PINNED host tensors and ``out_host`` a pinned host tensor that
    receives the result.  ``depth`` input / output device buffers are kept in flight (measured on B200, LNet B=128:
    13.7 ms device-only, 14.4 ms for the sequential loop, 14.0 ms with depth 3; depth 2 stalls at 16 ms).  Returns the number of batches;
    the outputs are complete when the function returns (it synchronises the copy-out stream).
    """
    dev = next(net.parameters()).device
    cur = torch.cuda.current_stream(dev)
    s_in, s_out = _copy_streams(dev)
    slots = [dict(inp=None, out=None, in_done=torch.cuda.Event(), run_done=torch.cuda.Event(), out_done=torch.cuda.Event(),
                  used=False) for _ in range(depth)]
    s_in.wait_stream(cur)
    n = 0
    pending = None                                  # (slot, out_host) whose forward has not been issued yet

    def issue_forward(slot, out_host):
        cur.wait_event(slot["in_done"])
        if slot["used"]:
            cur.wait_event(slot["out_done"])      # the slot's previous result has left the device
        res = net(*slot["inp"])
        if slot["out"] is None or slot["out"].shape != res.shape:
            slot["out"] = torch.empty_like(res)
        slot["out"].copy_(res)
        slot["run_done"].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(slot["run_done"])
            out_host.copy_(slot["out"], non_blocking=True)
            slot["out_done"].record(s_out)
        slot["used"] = True

    for inputs, out_host in batches:
        slot = slots[n % depth]
        with torch.cuda.stream(s_in):
            if slot["used"]:
                s_in.wait_event(slot["run_done"])    # the forward that read this slot's inputs has finished
            if slot["inp"] is None or any(a.shape != b.shape for a, b in zip(slot["inp"], inputs)):
                slot["inp"] = tuple(torch.empty(t.shape, dtype=t.dtype, device=dev) for t in inputs)
            for d, h in zip(slot["inp"], inputs):
                d.copy_(h, non_blocking=True)
            slot["in_done"].record(s_in)
        if pending is not None:                     # issue the previous batch's forward AFTER this batch's copy-in was queued
            issue_forward(*pending)
        pending = (slot, out_host)
        if depth == 1:                              # a single slot cannot be refilled before its forward has been issued
            issue_forward(*pending)
            pending = None
        n += 1
    if pending is not None:
        issue_forward(*pending)
    s_out.synchronize()
    return n
