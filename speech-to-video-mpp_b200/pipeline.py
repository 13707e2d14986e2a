"""Full per-frame path on one GPU for a contiguous frame range: mel -> windows -> DNet -> glue -> LNet
(BASELINE.json configs[3]/[4]).  Mirrors the order of the reference's inference.py (:204-216 mel +
windows, preprocessing/facing.py:176-194 DNet per frame, inference.py:259-267 LNet batches) with the
per-frame CPU<->GPU round trips removed: everything stays in HBM between stages.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from . import parallel
from .futils import audio, inference_utils


def glue_fake_to_face(fake: torch.Tensor, size: int = 96) -> torch.Tensor:
    """fake [B,3,H,W] in [-1,1] -> LNet face input [B,6,size,size] (synthetic glue)."""
    fake = fake.contiguous().float()
    b, c, h, w = fake.shape
    lib = L.require_device(fake.device.index)
    out = torch.empty(b, 2 * c, size, size, dtype=torch.float32, device=fake.device)
    with torch.cuda.device(fake.device):
        L.check(lib.s2v_glue_fake_to_face_f32(fake.data_ptr(), out.data_ptr(), b, c, h, w, size, size, size // 2,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s2v_glue_fake_to_face_f32")
    return out


class LipSyncPipeline:
    def __init__(self, lnet, dnet, lnet_batch: int = 128, dnet_batch: int = 64, fps: float = 25.0):
        self.lnet, self.dnet, self.lb, self.db, self.fps = lnet, dnet, lnet_batch, dnet_batch, fps

    def n_frames(self, n_samples: int) -> int:
        return audio.mel_window_count(1 + n_samples // 200, self.fps)

    @torch.no_grad()
    def run(self, wav: torch.Tensor, sources: torch.Tensor, coeffs: torch.Tensor | None, rank: int = 0, world: int = 1,
            semantic: torch.Tensor | None = None, crop_norm_ratio=None):
        """wav: float32 CUDA [n_samples]; sources [N,3,256,256], coeffs [N,73,26] for THIS rank's frame
        range (or all N frames when world == 1).  Instead of ``coeffs``, ``semantic`` may hold the clip's whole 3DMM
        coefficient table [T,262] on the device (+ ``crop_norm_ratio``): the [73,26] windows of this rank's frames are
        then built per DNet batch by s2v_semantic_windows (futils/inference_utils.py:78-91 of the reference, which runs
        it per frame on the CPU at preprocessing/facing.py:184).  Returns this rank's generated frames [n_r,3,96,96]."""
        mel = audio.melspectrogram_device(wav)
        total = audio.mel_window_count(mel.shape[1], self.fps)
        lo, hi = parallel.shard_range(total, rank, world)
        n = hi - lo
        assert sources.shape[0] >= n and (coeffs is None) != (semantic is None)
        assert coeffs is None or coeffs.shape[0] >= n
        windows = audio.mel_windows(mel, self.fps, lo, n)
        faces = torch.empty(n, 6, 96, 96, dtype=torch.float32, device=wav.device)
        for s in range(0, n, self.db):
            e = min(n, s + self.db)
            drive = coeffs[s:e] if coeffs is not None else inference_utils.semantic_windows(semantic, (lo + s, e - s), crop_norm_ratio)
            out = self.dnet(sources[s:e], drive)
            faces[s:e] = glue_fake_to_face(out["fake_image"])
        frames = torch.empty(n, 3, 96, 96, dtype=torch.float32, device=wav.device)
        for s in range(0, n, self.lb):
            e = min(n, s + self.lb)
            frames[s:e] = self.lnet(windows[s:e], faces[s:e])
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
