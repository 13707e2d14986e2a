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
