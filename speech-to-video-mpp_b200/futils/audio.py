"""Drop-in for the reference's futils/audio.py mel front end, backed by libs2v's fused CUDA kernel.

    melspectrogram(wav) -> float64 numpy [80, 1 + len(wav)//200]      (futils/audio.py:45-51)

plus the window rule of inference.py:209-216 as functions (``mel_window_starts``, ``mel_windows``).
There is no CPU path: a CUDA device is required.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib as L
from .hparams import hparams as hp

_basis_cache: dict = {}


def get_hop_size():
    hop_size = hp.hop_size
    if hop_size is None:
        assert hp.frame_shift_ms is not None
        hop_size = int(hp.frame_shift_ms / 1000 * hp.sample_rate)
    return hop_size


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_hz / f_sp + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def _build_mel_basis():
    """Slaney-scale, Slaney-normalised triangular filterbank, float32 [num_mels, 1+n_fft//2]
    (what librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) returns; audio.py:98-103)."""
    assert hp.fmax <= hp.sample_rate // 2
    n_bins = 1 + hp.n_fft // 2
    fftfreqs = np.linspace(0, hp.sample_rate / 2.0, n_bins)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(hp.fmin), _hz_to_mel(hp.fmax), hp.num_mels + 2))
    lower = (fftfreqs[None, :] - edges[:-2, None]) / np.diff(edges)[:-1, None]
    upper = (edges[2:, None] - fftfreqs[None, :]) / np.diff(edges)[1:, None]
    w = np.maximum(0, np.minimum(lower, upper)) * (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return w.astype(np.float32)


def _device_basis(device: torch.device):
    key = (device.index, hp.num_mels, hp.n_fft, hp.sample_rate, hp.fmin, hp.fmax)
    if key not in _basis_cache:
        w = _build_mel_basis()
        rng = np.zeros((hp.num_mels, 2), dtype=np.int32)
        for m in range(hp.num_mels):
            nz = np.nonzero(w[m])[0]
            rng[m] = (nz[0], nz[-1] + 1) if len(nz) else (0, 0)
        _basis_cache[key] = (torch.from_numpy(w).to(device), torch.from_numpy(rng).to(device))
    return _basis_cache[key]


def _check_hparams():
    if (hp.n_fft, hp.win_size, get_hop_size(), hp.num_mels) != (800, 800, 200, 80) or hp.use_lws:
        raise ValueError("the CUDA mel kernel is specialised for n_fft=win_size=800, hop=200, num_mels=80, use_lws=False")
    if not (hp.preemphasize and abs(hp.preemphasis - 0.97) < 1e-12 and hp.signal_normalization and
            hp.allow_clipping_in_normalization and hp.symmetric_mels and hp.max_abs_value == 4. and
            hp.min_level_db == -100 and hp.ref_level_db == 20):
        raise ValueError("the CUDA mel kernel bakes the reference's normalisation constants (hparams.py:39-57)")


def melspectrogram_device(wav: torch.Tensor, pad_mode: str = "constant") -> torch.Tensor:
    """wav: 1-D float32 CUDA tensor -> float32 CUDA tensor [80, T] (stream-ordered, no sync)."""
    _check_hparams()
    if not (wav.is_cuda and wav.dim() == 1):
        raise ValueError("wav must be a 1-D CUDA tensor")
    wav = wav.contiguous().float()
    lib = L.require_device(wav.device.index)
    basis, band_range = _device_basis(wav.device)
    t = 1 + wav.numel() // 200
    out = torch.empty(hp.num_mels, t, dtype=torch.float32, device=wav.device)
    with torch.cuda.device(wav.device):
        L.check(lib.s2v_melspectrogram_f32(wav.data_ptr(), wav.numel(), basis.data_ptr(), band_range.data_ptr(),
                                           out.data_ptr(), 1 if pad_mode == "reflect" else 0,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s2v_melspectrogram_f32")
    return out


def melspectrogram(wav, pad_mode: str = "constant"):
    """Reference signature: 1-D float array -> float64 numpy [80, T] (audio.py:45-51).
    Also accepts a CUDA tensor, in which case a float32 CUDA tensor is returned."""
    if isinstance(wav, torch.Tensor) and wav.is_cuda:
        return melspectrogram_device(wav, pad_mode)
    x = torch.as_tensor(np.ascontiguousarray(np.asarray(wav, dtype=np.float32)))
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    if dev is None:
        raise L.S2VError("a CUDA device is required: this package has no CPU path")
    return melspectrogram_device(x.to(dev), pad_mode).cpu().numpy().astype(np.float64)


def mel_window_count(n_cols: int, fps: float = None) -> int:
    fps = hp.fps if fps is None else fps
    n = L.load_library().s2v_mel_window_count(int(n_cols), float(fps))
    if n < 0:
        raise ValueError("mel has fewer than 16 columns")
    return int(n)


def mel_window_starts(n_cols: int, fps: float = None) -> list:
    """Start column of every 80x16 window - the index list of inference.py:209-216, bit-exact."""
    fps = hp.fps if fps is None else fps
    n = mel_window_count(n_cols, fps)
    buf = (C.c_int32 * n)()
    L.check(L.load_library().s2v_mel_window_starts_host(int(n_cols), float(fps), buf, n), "s2v_mel_window_starts_host")
    return list(buf)


def mel_windows(mel: torch.Tensor, fps: float = None, first: int = 0, count: int = None) -> torch.Tensor:
    """mel [80,T] float32 CUDA -> float32 [count,1,80,16]: windows [first, first+count)
    (inference.py:209-216 + the [B,1,80,16] layout of :399/:261)."""
    fps = hp.fps if fps is None else fps
    if not (mel.is_cuda and mel.dim() == 2 and mel.shape[0] == 80 and mel.dtype == torch.float32):
        raise ValueError("mel must be a float32 CUDA tensor [80, T]")
    mel = mel.contiguous()
    total = mel_window_count(mel.shape[1], fps)
    count = total - first if count is None else count
    lib = L.require_device(mel.device.index)
    out = torch.empty(count, 1, 80, 16, dtype=torch.float32, device=mel.device)
    with torch.cuda.device(mel.device):
        L.check(lib.s2v_mel_windows_f32(mel.data_ptr(), mel.shape[1], float(fps), first, count, out.data_ptr(),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s2v_mel_windows_f32")
    return out
