"""Drop-in for the reference's futils/audio.py front end, backed by libs2v's CUDA kernels.

    load_wav(path, sr)  -> float32 numpy [n]                          (futils/audio.py:9-10, librosa.core.load)
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


# ---- load_wav: PCM decode + librosa's default resampler (resampy 'kaiser_best') -------------------------------------------------
_KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)   # resampy/filters.py
_filter_cache: dict = {}
_PCM_KIND = {np.dtype(np.int16): 0, np.dtype(np.int32): 1, np.dtype(np.uint8): 2, np.dtype(np.float32): 3}


def _kaiser_best_table(device: torch.device, ratio: float):
    """resampy.filters.sinc_window for 'kaiser_best' (right half of a Kaiser-windowed sinc, 512 samples per zero crossing),
    scaled by the ratio when down-sampling, plus its forward differences - float64 device tensors."""
    key = (device.index, ratio if ratio < 1 else 1.0)
    if key not in _filter_cache:
        k = _KAISER_BEST
        num_bits = 2 ** k["precision"]
        n = num_bits * k["num_zeros"]
        win = k["rolloff"] * np.sinc(k["rolloff"] * np.linspace(0, k["num_zeros"], num=n + 1, endpoint=True)) * np.kaiser(2 * n + 1, k["beta"])[n:]
        if ratio < 1:
            win = win * ratio
        delta = np.zeros_like(win)
        delta[:-1] = np.diff(win)
        _filter_cache[key] = (torch.from_numpy(win).to(device), torch.from_numpy(delta).to(device), num_bits)
    return _filter_cache[key]


def resample_device(y: torch.Tensor, orig_sr: int, target_sr: int) -> torch.Tensor:
    """librosa.resample(y, orig_sr, target_sr) with its default res_type='kaiser_best' (what librosa.load calls) on a 1-D float32
    CUDA tensor: resampy's int(n * ratio) samples, padded / trimmed to ceil(n * ratio) (util.fix_length)."""
    if not (y.is_cuda and y.dim() == 1):
        raise ValueError("y must be a 1-D CUDA tensor")
    y = y.contiguous().float()
    if orig_sr == target_sr:
        return y
    lib = L.require_device(y.device.index)
    ratio = float(target_sr) / float(orig_sr)
    n_res = int(lib.s2v_resample_out_len(y.numel(), int(orig_sr), int(target_sr)))
    if n_res < 1:
        raise ValueError("Input signal length=%d is too small to resample from %d->%d" % (y.numel(), orig_sr, target_sr))
    n_out = int(np.ceil(y.numel() * ratio))
    win, delta, num_table = _kaiser_best_table(y.device, ratio)
    out = torch.zeros(n_out, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        L.check(lib.s2v_resample_f32(y.data_ptr(), y.numel(), int(orig_sr), int(target_sr), win.data_ptr(), delta.data_ptr(), win.numel(),
                                     num_table, out.data_ptr(), min(n_res, n_out), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                "s2v_resample_f32")
    return out


def load_array_device(data, sr_native: int, sr: int, device=None) -> torch.Tensor:
    """librosa.load behind the file read: PCM samples [n] or [n, channels] (int16 / int32 / uint8 / float32 numpy) at
    ``sr_native`` -> float32 mono CUDA tensor at ``sr``."""
    data = np.ascontiguousarray(data)
    if data.dtype == np.float64:
        data = data.astype(np.float32)
    if data.dtype not in _PCM_KIND:
        raise TypeError("unsupported PCM dtype %s" % data.dtype)
    if not torch.cuda.is_available():
        raise L.S2VError("a CUDA device is required: this package has no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = L.require_device(dev.index)
    n = data.shape[0]
    ch = data.shape[1] if data.ndim == 2 else 1
    out = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        # torch has no uint16/etc. issue here: the bytes are shipped as they are and decoded on the device
        raw = torch.from_numpy(data.reshape(-1).view(np.uint8)).to(dev)
        with torch.cuda.device(dev):
            L.check(lib.s2v_pcm_to_mono_f32(raw.data_ptr(), _PCM_KIND[data.dtype], ch, n, out.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "s2v_pcm_to_mono_f32")
    return resample_device(out, int(sr_native), int(sr))


def load_wav(path, sr):
    """futils/audio.py:9-10: ``librosa.core.load(path, sr=sr)[0]`` - float32 numpy, mono, resampled to ``sr``.  The file is read on
    the host (scipy.io.wavfile; the reference hands every non-wav input to ffmpeg first, inference.py:200-203); decode, channel
    mix and the resampler run on the GPU."""
    from scipy.io import wavfile
    sr_native, data = wavfile.read(path)
    return load_array_device(data, int(sr_native), int(sr)).cpu().numpy()


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
