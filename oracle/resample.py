"""numpy restatement of the reference's ``load_wav`` (TEST INFRASTRUCTURE).

``futils/audio.py:9-10``: ``librosa.core.load(path, sr=sr)[0]``.  The arithmetic lives in third-party code that is absent
from /root/reference and from this image: **librosa==0.9.2** (``requirements.txt:6``) -> ``soundfile`` (PCM decode to float32)
-> ``librosa.to_mono`` (mean over channels) -> ``librosa.resample(res_type='kaiser_best')`` -> **resampy** (un-pinned,
``resampy>=0.2.2`` is librosa's requirement; restated here in its 0.4.x form: output times ``t * (sr_orig / sr_new)``).
Published algorithm (J. O. Smith's band-limited interpolation, resampy/core.py + resampy/interpn.py + resampy/filters.py):

* filter ``kaiser_best`` = ``sinc_window(num_zeros=64, precision=9, rolloff=0.9475937167399596, window=kaiser(beta=14.769656459379492))``:
  the right half of a Kaiser-windowed sinc sampled 512 times per zero crossing (32 769 taps), times ``sr_new/sr_orig`` when down-sampling;
* output sample t at input time ``tau = t * sr_orig / sr_new``: ``n = int(tau)``, left wing over x[n], x[n-1], ... and right wing over
  x[n+1], ... with the table read at ``offset + i * index_step`` and linearly interpolated (``interp_delta``) by the fractional index;
* output length ``int(n_in * ratio)``; ``librosa.resample`` then pads / trims to ``ceil(n_in * ratio)`` (``util.fix_length``).

**Parity unpinned**: neither librosa nor resampy can be installed here and the reference holds no vectors for this call; the
restatement is checked for self-consistency only (identity at equal rates, band-limited sine reconstruction, the length rule).
resampy accumulates into a float32 output array; this restatement accumulates in float64 and rounds once (difference <= 1e-6).
"""
from __future__ import annotations

import numpy as np

NUM_ZEROS = 64
PRECISION = 9
ROLLOFF = 0.9475937167399596
KAISER_BETA = 14.769656459379492

_cache: dict = {}


def kaiser_best():
    """(interp_win float64 [32769], num_table = 512) - resampy.filters.sinc_window with the 'kaiser_best' parameters."""
    if "w" not in _cache:
        num_bits = 2 ** PRECISION
        n = num_bits * NUM_ZEROS
        sinc_win = ROLLOFF * np.sinc(ROLLOFF * np.linspace(0, NUM_ZEROS, num=n + 1, endpoint=True))
        taper = np.kaiser(2 * n + 1, KAISER_BETA)[n:]          # scipy.signal.windows.kaiser(sym=True) == np.kaiser
        _cache["w"] = (taper * sinc_win, num_bits)
    return _cache["w"]


def resample(x: np.ndarray, sr_orig: int, sr_new: int) -> np.ndarray:
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_best') for a 1-D signal -> float32 [int(len * ratio)]."""
    x = np.asarray(x, dtype=np.float64)
    ratio = float(sr_new) / float(sr_orig)
    n_out = int(x.shape[0] * ratio)
    if n_out < 1:
        raise ValueError("input signal is too short to resample")
    win, num_table = kaiser_best()
    win = win * ratio if ratio < 1 else win
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    scale = min(1.0, ratio)
    index_step = int(scale * num_table)
    nwin, n_orig = win.shape[0], x.shape[0]
    t_reg = np.arange(n_out) * (1.0 / ratio)
    n = t_reg.astype(np.int64)
    y = np.zeros(n_out, dtype=np.float64)
    for wing in (0, 1):
        frac = scale * (t_reg - n)
        if wing:
            frac = scale - frac
        index_frac = frac * num_table
        offset = index_frac.astype(np.int64)
        eta = index_frac - offset
        cnt = np.minimum(n + 1 if wing == 0 else n_orig - n - 1, (nwin - offset) // index_step)
        for i in range(int(cnt.max()) if cnt.size else 0):
            ok = i < cnt
            idx = np.where(ok, offset + i * index_step, 0)
            src = np.where(ok, n - i if wing == 0 else n + i + 1, 0)
            w = win[idx] + eta * delta[idx]
            y += np.where(ok, w * x[src], 0.0)
    return y.astype(np.float32)


def pcm_to_float_mono(data: np.ndarray) -> np.ndarray:
    """soundfile's float32 decode of PCM samples + librosa.to_mono: int16 / 2^15, int32 / 2^31, uint8 (x - 128) / 2^7."""
    data = np.asarray(data)
    if data.dtype == np.int16:
        y = data.astype(np.float32) / np.float32(32768.0)
    elif data.dtype == np.int32:
        y = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        y = (data.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)
    else:
        y = data.astype(np.float32)
    if y.ndim == 2:
        y = y.mean(axis=1, dtype=np.float32) if y.shape[1] > 1 else y[:, 0]
    return np.ascontiguousarray(y, dtype=np.float32)


def load_array(data: np.ndarray, sr_native: int, sr: int) -> np.ndarray:
    """librosa.load on already-read samples: decode, mono, resample, fix_length(ceil(n * ratio)); float32."""
    y = pcm_to_float_mono(data)
    if sr_native == sr:
        return y
    out = resample(y, sr_native, sr)
    n_samples = int(np.ceil(y.shape[0] * float(sr) / sr_native))
    if out.shape[0] < n_samples:
        out = np.pad(out, (0, n_samples - out.shape[0]))
    return out[:n_samples].astype(np.float32)


def load_wav(path: str, sr: int) -> np.ndarray:
    """futils/audio.py:9-10."""
    from scipy.io import wavfile
    sr_native, data = wavfile.read(path)
    return load_array(data, int(sr_native), int(sr))
