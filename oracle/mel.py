"""fp64 numpy restatement of the reference mel front end (TEST INFRASTRUCTURE).

Follows ``futils/audio.py`` (reference) with the constants of
``futils/hparams.py:21-61`` and the semantics of librosa 0.9.2 (the pinned,
un-vendored dependency that holds the STFT / mel-basis arithmetic):

* ``preemphasis``      futils/audio.py:20-23   (scipy.signal.lfilter([1,-k],[1]))
* ``stft``             futils/audio.py:57-61   (librosa.stft n_fft=800 hop=200 win=800,
                                                center=True, periodic Hann, zero pad)
* ``mel_basis``        futils/audio.py:98-103  (librosa.filters.mel: Slaney scale + norm, float32)
* ``amp_to_db`` / ``normalize``  futils/audio.py:104-117
* ``melspectrogram``   futils/audio.py:45-51
* ``mel_window_starts`` / ``mel_windows``  inference.py:209-216, :399, :261

Parity status: "parity unpinned" at the librosa boundary (librosa is absent and
the reference holds no golden vectors); cross-checked against torch.stft and
torchaudio's Slaney filterbank in tests/test_oracle_mel.py.
"""
from __future__ import annotations

import numpy as np

# futils/hparams.py:21-61
NUM_MELS = 80
N_FFT = 800
HOP = 200
WIN = 800
SR = 16000
PREEMPH = 0.97
MIN_LEVEL_DB = -100.0
REF_LEVEL_DB = 20.0
FMIN = 55.0
FMAX = 7600.0
MAX_ABS = 4.0
FPS = 25
MEL_STEP = 16          # inference.py:209  mel_step_size


def preemphasis(wav: np.ndarray, k: float = PREEMPH) -> np.ndarray:
    """y[n] = x[n] - k x[n-1], y[0] = x[0]; float64 (audio.py:20-23)."""
    x = np.asarray(wav, dtype=np.float64)
    y = x.copy()
    y[1:] -= k * x[:-1]
    return y


def hann_periodic(n: int = WIN) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y: np.ndarray, pad_mode: str = "constant") -> np.ndarray:
    """librosa.stft(y, n_fft=800, hop_length=200, win_length=800) -> complex128 [401, T].

    center=True pads n_fft//2 on both sides; librosa>=0.9 pads zeros
    ('constant'), <=0.8 reflected - parameterised, default zeros.
    T = 1 + len(y)//hop.
    """
    y = np.asarray(y, dtype=np.float64)
    yp = np.pad(y, N_FFT // 2, mode=pad_mode)
    t = 1 + (len(yp) - N_FFT) // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(t)[:, None]
    frames = yp[idx] * hann_periodic()[None, :]
    return np.fft.rfft(frames, n=N_FFT, axis=1).T


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_basis() -> np.ndarray:
    """librosa.filters.mel(sr=16000, n_fft=800, n_mels=80, fmin=55, fmax=7600) -> float32 [80, 401]."""
    assert FMAX <= SR // 2                       # audio.py:99
    n_bins = 1 + N_FFT // 2
    fftfreqs = np.linspace(0, SR / 2.0, n_bins)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(FMIN), _hz_to_mel(FMAX), NUM_MELS + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((NUM_MELS, n_bins), dtype=np.float64)
    for i in range(NUM_MELS):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:NUM_MELS + 2] - mel_f[:NUM_MELS])
    w *= enorm[:, None]
    return w.astype(np.float32)


def amp_to_db(x: np.ndarray) -> np.ndarray:
    min_level = np.exp(MIN_LEVEL_DB / 20 * np.log(10))      # audio.py:105
    return 20 * np.log10(np.maximum(min_level, x))


def normalize(s: np.ndarray) -> np.ndarray:
    """allow_clipping & symmetric branch (audio.py:112-115)."""
    return np.clip((2 * MAX_ABS) * ((s - MIN_LEVEL_DB) / (-MIN_LEVEL_DB)) - MAX_ABS, -MAX_ABS, MAX_ABS)


def melspectrogram(wav: np.ndarray, pad_mode: str = "constant") -> np.ndarray:
    """audio.py:45-51 -> float64 [80, 1 + len//200]."""
    d = stft(preemphasis(wav), pad_mode=pad_mode)
    s = amp_to_db(np.dot(mel_basis(), np.abs(d))) - REF_LEVEL_DB
    return normalize(s)


def mel_window_starts(n_cols: int, fps: float = FPS) -> list[int]:
    """Start columns of the 80x16 windows, exactly as inference.py:209-216.

    ``int(i * (80./fps))`` is a float64 product truncated toward zero; when
    ``start + 16 > T`` the tail window ``T-16`` is appended once and the loop
    stops.
    """
    mult = 80.0 / fps
    starts = []
    i = 0
    while True:
        s = int(i * mult)
        if s + MEL_STEP > n_cols:
            starts.append(n_cols - MEL_STEP)
            break
        starts.append(s)
        i += 1
    return starts


def mel_windows(mel: np.ndarray, fps: float = FPS) -> np.ndarray:
    """[80,T] -> float32 [N,1,80,16] (inference.py:209-216 + :399 + :261)."""
    starts = mel_window_starts(mel.shape[1], fps)
    out = np.stack([mel[:, s:s + MEL_STEP] for s in starts], 0)
    return out[:, None].astype(np.float32)
