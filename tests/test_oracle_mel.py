"""Pins the mel restatement (oracle/mel.py) against independent implementations
(torch.stft, torchaudio's Slaney filterbank) and the window-index rule against
the literal loop of inference.py:209-216 and the committed counts."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import mel as omel
from oracle import synth

from conftest import GOLDEN


def test_preemphasis_matches_lfilter():
    from scipy import signal
    x = synth.wav(0.5, seed=3)
    ref = signal.lfilter([1, -0.97], [1], x)
    np.testing.assert_allclose(omel.preemphasis(x), ref, rtol=0, atol=1e-15)


def test_stft_matches_torch_stft():
    y = omel.preemphasis(synth.wav(1.0, seed=1))
    d = omel.stft(y)
    t = torch.stft(torch.from_numpy(y), n_fft=800, hop_length=200, win_length=800,
                   window=torch.hann_window(800, periodic=True, dtype=torch.float64),
                   center=True, pad_mode="constant", return_complex=True).numpy()
    assert d.shape == (401, 1 + len(y) // 200) == t.shape
    assert np.abs(d - t).max() < 1e-10


def test_mel_basis_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(401, 55.0, 7600.0, 80, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    w = omel.mel_basis()
    assert w.dtype == np.float32 and w.shape == (80, 401)
    assert np.abs(w - fb).max() < 1e-6
    assert (w.sum(1) > 0).all()


def test_melspectrogram_shape_range_and_golden():
    wav = synth.wav(1.0, seed=0)
    m = omel.melspectrogram(wav)
    assert m.dtype == np.float64 and m.shape == (80, 81)
    assert m.min() >= -4 and m.max() <= 4
    gold = np.load(os.path.join(GOLDEN, "mel_oracle_seed0_1s.npy"))
    np.testing.assert_allclose(m, gold, atol=1e-5)


def test_melspectrogram_empty_and_short():
    assert omel.melspectrogram(np.zeros(0, np.float32)).shape == (80, 1)
    m = omel.melspectrogram(np.zeros(199, np.float32))
    assert m.shape == (80, 1) and np.all(m == -4.0)     # silence clips at -max_abs


def _literal_loop(n_cols, fps):
    # the loop of inference.py:209-216, kept literal on purpose
    mel_step_size, mel_idx_multiplier, i, starts = 16, 80. / fps, 0, []
    while True:
        start_idx = int(i * mel_idx_multiplier)
        if start_idx + mel_step_size > n_cols:
            starts.append(n_cols - mel_step_size)
            break
        starts.append(start_idx)
        i += 1
    return starts


@settings(max_examples=200, deadline=None)
@given(st.integers(16, 6000), st.sampled_from([23.976, 24.0, 25.0, 29.97, 30.0, 50.0, 60.0]))
def test_window_starts_property(n_cols, fps):
    s = omel.mel_window_starts(n_cols, fps)
    assert s == _literal_loop(n_cols, fps)
    assert all(0 <= a and a + 16 <= n_cols for a in s)
    assert s[-1] == n_cols - 16
    assert all(b >= a for a, b in zip(s[:-2], s[1:-1]))


def test_window_counts_committed():
    with open(os.path.join(GOLDEN, "mel_window_counts.json")) as f:
        g = json.load(f)
    assert g["counts"] == {"5": 122, "60": 1497, "600": 14997}
    assert omel.mel_window_starts(401) == g["starts_5s"]
    assert g["starts_5s"][:7] == [0, 3, 6, 9, 12, 16, 19] and g["starts_5s"][-3:] == [380, 384, 385]


def test_mel_windows_layout():
    m = omel.melspectrogram(synth.wav(1.0, seed=0))
    w = omel.mel_windows(m)
    assert w.dtype == np.float32 and w.shape[1:] == (1, 80, 16)
    s = omel.mel_window_starts(m.shape[1])
    np.testing.assert_array_equal(w[3, 0], m[:, s[3]:s[3] + 16].astype(np.float32))
