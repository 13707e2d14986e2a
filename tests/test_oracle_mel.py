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


# ---- goldens produced by the UNMODIFIED reference futils/audio.py (oracle/make_golden_mel.py; stub librosa) ------------------
def _ref_golden():
    return np.load(os.path.join(GOLDEN, "mel_ref_golden.npz"))


def test_oracle_vs_reference_audio_py_goldens():
    """a1 / a4 + the glue of a2 / a3 pinned to the reference's own file (STFT / mel basis stay unpinned at librosa)."""
    from oracle import resample
    g = _ref_golden()
    np.testing.assert_allclose(omel.preemphasis(synth.wav(1.0, seed=0)), g["preemph_synth"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(omel.melspectrogram(synth.wav(1.0, seed=0)), g["mel_synth_seed0_1s"], rtol=0, atol=1e-12)
    speech = resample.pcm_to_float_mono(g["speech_pcm"])
    m = omel.melspectrogram(speech)
    assert m.shape == g["mel_speech"].shape == (80, 321)
    np.testing.assert_allclose(m, g["mel_speech"], rtol=0, atol=1e-12)
    assert (g["mel_speech"] <= -4).mean() > 0.1          # real speech reaches the lower clip rail (white noise never does)


def test_reference_audio_py_importable_with_stub_librosa():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree only exists in the build container")
    a = ref_shim.load_audio()
    w = synth.wav(0.5, seed=5)
    np.testing.assert_array_equal(a.melspectrogram(w), omel.melspectrogram(w))
    assert a.get_hop_size() == 200 and a.hp.num_mels == 80
    import sys
    assert "librosa" not in sys.modules or not hasattr(sys.modules["librosa"], "_s2v_stub")


# ---- load_wav restatement (oracle/resample.py; parity unpinned: librosa / resampy absent) ---------------------------------------
def test_resample_identity_and_length_rule():
    from oracle import resample as R
    x = synth.wav(0.5, seed=2)
    np.testing.assert_array_equal(R.load_array(x, 16000, 16000), x)                     # equal rates: no resampling at all
    for sr0, n in ((44100, 4410), (48000, 4801), (22050, 2207), (8000, 801)):
        y = R.load_array(x[:n], sr0, 16000)
        assert y.dtype == np.float32 and y.shape[0] == int(np.ceil(n * 16000.0 / sr0))   # librosa.resample: fix_length(ceil(n * ratio))


def test_resample_reconstructs_band_limited_sine():
    from oracle import resample as R
    for sr0, tol in ((8000, 1e-6), (22050, 1e-3), (44100, 4e-3), (48000, 4e-3)):        # int(index_step) truncation = resampy's known gain error when down-sampling
        t = np.arange(sr0 // 2) / sr0
        y = R.resample(np.sin(2 * np.pi * 440 * t), sr0, 16000).astype(np.float64)
        ref = np.sin(2 * np.pi * 440 * np.arange(len(y)) / 16000)
        assert np.abs(y - ref)[1000:-1000].max() < tol


def test_pcm_decode():
    from oracle import resample as R
    pcm = np.array([[-32768, 32767], [0, 16384]], dtype=np.int16)
    np.testing.assert_array_equal(R.pcm_to_float_mono(pcm), np.array([(-1.0 + 32767 / 32768) / 2, 0.25], dtype=np.float32))
    np.testing.assert_array_equal(R.pcm_to_float_mono(np.array([0, 128, 255], dtype=np.uint8)), np.array([-1, 0, 127 / 128], dtype=np.float32))
