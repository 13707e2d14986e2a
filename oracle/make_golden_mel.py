"""Generates the mel goldens from the UNMODIFIED reference ``futils/audio.py`` (TEST INFRASTRUCTURE; build container only).

    python -m oracle.make_golden_mel

``oracle/ref_shim.load_audio`` imports the reference file with a stub ``librosa`` whose ``stft`` / ``filters.mel`` are the
oracle's restatements (librosa 0.9.2 itself cannot be installed here).  What the goldens pin to the reference itself:
``preemphasis`` (scipy's lfilter, audio.py:20-23), ``_linear_to_mel`` / ``_amp_to_db`` / ``_normalize`` (:92-117), the order of
operations of ``melspectrogram`` (:45-51) and the ``hparams`` binding (:7) - rows a1 and a4 of SURVEY section 8 and the glue of
a2 / a3.  The STFT and mel-basis arithmetic stay "unpinned at librosa".

Inputs: the seeded 1 s synthetic wav (oracle/synth.py) and a 4 s excerpt of the reference's own speech sample
``tts_output.wav`` (16 kHz int16 mono; decoded as librosa.load does, int16 / 32768).  Real speech matters because its
spectrum reaches both clip rails of ``_normalize`` (silence -> -4), which white noise never does.
"""
from __future__ import annotations

import os

import numpy as np

from . import ref_shim, resample, synth, weights

GOLDEN = weights._GOLDEN
EXCERPT = (48000, 48000 + 64000)          # samples of tts_output.wav kept as the fixture (4 s: speech + pauses, 17 % of the bins on the -4 rail)


def main():
    ref_audio = ref_shim.load_audio()
    from scipy.io import wavfile
    sr, pcm = wavfile.read(os.path.join(ref_shim.REF_ROOT, "tts_output.wav"))
    assert sr == 16000 and pcm.dtype == np.int16 and pcm.ndim == 1
    pcm = np.ascontiguousarray(pcm[EXCERPT[0]:EXCERPT[1]])
    speech = resample.pcm_to_float_mono(pcm)
    m_speech = ref_audio.melspectrogram(speech)
    m_synth = ref_audio.melspectrogram(synth.wav(1.0, seed=0))
    assert m_speech.dtype == np.float64 and m_speech.shape == (80, 1 + len(speech) // 200)
    frac_lo, frac_hi = float((m_speech <= -4).mean()), float((m_speech >= 4).mean())
    np.savez_compressed(os.path.join(GOLDEN, "mel_ref_golden.npz"), speech_pcm=pcm, mel_speech=m_speech, mel_synth_seed0_1s=m_synth,
                        preemph_synth=ref_audio.preemphasis(synth.wav(1.0, seed=0), 0.97))
    print("speech mel", m_speech.shape, "clipped low %.3f high %.3f" % (frac_lo, frac_hi), "synth mel", m_synth.shape)


if __name__ == "__main__":
    main()
