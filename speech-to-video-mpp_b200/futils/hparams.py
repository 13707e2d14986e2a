"""Front-end constants, same names and values as the reference's futils/hparams.py:20-84
(only the keys the mel / windowing path reads).  The CUDA mel kernel is specialised for
n_fft = win_size = 800, hop_size = 200; other values raise at call time."""


class HParams:
    def __init__(self, **kwargs):
        self.data = dict(kwargs)

    def __getattr__(self, key):
        data = self.__dict__.get("data", {})
        if key not in data:
            raise AttributeError("'HParams' object has no attribute %s" % key)
        return data[key]

    def set_hparam(self, key, value):
        self.data[key] = value


hparams = HParams(
    num_mels=80, rescale=True, rescaling_max=0.9, use_lws=False,
    n_fft=800, hop_size=200, win_size=800, sample_rate=16000, frame_shift_ms=None,
    signal_normalization=True, allow_clipping_in_normalization=True, symmetric_mels=True, max_abs_value=4.,
    preemphasize=True, preemphasis=0.97,
    min_level_db=-100, ref_level_db=20, fmin=55, fmax=7600,
    img_size=96, fps=25,
)
