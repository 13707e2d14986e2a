"""speech-to-video-mpp_b200: B200-native kernels behind VideoReTalking's per-frame lip-sync path.

Drop-in mirror of the reference's call surface for this path only:
  futils.audio.melspectrogram / mel_windows      (futils/audio.py:45, inference.py:209-216)
  futils.flow_util.*                             (futils/flow_util.py)
  models.LNet.LNet, models.DNet.DNet             (models/LNet.py:80, models/DNet.py:12)
Import as ``s2v_b200`` (the directory name has hyphens).
"""
