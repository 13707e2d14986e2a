"""Importable alias of the ``speech-to-video-mpp_b200/`` package directory.

The project directory name contains hyphens (it mirrors the reference repo's name),
which Python cannot import; this shim points the package path at it, so
``import s2v_b200.models.LNet`` resolves to ``speech-to-video-mpp_b200/models/LNet.py``.
"""
import os as _os

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speech-to-video-mpp_b200")
__path__.insert(0, _pkg_dir)

from ._lib import lib_path, load_library  # noqa: E402,F401
