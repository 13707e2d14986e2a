"""Import shim for the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only).

/root/reference exists only in the build container.  Two non-arithmetic imports
of the reference are absent there and are stubbed: ``torchsummary``
(models/__init__.py:6) and ``basicsr.archs.arch_util.default_init_weights``
(models/base_blocks.py:9, used only by ENet's ModulatedConv2d).  ``librosa`` is
absent too, so ``futils/audio.py`` cannot be imported (see oracle/mel.py).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("S2V_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models"))


def load():
    """Returns (LNet, DNet, flow_util) classes/modules of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    for name in ("torchsummary", "basicsr", "basicsr.archs", "basicsr.archs.arch_util"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["torchsummary"].summary = lambda *a, **k: None
    sys.modules["basicsr.archs.arch_util"].default_init_weights = lambda *a, **k: None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # our own drop-in mirror also has top-level names `models` / `futils` inside its
    # package, never at top level, so these resolve to the reference.
    from models.LNet import LNet          # noqa
    from models.DNet import DNet          # noqa
    from futils import flow_util          # noqa
    return LNet, DNet, flow_util
