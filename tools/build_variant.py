"""Development: builds a variant of libs2v.so with extra -D flags into its own object directory.

    python tools/build_variant.py prof -DS2V_EPI_PROF        -> speech-to-video-mpp_b200/libs2v_prof.so

Select it with S2V_LIB=<path>.  The product build (speech-to-video-mpp_b200/build.py) is untouched."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("s2v_build", os.path.join(ROOT, "speech-to-video-mpp_b200", "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
tag, defs = sys.argv[1], sys.argv[2:]
b.OBJ = os.path.join(b.HERE, "build_" + tag)
b.LIB = os.path.join(b.HERE, "libs2v_%s.so" % tag)
b.FLAGS = b.FLAGS + defs
print(b.build())
