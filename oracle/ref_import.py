"""Import the UNMODIFIED reference (only possible where /root/reference exists).

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (golden-vector
generation) and by CPU tests that are skipped when the reference tree is
absent (it does not exist on the GPU box).  matplotlib / eyed3 / pydub are not
installed here, so empty stand-in modules are registered first; only plotting
and mp3 paths would touch them.
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("PYAA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyAudioAnalysis"))


def load_reference():
    """Returns (ShortTermFeatures, MidTermFeatures, audioBasicIO) modules of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "eyed3", "pydub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["pydub"], "AudioSegment"):
        sys.modules["pydub"].AudioSegment = None
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pyAudioAnalysis import ShortTermFeatures, MidTermFeatures, audioBasicIO
    return ShortTermFeatures, MidTermFeatures, audioBasicIO
