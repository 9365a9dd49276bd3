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
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/make_ref.py (byte-for-byte copy)


def reference_available() -> bool:
    """The reference tree itself (dev container only)."""
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyAudioAnalysis"))


def staged_available() -> bool:
    """The staged copy of the hot-path modules (travels to the GPU box; bench.py's CPU baseline)."""
    return os.path.isfile(os.path.join(STAGED_ROOT, "pyAudioAnalysis", "ShortTermFeatures.py"))


def load_reference(staged_ok=False):
    """Returns (ShortTermFeatures, MidTermFeatures, audioBasicIO) modules of the reference.

    ``staged_ok``: fall back to (or, for bench.py, use) the copy under oracle/_ref when the tree is absent."""
    root = REFERENCE_ROOT
    if not reference_available():
        if not (staged_ok and staged_available()):
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        root = STAGED_ROOT
    for name in ("matplotlib", "matplotlib.pyplot", "eyed3", "pydub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["pydub"], "AudioSegment"):
        sys.modules["pydub"].AudioSegment = None
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from pyAudioAnalysis import ShortTermFeatures, MidTermFeatures, audioBasicIO
    return ShortTermFeatures, MidTermFeatures, audioBasicIO
