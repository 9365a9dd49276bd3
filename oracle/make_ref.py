"""Stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference is pure Python, so there is nothing to compile: this recipe copies the four modules the path needs
(ShortTermFeatures, MidTermFeatures and their two imports audioBasicIO / utilities, plus the package __init__) byte
for byte from /root/reference into ``oracle/_ref/pyAudioAnalysis/``.  ``oracle/_ref/`` is git-ignored (no reference
source enters the history) but travels to the GPU box with the snapshot, where ``bench.py`` times it as the CPU
baseline (``cpu_baseline.kind = "reference"``) through ``oracle/ref_import.py`` (which registers empty stand-ins for
matplotlib / eyed3 / pydub -- only plotting and mp3 code would touch them).  Run by ``__graft_entry__.build()``
wherever /root/reference exists; a sha256 manifest of the copied files is written beside them.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("PYAA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "ShortTermFeatures.py", "MidTermFeatures.py", "audioBasicIO.py", "utilities.py"]


def make(force=False):
    """Returns the staged root (oracle/_ref) or None when the reference tree is not present."""
    src_pkg = os.path.join(REF_SRC, "pyAudioAnalysis")
    if not os.path.isdir(src_pkg):
        return DST if os.path.isdir(os.path.join(DST, "pyAudioAnalysis")) else None
    dst_pkg = os.path.join(DST, "pyAudioAnalysis")
    os.makedirs(dst_pkg, exist_ok=True)
    manifest = {}
    for name in FILES:
        s, d = os.path.join(src_pkg, name), os.path.join(dst_pkg, name)
        if force or not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
        with open(s, "rb") as f:
            assert manifest[name] == hashlib.sha256(f.read()).hexdigest(), "staged copy differs from the reference: " + name
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": manifest}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(make(force=True))
