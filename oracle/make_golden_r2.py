"""Round-2 golden vectors from the UNMODIFIED reference (run in the dev container only):

    python -m oracle.make_golden_r2

* bigwin.npz -- 1 s windows at CD rates, as music_thumbnailing calls the path (audioSegmentation.py:1137-1139):
  feature_extraction(x, fs, fs, fs) for fs = 44100 and 22050, plus a window of 30 000 samples at step 15 000.
* beat.npz -- MidTermFeatures.beat_extraction (MidTermFeatures.py:18-84) on the short-term features of seeded clips with a
  pulse train.  The reference's peakdet uses numpy.Inf / numpy.NaN, removed in NumPy 2: they are provided as aliases
  of numpy.inf / numpy.nan for this run (the only patch, documented here; values are unchanged).
The existing fixtures (oracle/make_golden.py) are left untouched.
"""
import os

import numpy as np

from oracle.ref_import import load_reference
from oracle.st_oracle import synth_clip

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def pulse_clip(seed, n, fs, bpm):
    """Noise bursts at a steady tempo over a quiet tone (something with a detectable beat)."""
    rng = np.random.default_rng(seed)
    x = 300.0 * rng.standard_normal(n) + 1500.0 * np.sin(2 * np.pi * 220.0 * np.arange(n) / fs)
    period = int(round(fs * 60.0 / bpm))
    for s in range(0, n, period):
        e = min(n, s + fs // 20)
        x[s:e] += 9000.0 * rng.standard_normal(e - s) * np.linspace(1.0, 0.0, e - s)
    return np.round(np.clip(x, -32768, 32767)).astype(np.int16)


def main():
    S, M, A = load_reference()
    big = {}
    for fs, n in ((44100, 44100 * 4 + 1234), (22050, 22050 * 5)):
        x = synth_clip(31 + fs % 7, n, fs)
        big["x_%d" % fs] = x
        big["st_%d" % fs] = S.feature_extraction(x, fs, fs, fs)[0]
    x = synth_clip(40, 100000, 32000)
    big["x_30000"] = x
    big["st_30000"] = S.feature_extraction(x, 32000, 30000, 15000)[0]
    np.savez_compressed(os.path.join(OUT, "bigwin.npz"), **big)

    if not hasattr(np, "Inf"):
        np.Inf, np.NaN = np.inf, np.nan            # NumPy >= 2 dropped the aliases peakdet uses (utilities.py:62-63)
    beat = {}
    for i, (bpm, win) in enumerate(((120, 0.05), (90, 0.05), (140, 0.025), (75, 0.1))):
        fs = 16000
        x = pulse_clip(50 + i, fs * 12, fs, bpm)
        st, _ = S.feature_extraction(x, fs, int(win * fs), int(win * fs))
        b, r = M.beat_extraction(st, win)
        beat["st_%d" % i] = st
        beat["win_%d" % i] = win
        beat["bpm_%d" % i] = b
        beat["ratio_%d" % i] = r
    # a few synthetic feature matrices (random walks): exercises peakdet on rough data
    rng = np.random.default_rng(9)
    for i in range(4, 8):
        st = np.cumsum(rng.standard_normal((68, 300 + 40 * i)), axis=1) * 0.05 + rng.standard_normal((68, 1))
        b, r = M.beat_extraction(st, 0.05)
        beat["st_%d" % i] = st
        beat["win_%d" % i] = 0.05
        beat["bpm_%d" % i] = b
        beat["ratio_%d" % i] = r
    beat["n"] = 8
    np.savez_compressed(os.path.join(OUT, "beat.npz"), **beat)
    for f in ("bigwin.npz", "beat.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
