"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the dev container only).

    python -m oracle.make_golden

Inputs are stored next to the outputs so the fixtures are self-contained on
the GPU box, where /root/reference does not exist.
"""
import contextlib
import io
import os

import numpy as np

from oracle.ref_import import load_reference, REFERENCE_ROOT
from oracle.st_oracle import synth_clip

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def main():
    S, M, A = load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- config 1: doremi.wav 50/25 ms (BASELINE.json configs[0])
    fs, x = A.read_audio_file(os.path.join(REFERENCE_ROOT, "pyAudioAnalysis", "data", "doremi.wav"))
    F, names = S.feature_extraction(x, fs, 0.050 * fs, 0.025 * fs)
    sp, sp_t, sp_f = quiet(S.spectrogram, x, fs, 800, 400)
    ch, ch_t, ch_n = S.chromagram(x, fs, 800, 400)
    mid, st, mid_names = M.mid_feature_extraction(x, fs, fs, fs, 0.05 * fs, 0.025 * fs)
    np.savez_compressed(os.path.join(OUT, "doremi.npz"), fs=fs, x=x, st=F, names=np.array(names),
                        spectrogram=sp.astype(np.float32), spectrogram_sum=sp.sum(),
                        spec_time=np.array(sp_t), spec_freq=np.array(sp_f),
                        chromagram=ch, chroma_time=np.array(ch_t), chroma_names=np.array(ch_n),
                        mid=mid, mid_names=np.array(mid_names))

    # ---- the reference's own pytest inputs (shape pins, pytests/test_feature_extraction.py)
    fs1, x1 = A.read_audio_file(os.path.join(REFERENCE_ROOT, "pytests", "test_data", "1_sec_wav.wav"))
    F1, n1 = S.feature_extraction(x1, fs1, 0.050 * fs1, 0.050 * fs1)
    fs5, x5 = A.read_audio_file(os.path.join(REFERENCE_ROOT, "pytests", "test_data", "5_sec_wav.wav"))
    m5, s5, mn5 = M.mid_feature_extraction(x5, fs5, 1 * fs5, 1 * fs5, 0.05 * fs5, 0.05 * fs5)
    np.savez_compressed(os.path.join(OUT, "pytests.npz"), fs1=fs1, x1=x1, st1=F1,
                        fs5=fs5, x5=x5, mid5=m5, st5=s5, mid_names5=np.array(mn5))

    # ---- seeded synthetic clips (SURVEY 8d recipe): 16 kHz 50/25 and 44.1 kHz 20/10
    syn = {}
    for idx in (0, 1, 2):
        c = synth_clip(idx, 32000, 16000)
        f, _ = S.feature_extraction(c, 16000, 800, 400)
        syn[f"st16_{idx}"] = f
    c44 = synth_clip(7, 44100, 44100)
    f44, _ = S.feature_extraction(c44, 44100, 882, 441)
    sp44 = quiet(S.spectrogram, c44, 44100, 882, 441)[0]
    ch44 = S.chromagram(c44, 44100, 882, 441)[0]
    syn["st44"] = f44
    syn["sp44"] = sp44
    syn["ch44"] = ch44
    # a float-valued input, no deltas, odd window, non-50% hop
    cf = synth_clip(11, 20000, 22050).astype(np.float64) * 0.37 + 11.5
    ff, _ = S.feature_extraction(cf, 22050, 551, 200, deltas=False)
    syn["st_float_551"] = ff
    # 1 s windows as music_thumbnailing uses them (audioSegmentation.py:1137-1139)
    cl = synth_clip(13, 16000 * 5, 16000)
    fl, _ = S.feature_extraction(cl, 16000, 16000, 16000)
    syn["st_win16000"] = fl
    # mid-term with an awkward ratio and a short last window
    mm, ss, _ = M.mid_feature_extraction(synth_clip(3, 50000, 16000), 16000, 16000, 8000, 800, 400)
    syn["mid_16000_8000"] = mm
    np.savez_compressed(os.path.join(OUT, "synthetic.npz"), **syn)

    # ---- edge cases
    edge = {}
    z = np.zeros(4000, dtype=np.int16)
    edge["zeros"] = S.feature_extraction(z, 16000, 800, 400)[0]
    k = np.full(4000, 1234, dtype=np.int16)
    edge["const"] = S.feature_extraction(k, 16000, 800, 400)[0]
    for n in (800, 1199, 1200):
        edge[f"n{n}"] = S.feature_extraction(synth_clip(5, n, 16000), 16000, 800, 400)[0]
    # digital silence in the middle of a clip with a DC offset
    s = synth_clip(21, 12000, 16000).astype(np.int32) // 4 + 700
    s[3000:7000] = 0
    s = s.astype(np.int16)
    edge["silence_x"] = s
    edge["silence"] = S.feature_extraction(s, 16000, 800, 400)[0]
    # chromagram whose last frame is clipped at the end of the clip
    cc = synth_clip(22, 16300, 16000)
    edge["chroma_clipped"] = S.chromagram(cc, 16000, 800, 400)[0]
    edge["spec_16300"] = quiet(S.spectrogram, cc, 16000, 800, 400)[0]
    np.savez_compressed(os.path.join(OUT, "edges.npz"), **edge)
    # ---- directory_feature_extraction (long-term averaged mid-term vectors per file)
    # 12 files of each class of the reference's own pytests/test_data/3_class (8 kHz, 1 s clips)
    import glob
    import shutil
    import tempfile
    dirs = {}
    for cls in ("music", "silence", "speech"):
        files = sorted(glob.glob(os.path.join(REFERENCE_ROOT, "pytests", "test_data", "3_class", cls, "*.wav")))[:12]
        tmp = tempfile.mkdtemp()
        for f in files:
            shutil.copy(f, tmp)
        feats, flist, names = quiet(M.directory_feature_extraction, tmp, 1.0, 1.0, 0.05, 0.05, compute_beat=False)
        xs = np.stack([A.read_audio_file(f)[1] for f in sorted(glob.glob(os.path.join(tmp, "*.wav")))])
        dirs[cls + "_x"] = xs
        dirs[cls + "_files"] = np.array([os.path.basename(f) for f in flist])
        dirs[cls + "_feats"] = feats
        dirs["names"] = np.array(names)
        shutil.rmtree(tmp)
    dirs["fs"] = 8000
    np.savez_compressed(os.path.join(OUT, "dirs.npz"), **dirs)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
