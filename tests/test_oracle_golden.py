"""CPU: pin the oracle (oracle/st_oracle.py) against golden vectors made by the unmodified reference."""
import ctypes

import numpy as np
import pytest

from oracle import st_oracle as O
from oracle.build_c import build as build_c

RT, AT = 1e-9, 1e-11     # float64 restatement vs float64 reference (different summation order only)


def close(a, b, rt=RT, at=AT):
    np.testing.assert_allclose(a, b, rtol=rt, atol=at)


def test_names():
    n = O.feature_names(True)
    assert len(n) == 68 and n[0] == "zcr" and n[8] == "mfcc_1" and n[33] == "chroma_std"
    assert n[34] == "delta zcr" and n[67] == "delta chroma_std"
    assert len(O.feature_names(False)) == 34


def test_doremi_short_term(golden_doremi):
    g = golden_doremi
    F, names = O.feature_extraction(g["x"], int(g["fs"]), 0.050 * 16000, 0.025 * 16000)
    assert F.shape == (68, 319)
    assert names == list(g["names"])
    close(F[:, 5:], g["st"][:, 5:])
    # frames 0-3 are digital silence: non-DC bins are float64 round-off in the reference, so a few
    # features there are defined by that noise (spread ~4e-9); absolute agreement only.
    np.testing.assert_allclose(F[:, :5], g["st"][:, :5], rtol=0, atol=1e-7)
    # anchors recorded in SURVEY.md 8c
    assert abs(g["st"].sum() - (-7234.680904945822)) < 1e-6
    assert abs(np.abs(g["st"]).sum() - 13662.223468905411) < 1e-6


def test_doremi_loop_flavour(golden_doremi):
    g = golden_doremi
    x = g["x"][:20000]
    Fl, _ = O.feature_extraction_loop(x, 16000, 800, 400)
    Fv, _ = O.feature_extraction(x, 16000, 800, 400)
    close(Fl, Fv)
    Fn, _ = O.feature_extraction_loop(x, 16000, 800, 400, tables_per_frame=False)
    close(Fl, Fn, 0, 0)


def test_doremi_spectrogram_chromagram_mid(golden_doremi):
    g = golden_doremi
    sp, t, f = O.spectrogram(g["x"], 16000, 800, 400)
    assert sp.shape == (319, 400) and not sp[317:].any()
    np.testing.assert_allclose(sp, g["spectrogram"], rtol=2e-6, atol=1e-9)   # fixture stored as f32
    assert abs(sp.sum() - float(g["spectrogram_sum"])) < 1e-9
    close(np.array(t), g["spec_time"]); close(np.array(f), g["spec_freq"])
    ch, t, n = O.chromagram(g["x"], 16000, 800, 400)
    assert ch.shape == (318, 12) and n == list(g["chroma_names"])
    close(ch, g["chromagram"]); close(np.array(t), g["chroma_time"])
    mid, st, mn = O.mid_feature_extraction(g["x"], 16000, 16000, 16000, 800, 400)
    assert mid.shape == (136, 8) and mn == list(g["mid_names"])
    close(mid, g["mid"], 1e-6, 1e-9)      # window 0 pools the noise-defined silent frames


def test_reference_pytest_inputs(golden_pytests):
    g = golden_pytests
    F, n = O.feature_extraction(g["x1"], int(g["fs1"]), 0.05 * 16000, 0.05 * 16000)
    assert F.shape[1] == 20 and F.shape[0] == len(n)          # pytests/test_feature_extraction.py:14-15
    close(F, g["st1"])
    mid, st, mn = O.mid_feature_extraction(g["x5"], 16000, 16000, 16000, 800, 800)
    assert mid.shape[1] == 5 and mid.shape[0] == len(mn) == 136    # :27-28
    close(mid, g["mid5"]); close(st, g["st5"])
    assert mn == list(g["mid_names5"])


def test_synthetic(golden_synth):
    g = golden_synth
    for idx in (0, 1, 2):
        close(O.feature_extraction(O.synth_clip(idx, 32000, 16000), 16000, 800, 400)[0], g[f"st16_{idx}"])
    c44 = O.synth_clip(7, 44100, 44100)
    close(O.feature_extraction(c44, 44100, 882, 441)[0], g["st44"])
    close(O.spectrogram(c44, 44100, 882, 441)[0], g["sp44"])
    close(O.chromagram(c44, 44100, 882, 441)[0], g["ch44"])
    cf = O.synth_clip(11, 20000, 22050).astype(np.float64) * 0.37 + 11.5
    close(O.feature_extraction(cf, 22050, 551, 200, deltas=False)[0], g["st_float_551"])
    close(O.feature_extraction(O.synth_clip(13, 80000, 16000), 16000, 16000, 16000)[0], g["st_win16000"])
    close(O.mid_feature_extraction(O.synth_clip(3, 50000, 16000), 16000, 16000, 8000, 800, 400)[0],
          g["mid_16000_8000"])


def test_edges(golden_edges):
    g = golden_edges
    z = np.zeros(4000, dtype=np.int16)
    Fz = O.feature_extraction(z, 16000, 800, 400)[0]
    close(Fz, g["zeros"])
    assert abs(Fz[8, 0] - (-99.00180475419432)) < 1e-9
    # constant / digitally silent frames: the reference's non-DC bins are float64 round-off, so
    # log-of-noise features (mfcc) are defined by that noise; compare those loosely.
    Fk = O.feature_extraction(np.full(4000, 1234, dtype=np.int16), 16000, 800, 400)[0]
    np.testing.assert_allclose(Fk[:8], g["const"][:8], rtol=1e-6, atol=1e-6)
    for n in (800, 1199, 1200):
        F = O.feature_extraction(O.synth_clip(5, n, 16000), 16000, 800, 400)[0]
        assert F.shape == g[f"n{n}"].shape
        close(F, g[f"n{n}"])
    with pytest.raises(ValueError):
        O.feature_extraction(O.synth_clip(5, 799, 16000), 16000, 800, 400)
    Fs = O.feature_extraction(g["silence_x"], 16000, 800, 400)[0]
    ok = np.ones(Fs.shape[1], bool); ok[7:18] = False          # frames inside / touching the silent span
    close(Fs[:, ok], g["silence"][:, ok])
    cc = O.synth_clip(22, 16300, 16000)
    close(O.chromagram(cc, 16000, 800, 400)[0], g["chroma_clipped"])
    close(O.spectrogram(cc, 16000, 800, 400)[0], g["spec_16300"])
    with pytest.raises(ValueError):
        O.chroma_operator(8000, 80)       # reference's chroma else-branch cannot succeed


def test_mid_ratios():
    assert O.mid_ratios(16000, 16000, 800, 400) == (39, 40)
    assert O.mid_ratios(16000, 16000, 800, 800) == (20, 20)


def test_transform_definitions():
    """numpy.fft / the DCT matrix agree with the naive C definitions (SciPy's published formulas)."""
    lib = ctypes.CDLL(build_c())
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(3)
    for n in (800, 882, 551, 97):
        x = rng.standard_normal(n)
        re = np.zeros(n); im = np.zeros(n)
        lib.oracle_dft_real(x.ctypes.data_as(dp), ctypes.c_size_t(n), re.ctypes.data_as(dp), im.ctypes.data_as(dp))
        Z = np.fft.fft(x)
        np.testing.assert_allclose(Z.real, re, atol=1e-10); np.testing.assert_allclose(Z.imag, im, atol=1e-10)
    m = rng.standard_normal(40)
    y = np.zeros(13)
    lib.oracle_dct2_ortho(m.ctypes.data_as(dp), ctypes.c_size_t(40), y.ctypes.data_as(dp), ctypes.c_size_t(13))
    np.testing.assert_allclose(O.dct_matrix() @ m, y, atol=1e-13)
    from scipy.fft import dct
    np.testing.assert_allclose(dct(m, type=2, norm="ortho")[:13], y, atol=1e-13)
