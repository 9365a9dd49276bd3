"""CPU, dev container only: the oracle against the UNMODIFIED reference on randomised configurations.

Skipped where /root/reference does not exist (the GPU box); the committed golden vectors cover that case.
"""
import contextlib
import io

import numpy as np
import pytest

from oracle import st_oracle as O
from oracle.ref_import import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def REF():
    return load_reference()


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


CASES = []
_rng = np.random.default_rng(20260922)
for _ in range(14):
    fs = int(_rng.choice([8000, 16000, 22050, 32000, 44100, 48000]))
    w = int(_rng.integers(max(240, fs // 70), fs // 12))
    s = int(_rng.integers(max(1, w // 8), w + 1))
    n = int(_rng.integers(3 * w, 12 * w))
    CASES.append((fs, w, s, n, int(_rng.integers(0, 10 ** 6))))


@pytest.mark.parametrize("fs,w,s,n,seed", CASES)
def test_random_configurations(REF, fs, w, s, n, seed):
    S, M, A = REF
    rng = np.random.default_rng(seed)
    x = O.synth_clip(seed, n, fs)
    if seed % 3 == 0:                       # float-valued input with a DC offset
        x = x.astype(np.float64) * float(rng.uniform(0.01, 3.0)) + float(rng.uniform(-500, 500))
    try:
        ref, names = S.feature_extraction(x, fs, w, s, deltas=bool(seed % 2))
    except (ValueError, IndexError) as exc:
        with pytest.raises(type(exc)):
            O.feature_extraction(x, fs, w, s, deltas=bool(seed % 2))
        return
    got, gnames = O.feature_extraction(x, fs, w, s, deltas=bool(seed % 2))
    assert gnames == names
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-10)
    loop, _ = O.feature_extraction_loop(x[: 4 * w], fs, w, s, deltas=bool(seed % 2))
    np.testing.assert_allclose(loop, S.feature_extraction(x[: 4 * w], fs, w, s, deltas=bool(seed % 2))[0], rtol=1e-8, atol=1e-10)
    sp_ref = _quiet(S.spectrogram, x, fs, w, s)
    sp = O.spectrogram(x, fs, w, s)
    np.testing.assert_allclose(sp[0], sp_ref[0], rtol=1e-9, atol=1e-12)
    assert sp[1] == sp_ref[1] and sp[2] == sp_ref[2]
    try:
        ch_ref = S.chromagram(x, fs, w, s)
    except ValueError:
        with pytest.raises(ValueError):
            O.chromagram(x, fs, w, s)
    else:
        ch = O.chromagram(x, fs, w, s)
        np.testing.assert_allclose(ch[0], ch_ref[0], rtol=1e-9, atol=1e-12)
        assert ch[1] == ch_ref[1] and ch[2] == ch_ref[2]
    mw, ms = int(rng.integers(2, 9)) * s + w, int(rng.integers(1, 9)) * s
    mid_ref = M.mid_feature_extraction(x, fs, mw, ms, w, s)
    mid = O.mid_feature_extraction(x, fs, mw, ms, w, s)
    np.testing.assert_allclose(mid[0], mid_ref[0], rtol=1e-8, atol=1e-10)
    assert mid[2] == mid_ref[2]


def test_tables_match_reference(REF):
    S, M, A = REF
    for fs, K in [(16000, 400), (44100, 441), (8000, 200), (22050, 551), (48000, 1200)]:
        np.testing.assert_array_equal(O.mel_filterbank(fs, K), S.mfcc_filter_banks(fs, K)[0])
        semis, share = S.chroma_features_init(K, fs)
        os_, osh = O.chroma_tables(fs, K)
        np.testing.assert_array_equal(os_, semis)
        np.testing.assert_array_equal(osh, share)
        rng = np.random.default_rng(K)
        X = rng.random(K)
        ref = S.chroma_features(X, fs, K)[1][:, 0]
        np.testing.assert_allclose(O.chroma_operator(fs, K) @ (X ** 2) / (X ** 2).sum(), ref, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("fs,w,s,n", [(4000, 100, 50, 50), (4000, 100, 50, 1000), (8000, 160, 80, 100),
                                      (8000, 160, 80, 1000), (4000, 400, 200, 100), (16000, 800, 400, 799)])
def test_error_precedence(REF, fs, w, s, n):
    """mel bank IndexError (before the loop) > no frames ValueError > chroma ValueError (frame 0)."""
    S, M, A = REF
    x = O.synth_clip(1, n, fs)
    with pytest.raises((ValueError, IndexError)) as ref:
        S.feature_extraction(x, fs, w, s)
    with pytest.raises(ref.type) as got:
        O.feature_extraction(x, fs, w, s)
    assert ("need at least one array" in str(ref.value)) == ("need at least one array" in str(got.value))
    with pytest.raises(ref.type):
        O.feature_extraction_loop(x, fs, w, s)
