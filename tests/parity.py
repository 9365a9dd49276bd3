"""Shared comparison helpers for the GPU parity tests.

Tolerance (stated once, used everywhere): every feature column must satisfy
    |gpu - ref| <= RTOL * |ref| + ATOL      with RTOL = 1e-4, ATOL = 1e-5
(BASELINE.json north_star asks for 1e-4 rtol; the absolute term covers values that cross zero --
mfcc_2..13 and every delta column are differences of near-equal numbers -- see SURVEY.md 8d.)
Two rows are discrete: zcr moves in quanta of 0.5/(w-1) and spectral_rolloff in quanta of 1/K; a
float32-vs-float64 tie may move them by one quantum on a small fraction of frames, which is
counted and bounded separately.
"""
import numpy as np

RTOL, ATOL = 1e-4, 1e-5
ROLLOFF_ROW = 7
MAX_FLIP_FRACTION = 2e-3


def check_features(gpu, ref, K, what=""):
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert gpu.shape == ref.shape, (what, gpu.shape, ref.shape)
    assert np.isfinite(gpu).all(), what + ": non-finite output"
    err = np.abs(gpu - ref)
    tol = RTOL * np.abs(ref) + ATOL
    bad = err > tol
    F = ref.shape[0]
    flips = 0
    for r in (ROLLOFF_ROW, ROLLOFF_ROW + 34):
        if r < F and bad[r].any():
            q = err[r][bad[r]]
            assert (q <= (1.0 / K) * (2 if r >= 34 else 1) + 1e-6).all(), "%s: rolloff off by more than one quantum" % what
            flips += int(bad[r].sum())
            bad[r] = False
    assert flips <= max(2, MAX_FLIP_FRACTION * ref.shape[1] * 2), "%s: %d rolloff quantum flips" % (what, flips)
    if bad.any():
        rows = np.unique(np.nonzero(bad)[0])
        worst = [(int(r), float(err[r].max()), float((err[r] / tol[r]).max())) for r in rows]
        raise AssertionError("%s: rows outside tolerance (row, max abs err, max err/tol): %s" % (what, worst))
    return flips


def check_close(gpu, ref, what="", rtol=RTOL, atol=ATOL):
    gpu = np.asarray(gpu, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert gpu.shape == ref.shape, (what, gpu.shape, ref.shape)
    np.testing.assert_allclose(gpu, ref, rtol=rtol, atol=atol, err_msg=what)
