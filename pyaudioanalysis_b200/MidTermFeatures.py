"""Drop-in for ``pyAudioAnalysis.MidTermFeatures.mid_feature_extraction`` (MidTermFeatures.py:87-127)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check, get_plan
from . import ShortTermFeatures
from .ShortTermFeatures import _as_clip, _fs_int, _ptr
from .batch import mid_ratios


def mid_feature_extraction(signal, sampling_rate, mid_window, mid_step, short_window, short_step):
    """Mid-term feature extraction: (mid float64 [136 x M], short float64 [68 x T], 136 names).

    Short-term features with deltas (MidTermFeatures.py:93-95), then the mean and population
    standard deviation of every row over runs of ``ratio`` frames every ``step_ratio`` frames
    (:100-124), ``np.nan_to_num`` (:126).  All window arguments are in samples.
    """
    w, s = int(short_window), int(short_step)
    x, code = _as_clip(signal)
    plan = get_plan(_fs_int(sampling_rate), w, s)
    T = lib().b200aa_num_frames(x.shape[0], w, s)
    if T <= 0:
        check(_lib.ERR_TOO_SHORT)
    ratio, stepr = mid_ratios(mid_window, mid_step, short_window, short_step)
    if ratio < 1 or stepr < 1:
        raise ValueError("mid-term window / step shorter than one short-term step")
    M = lib().b200aa_mid_windows(T, stepr)
    mid = np.empty((136, M), dtype=np.float32)
    st = np.empty((68, T), dtype=np.float32)
    check(lib().b200aa_mid_features_host(plan.handle, _ptr(x), code, x.shape[0], ratio, stepr, _ptr(mid), _ptr(st)))
    st_names = ShortTermFeatures.feature_names(True)
    names = [n + "_mean" for n in st_names] + [n + "_std" for n in st_names]
    return mid.astype(np.float64), st.astype(np.float64), names
