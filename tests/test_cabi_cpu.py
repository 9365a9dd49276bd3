"""CPU: the C-ABI library builds, loads and exports every symbol include/b200aa.h declares; host-only
entry points (tables, frame counts) agree with the oracle.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import st_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from pyaudioanalysis_b200.build import build
    build()
    from pyaudioanalysis_b200 import _lib
    return _lib


def test_header_symbols_exported(L):
    header = open(os.path.join(ROOT, "include", "b200aa.h")).read()
    declared = set(re.findall(r"\b(b200aa_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(L.SIGNATURES), (declared ^ set(L.SIGNATURES))
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200aa_abi_version() == 1
    assert b"concatenate" in lib.b200aa_status_string(-2)


@pytest.mark.parametrize("fs,w", [(16000, 800), (44100, 882), (22050, 551), (8000, 400), (48000, 2400), (16000, 16000)])
def test_host_tables_match_oracle(L, fs, w):
    K = w // 2
    np.testing.assert_allclose(L.host_table(fs, w, "mel"), O.mel_filterbank(fs, K), rtol=0, atol=1e-15)
    np.testing.assert_allclose(L.host_table(fs, w, "chroma"), O.chroma_operator(fs, K), rtol=0, atol=1e-15)
    np.testing.assert_allclose(L.host_table(fs, w, "dct"), O.dct_matrix(), rtol=0, atol=1e-15)


def test_table_sparsity_matches_survey(L):
    mel = L.host_table(16000, 800, "mel")
    assert np.count_nonzero(mel) == 323 and np.nonzero(mel.any(axis=0))[0].max() == 171
    assert np.count_nonzero(L.host_table(16000, 800, "chroma")) == 71                         # row 14
    assert np.count_nonzero(L.host_table(44100, 882, "chroma")) == 74


def test_reference_error_cases(L):
    with pytest.raises(ValueError):
        L.host_table(8000, 160, "chroma")          # reference: chroma else-branch raises
    with pytest.raises(IndexError):
        L.host_table(4000, 400, "mel")             # reference: fancy store beyond num_fft raises IndexError
    with pytest.raises(IndexError):
        O.mel_filterbank(4000, 200)


def test_counts(L):
    lib = L.lib()
    for n, w, s in [(128164, 800, 400), (16000, 800, 800), (799, 800, 400), (800, 800, 400), (1199, 800, 400),
                    (1200, 800, 400), (2646000, 882, 441), (16300, 800, 400), (16000, 800, 300), (16000, 800, 200)]:
        assert lib.b200aa_num_frames(n, w, s) == O.frame_count(n, w, s)
        x = np.zeros(n, dtype=np.int16)
        if n > 2 * w:
            assert lib.b200aa_spectrogram_rows(n, w, s) == O.spectrogram(x, 16000, w, s)[0].shape[0]
            assert lib.b200aa_chromagram_rows(n, w, s) == int((n - s - w) / s) + 1
    assert lib.b200aa_mid_windows(319, 40) == 8 and lib.b200aa_mid_windows(143999, 40) == 3600


def test_python_mirror_names():
    from pyaudioanalysis_b200 import ShortTermFeatures as S
    assert S.feature_names(True) == O.feature_names(True)
    assert S.feature_names(False) == O.feature_names(False)
    from pyaudioanalysis_b200.batch import mid_ratios
    assert mid_ratios(16000, 16000, 800, 400) == O.mid_ratios(16000, 16000, 800, 400) == (39, 40)
    assert mid_ratios(1.0 * 16000, 0.1 * 16000, 0.05 * 16000, 0.05 * 16000) == O.mid_ratios(16000.0, 1600.0, 800.0, 800.0)


def test_no_cpu_fallback():
    """The product path must refuse CPU tensors instead of computing on the host."""
    import torch
    import pyaudioanalysis_b200 as pkg
    with pytest.raises(TypeError):
        pkg.feature_extraction_batch(torch.zeros(2, 16000, dtype=torch.int16), 16000, 800, 400)
    src = "".join(open(os.path.join(ROOT, "pyaudioanalysis_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "pyaudioanalysis_b200")) if f.endswith(".py"))
    assert "oracle" not in src, "product code must never import the oracle"
