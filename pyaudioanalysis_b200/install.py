"""Rebind the reference's module attributes to the GPU path.

Every caller in the reference looks the four functions up on the module object at call time
(MidTermFeatures.py:93-95; audioSegmentation.py:371,423,482,573,707,842,1137; audioTrainTest.py:
1081,1128; audioAnalysis.py:46,71,80), so setattr on the imported modules is the drop-in.
"""
_saved = {}


def install(pyaudioanalysis_pkg=None):
    """Patch an importable ``pyAudioAnalysis``; returns the list of attributes rebound."""
    from . import ShortTermFeatures as ours_st, MidTermFeatures as ours_mt
    if pyaudioanalysis_pkg is None:
        import importlib
        ref_st = importlib.import_module("pyAudioAnalysis.ShortTermFeatures")
        ref_mt = importlib.import_module("pyAudioAnalysis.MidTermFeatures")
    else:
        ref_st, ref_mt = pyaudioanalysis_pkg.ShortTermFeatures, pyaudioanalysis_pkg.MidTermFeatures
    done = []
    for mod, ours, names in ((ref_st, ours_st, ("feature_extraction", "spectrogram", "chromagram")),
                             (ref_mt, ours_mt, ("mid_feature_extraction", "beat_extraction", "directory_feature_extraction",
                                                "multiple_directory_feature_extraction", "directory_feature_extraction_no_avg",
                                                "mid_feature_extraction_to_file", "mid_feature_extraction_file_dir"))):
        for n in names:
            _saved.setdefault((mod, n), getattr(mod, n))
            setattr(mod, n, getattr(ours, n))
            done.append("%s.%s" % (mod.__name__, n))
    return done


def uninstall():
    for (mod, n), fn in list(_saved.items()):
        setattr(mod, n, fn)
    _saved.clear()
