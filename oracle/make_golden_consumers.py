"""Golden vectors for pyaudioanalysis_b200/consumers.py from the UNMODIFIED reference (dev container only).

audioTrainTest / audioSegmentation cannot be imported here (imblearn, hmmlearn, plotly are absent), so the two pure
functions under test are executed from their source text: `labels_to_segments` (audioSegmentation.py:58-99) and the
`Knn` class (audioTrainTest.py:33-49).  Writes tests/golden/consumers.npz.
"""
import os

import numpy as np
from scipy.spatial import distance

REF = "/root/reference/pyAudioAnalysis"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "consumers.npz")


def _exec_between(path, start, stop, env):
    src = open(path).read()
    a = src.index(start)
    b = src.index(stop, a)
    exec(src[a:b], env)
    return env


def main():
    seg_env = _exec_between(os.path.join(REF, "audioSegmentation.py"), "def labels_to_segments", "def segments_to_labels", {"np": np})
    knn_env = _exec_between(os.path.join(REF, "audioTrainTest.py"), "class Knn", "def classifier_wrapper", {"np": np, "distance": distance})
    rng = np.random.default_rng(20260923)
    out = {}
    # ---- labels_to_segments
    seqs, segs, classes = [], [], []
    for _ in range(60):
        lab = rng.integers(0, 3, size=int(rng.integers(1, 24)))
        s, c = seg_env["labels_to_segments"](list(lab), 0.5)
        seqs.append(lab)
        segs.append(np.asarray(s, dtype=np.float64).reshape(-1))
        classes.append(np.asarray(c, dtype=np.int64))
    out["seg_n"] = np.int64(len(seqs))
    for i in range(len(seqs)):
        out["seg_labels_%d" % i] = seqs[i]
        out["seg_out_%d" % i] = segs[i]
        out["seg_classes_%d" % i] = classes[i]
    # ---- kNN
    feats = rng.normal(size=(90, 12))
    labels = rng.integers(0, 3, size=90)
    feats += labels[:, None] * 0.8
    knn = knn_env["Knn"](feats, labels, 7)
    test = rng.normal(size=(40, 12)) + 0.8
    ids, P = [], []
    for v in test:
        i, p = knn.classify(v)
        ids.append(i)
        P.append(p)
    out.update(knn_features=feats, knn_labels=labels, knn_neighbors=np.int64(7), knn_test=test, knn_ids=np.asarray(ids), knn_P=np.asarray(P))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
