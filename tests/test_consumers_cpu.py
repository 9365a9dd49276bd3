"""CPU: the host side of pyaudioanalysis_b200/consumers.py against values of the unmodified reference
(tests/golden/consumers.npz, written by oracle/make_golden_consumers.py): labels_to_segments, the library's kNN over a
matrix of test vectors, and batch == per-row for the scikit-learn branch of classifier_wrapper."""
import os
import types

import numpy as np

from pyaudioanalysis_b200 import consumers as C

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "consumers.npz"))


def test_labels_to_segments_golden():
    for i in range(int(G["seg_n"])):
        seg, classes = C.labels_to_segments(list(G["seg_labels_%d" % i]), 0.5)
        assert np.array_equal(np.asarray(seg, dtype=np.float64).reshape(-1), G["seg_out_%d" % i]), i
        assert [int(c) for c in classes] == [int(c) for c in G["seg_classes_%d" % i]], i


def test_knn_matrix_matches_reference_loop():
    knn = types.SimpleNamespace(features=G["knn_features"], labels=G["knn_labels"], neighbors=int(G["knn_neighbors"]))
    ids, P = C.classify_vectors(knn, "knn", G["knn_test"])
    assert np.array_equal(ids, G["knn_ids"])
    assert np.allclose(P, G["knn_P"], rtol=0, atol=1e-15)


def test_sklearn_batch_equals_per_row():
    from sklearn.svm import SVC
    rng = np.random.default_rng(3)
    X = rng.normal(size=(120, 10))
    y = (X[:, 0] + 0.5 * X[:, 1] > 0).astype(int)
    svm = SVC(C=1.0, kernel="linear", probability=True, random_state=0).fit(X, y)
    T = rng.normal(size=(25, 10))
    ids, P = C.classify_vectors(svm, "svm", T)
    for i in range(T.shape[0]):                       # classifier_wrapper, audioTrainTest.py:91-92
        assert ids[i] == svm.predict(T[i].reshape(1, -1))[0]
        assert np.allclose(P[i], svm.predict_proba(T[i].reshape(1, -1))[0], rtol=1e-12, atol=1e-15)
    ids, P = C.classify_vectors(object(), "no such type", T)
    assert (ids == -1).all() and (P == -1).all()
