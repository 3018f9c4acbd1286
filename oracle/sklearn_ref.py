"""Thin driver around the REAL scikit-learn k-means -- TEST INFRASTRUCTURE ONLY.

``sklearn.cluster.KMeans`` is the implementation the reference executes at
``members/jasraj/land_use_classification/core.py:227-228`` (third-party dependency,
``pyproject.toml:24``; scikit-learn 1.9.0 in this image, both in the build container and
on the GPU box).  This module runs it in float64 with an explicit ``init`` so both sides
of a parity test start from identical centroids, exposes the single-iteration Cython
routine for shard-wise single-step checks, and is what ``bench.py`` times for the
``cpu_baseline`` / ``--impl reference`` numbers.  Never imported by the product path.
"""
from __future__ import annotations

import time
import warnings

import numpy as np


def available() -> bool:
    try:
        import sklearn.cluster  # noqa: F401

        return True
    except Exception:
        return False


def n_threads() -> int:
    from sklearn.utils._openmp_helpers import _openmp_effective_n_threads

    return int(_openmp_effective_n_threads())


def fit(X, init, max_iter=300, tol=1e-4):
    """KMeans(n_clusters=K, init=<f64 array>, n_init=1, algorithm="lloyd").fit(X) in f64."""
    from sklearn.cluster import KMeans

    X = np.ascontiguousarray(X, dtype=np.float64)
    init = np.ascontiguousarray(init, dtype=np.float64)
    km = KMeans(
        n_clusters=init.shape[0], init=init, n_init=1, max_iter=max_iter, tol=tol, algorithm="lloyd"
    )
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        km.fit(X)
        wall = time.perf_counter() - t0
    return {
        "labels": km.labels_,
        "centers": km.cluster_centers_,
        "inertia": float(km.inertia_),
        "n_iter": int(km.n_iter_),
        "wall_s": wall,
    }


def lloyd_step(X, centers_old, threads=None, update_centers=True):
    """One call of ``sklearn.cluster._k_means_lloyd.lloyd_iter_chunked_dense`` (f64).

    Signature at sklearn:cluster/_k_means_lloyd.pyx:23-32.  Returns
    (labels, centers_new, weight_in_clusters, center_shift).
    """
    from sklearn.cluster._k_means_lloyd import lloyd_iter_chunked_dense

    X = np.ascontiguousarray(X, dtype=np.float64)
    c_old = np.ascontiguousarray(centers_old, dtype=np.float64)
    k = c_old.shape[0]
    c_new = np.zeros_like(c_old)
    w = np.zeros(k, dtype=np.float64)
    labels = np.full(X.shape[0], -1, dtype=np.int32)
    shift = np.zeros(k, dtype=np.float64)
    ones = np.ones(X.shape[0], dtype=np.float64)
    lloyd_iter_chunked_dense(
        X, ones, c_old, c_new, w, labels, shift, threads or n_threads(), update_centers
    )
    return labels, c_new, w, shift


def kmeans_plusplus(X, n_clusters, seed):
    """sklearn.cluster.kmeans_plusplus with a RandomState(seed)."""
    from sklearn.cluster import kmeans_plusplus as kpp

    return kpp(np.ascontiguousarray(X, dtype=np.float64), n_clusters, random_state=np.random.RandomState(seed))
