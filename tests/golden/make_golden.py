"""Generates the golden fixtures in this directory by running the REAL reference arithmetic.

The reference's k-means is ``sklearn.cluster.KMeans`` (third-party; called at
``members/jasraj/land_use_classification/core.py:227-228``); its unprojection tail is plain
numpy inside ``members/rafael/disparity/plugin.py:147-192`` (not importable here, restated
in ``oracle/unproject_oracle.py``).  This script was run in the build container with
scikit-learn 1.9.0 / numpy 2.3.5:

    python tests/golden/make_golden.py

Outputs (committed): kmeans_stack_small.npz, kmeans_c1_like.npz, kmeans_tol.npz,
kmeans_relocate.npz, kmeanspp.npz, unproject_small.npz, kmeans_default_call.npz
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import sklearn_ref, unproject_oracle  # noqa: E402

synth = importlib.import_module("3d-point-cloud-multiday-imagery_b200.synth")


def stack_case(name, D, H, W, k, max_iter, tol, seed, n_buildings=12):
    hm = synth.make_stack(D, H, W, seed=seed, n_buildings=n_buildings).numpy()
    P = unproject_oracle.unproject_stack(hm)  # float64 x,y,z
    P32 = P.astype(np.float32)
    assert np.array_equal(P32.astype(np.float64), P)
    init = synth.init_from_points(P32, k, seed)
    r = sklearn_ref.fit(P, init, max_iter=max_iter, tol=tol)
    mean = P.mean(axis=0)
    labels1, c_new1, w1, shift1 = sklearn_ref.lloyd_step(P - mean, init - mean)
    np.savez_compressed(
        os.path.join(HERE, name),
        height_maps=hm, init=init, k=k, max_iter=max_iter, tol=tol,
        n_points=P.shape[0], labels=r["labels"].astype(np.int32), centers=r["centers"],
        inertia=r["inertia"], n_iter=r["n_iter"],
        step_labels=labels1.astype(np.int32), step_centers=c_new1 + mean, step_counts=w1,
    )
    print(name, "N", P.shape[0], "n_iter", r["n_iter"], "inertia", r["inertia"])


def relocate_case():
    # 3-D version of sklearn/cluster/tests/test_k_means.py:85-111 plus a larger random one
    rs = np.random.RandomState(3)
    X = np.concatenate([rs.normal(0, 1, (300, 3)), rs.normal(8, 1, (300, 3))]).astype(np.float32)
    init = np.array([[0, 0, 0], [0.5, 0.5, 0.5], [100, 100, 100], [-90, 50, 3]], dtype=np.float64)
    r = sklearn_ref.fit(X.astype(np.float64), init, max_iter=100, tol=0.0)
    np.savez_compressed(os.path.join(HERE, "kmeans_relocate.npz"), X=X, init=init,
                        labels=r["labels"].astype(np.int32), centers=r["centers"],
                        inertia=r["inertia"], n_iter=r["n_iter"])
    print("relocate n_iter", r["n_iter"], "inertia", r["inertia"], np.bincount(r["labels"]))


def kpp_case():
    rs = np.random.RandomState(11)
    X = np.concatenate([rs.normal(c, 1.5, (400, 3)) for c in (0, 10, -7, 25)]).astype(np.float32)
    out = {}
    for seed, k in ((0, 4), (5, 8), (42, 16)):
        c, idx = sklearn_ref.kmeans_plusplus(X.astype(np.float64), k, seed)
        out[f"centers_{seed}_{k}"] = c
        out[f"indices_{seed}_{k}"] = idx
    np.savez_compressed(os.path.join(HERE, "kmeanspp.npz"), X=X, **out)
    print("kmeans++ cases", sorted(out))


def unproject_case():
    rs = np.random.RandomState(1)
    D, H, W = 3, 20, 28
    disp = rs.randint(-2000, 2000, size=(D, H, W)).astype(np.int16)
    disp[rs.rand(D, H, W) < 0.1] = 32767  # OpenCV-style sentinel -> |h| > 144
    mask = rs.rand(D, H, W) > 0.15
    hm = unproject_oracle.height_from_disparity(disp)
    P = unproject_oracle.unproject_stack(hm, mask, detrend=False)
    Pd = unproject_oracle.unproject_stack(hm, mask, detrend=True)
    z0, hn = unproject_oracle.ground_level(Pd[:, 2])
    Pt, hnt, lo, hi, off = unproject_oracle.reference_tail_stack(hm, mask, detrend=True)
    np.savez_compressed(os.path.join(HERE, "unproject_small.npz"), disparity=disp, mask=mask,
                        points=P, points_detrended=Pd, z_ground=z0, height_norm=hn,
                        tail_points=Pt, tail_height_norm=hnt, tail_h_min=lo, tail_h_max=hi, tail_offsets=off)
    print("unproject N", P.shape[0])


def default_call_case():
    """The reference's very call, ``KMeans(n_clusters, random_state=42, n_init=10)``
    (core.py:227-228; default init="k-means++", max_iter=300, tol=1e-4), on a small cloud."""
    from sklearn.cluster import KMeans

    hm = synth.make_stack(2, 40, 56, seed=7, n_buildings=6).numpy()
    P = unproject_oracle.unproject_stack(hm)
    out = {"height_maps": hm}
    for k, n_init in ((5, 10), (3, 1)):
        km = KMeans(n_clusters=k, random_state=42, n_init=n_init).fit(P)
        out[f"labels_{k}"] = km.labels_.astype(np.int32)
        out[f"centers_{k}"] = km.cluster_centers_
        out[f"inertia_{k}"] = km.inertia_
        out[f"n_iter_{k}"] = km.n_iter_
        km = KMeans(n_clusters=k, random_state=42, n_init=n_init, init="random").fit(P)
        out[f"rlabels_{k}"] = km.labels_.astype(np.int32)
        out[f"rcenters_{k}"] = km.cluster_centers_
        out[f"rinertia_{k}"] = km.inertia_
        out[f"rn_iter_{k}"] = km.n_iter_
    np.savez_compressed(os.path.join(HERE, "kmeans_default_call.npz"), **out)
    print("default call", {k: v for k, v in out.items() if k.startswith(("n_iter", "inertia", "rn_iter"))})


if __name__ == "__main__":
    only = set(sys.argv[1:])
    if not only or "stack" in only:
        stack_case("kmeans_stack_small.npz", 2, 48, 64, 5, 50, 1e-4, 0)
        stack_case("kmeans_c1_like.npz", 3, 96, 128, 8, 20, 0.0, 1)
        stack_case("kmeans_tol.npz", 2, 64, 64, 6, 300, 1e-4, 2)
    if not only or "relocate" in only:
        relocate_case()
    if not only or "kpp" in only:
        kpp_case()
    if not only or "unproject" in only:
        unproject_case()
    if not only or "default" in only:
        default_call_case()
