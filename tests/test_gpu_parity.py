"""GPU parity tests: every call goes through the C ABI (libmdkm.so via ctypes) and is checked
against the CPU oracle on identical inputs.  Bars (BASELINE.json north star): labels
bit-exact except near-ties within 1e-6 relative (counted and reported), centroids within
1e-5 relative (denominator max(|c|, per-dim data std))."""
import importlib

import numpy as np
import pytest

from oracle import c_oracle, kmeans_oracle as KO, unproject_oracle as UO

pytestmark = pytest.mark.gpu

synth = importlib.import_module("3d-point-cloud-multiday-imagery_b200.synth")

LABEL_BAND = 1e-6
CENTROID_TOL = 1e-5


def check_labels(P, centers, labels_ref, labels_gpu, allow_near=None):
    r = KO.compare_labels(P, centers, labels_ref, labels_gpu, rel_band=LABEL_BAND)
    print("label parity:", r)
    assert r["n_hard"] == 0, r
    if allow_near is not None:
        assert r["n_near_tie"] <= allow_near, r
    return r


def check_centroids(c_ref, c_gpu, P):
    err = KO.centroid_rel_err(c_ref, c_gpu, P.std(axis=0))
    print("centroid rel err:", err)
    assert err <= CENTROID_TOL, err
    return err


# ---------------------------------------------------------------------------------------
# K1 unprojection
# ---------------------------------------------------------------------------------------
def test_unproject_golden_disparity(engine, golden):
    g = golden("unproject_small.npz")
    n = engine.unproject(g["disparity"], g["mask"], disparity_scale=-1.0 / 16.0)
    assert n == g["points"].shape[0]
    cloud = engine.get_cloud(napari_order=False)
    np.testing.assert_array_equal(cloud.astype(np.float64), g["points"])  # bit-exact
    zyx = engine.get_cloud(napari_order=True)
    np.testing.assert_array_equal(zyx, cloud[:, ::-1])


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 5), (3, 17, 33), (2, 64, 64), (4, 130, 257), (1, 96, 4096)])
def test_unproject_ragged_shapes(engine, shape):
    D, H, W = shape
    hm = synth.make_stack(D, H, W, seed=D * 1000 + W, n_buildings=5).numpy()
    rs = np.random.RandomState(W)
    mask = rs.rand(D, H, W) > 0.1
    for m in (None, mask):
        P = UO.unproject_stack(hm, m)
        n = engine.unproject(hm, m)
        assert n == P.shape[0]
        if n:
            np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), P)


def test_unproject_all_invalid_and_empty(engine):
    assert engine.unproject(np.full((2, 8, 8), np.nan, dtype=np.float32)) == 0
    assert engine.unproject(np.full((1, 4, 4), 500.0, dtype=np.float32)) == 0
    assert engine.n_points == 0


def test_unproject_boundary_values(engine):
    hm = np.array([[[144.0, -144.0, 144.00002, np.inf, -np.inf, np.nan, 0.0, -0.0]]], dtype=np.float32)
    P = UO.unproject_stack(hm)
    assert engine.unproject(hm) == P.shape[0] == 4
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), P)


def test_unproject_device_input_and_sharded_slices(engine):
    import torch

    D, H, W = 3, 50, 70
    hm = synth.make_stack(D, H, W, seed=5, n_buildings=4)
    P = UO.unproject_stack(hm.numpy())
    assert engine.unproject(hm.cuda()) == P.shape[0]
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), P)
    # a rank's slice: rows [b, e) of the flattened stack, unaligned start
    flat = hm.numpy().reshape(-1)
    b, e = 70 * 13 + 0, 70 * 97
    n = engine.unproject(flat[b:e], stack_shape=(D, H, W), pix_begin=b)
    ys, xs = np.divmod(np.arange(b, e) % (H * W), W)
    ok = UO.valid_mask(flat[b:e])
    ref = np.stack([xs[ok], ys[ok], flat[b:e][ok]], axis=1).astype(np.float64)
    assert n == ref.shape[0]
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), ref)


def test_cuda_tensor_inputs_are_ordered_after_torch_stream(engine):
    """A CUDA tensor still being produced on torch's current stream when it is handed over: the
    handle's own stream must wait for the producer (and for the dtype conversion made on the way)."""
    import torch

    hm_ref = synth.make_stack(2, 512, 640, seed=5)
    mask_ref = torch.from_numpy(np.random.RandomState(5).rand(2, 512, 640) > 0.2)
    P = UO.unproject_stack(hm_ref.numpy(), mask_ref.numpy())
    hm_dev, mask_dev = hm_ref.cuda(), mask_ref.cuda()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        a = torch.randn(4096, 4096, device="cuda")
        for _ in range(30):  # keeps `side` busy for several milliseconds
            a = (a @ a) * 1e-4
        dep = (a[0, 0] != a[0, 0]).float() * 0.0  # 0.0, available only when the chain is done
        hm = torch.full_like(hm_dev, float("nan"))
        hm.copy_(hm_dev + dep)
        n = engine.unproject(hm, mask_dev)  # bool mask: converted to uint8 on `side`, too
        labels = torch.empty(n, dtype=torch.int32, device="cuda")
        init = synth.init_from_points(P.astype(np.float32), 4, 0)
        r = engine.fit(init, max_iter=3, tol=0.0, labels_out=labels)
        lab_dev = labels.clone()  # consumer on `side`, right after the call
    torch.cuda.synchronize()
    assert n == P.shape[0]
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), P)
    r2 = engine.fit(init, max_iter=3, tol=0.0)
    np.testing.assert_array_equal(lab_dev.cpu().numpy(), r2["labels"])
    assert r["labels"] is labels
    # device-resident cloud output
    out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
    engine.get_cloud(napari_order=False, out=out)
    np.testing.assert_array_equal(out.cpu().numpy().astype(np.float64), P)


def test_inputs_outlive_a_closed_engine(pkg):
    """A CUDA tensor that was handed to an engine must be releasable after that engine (and its
    stream) is gone: nothing may tie the tensor's release to the destroyed stream."""
    import gc

    import torch

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        hm = synth.make_stack(2, 256, 320, seed=3, device="cuda")
        with pkg.Engine(0) as solo:
            n = solo.unproject(hm)
            init = solo.gather_points(np.arange(4, dtype=np.int64) * (n // 5)).astype(np.float64)
            labels = torch.empty(n, dtype=torch.int32, device="cuda")
            solo.fit(init, max_iter=3, tol=0.0, labels_out=labels)
        del hm, labels
        gc.collect()
        torch.cuda.empty_cache()
        x = torch.ones(1 << 20, device="cuda").sum().item()  # the allocator still works
    torch.cuda.synchronize()
    assert x == float(1 << 20)


def test_output_buffers_must_be_usable_as_they_are(engine):
    hm = synth.make_stack(1, 32, 48, seed=1).numpy()
    n = engine.unproject(hm)
    with pytest.raises(ValueError, match="contiguous"):
        engine.get_cloud(out=np.empty((n, 6), dtype=np.float32)[:, ::2])
    with pytest.raises(ValueError, match="float32"):
        engine.get_cloud(out=np.empty((n, 3), dtype=np.float64))
    init = synth.init_from_points(engine.get_cloud(False), 3, 0)
    with pytest.raises(ValueError, match="int32"):
        engine.fit(init, max_iter=2, labels_out=np.empty(n, dtype=np.int64))
    out = np.empty(n, dtype=np.int32)
    r = engine.fit(init, max_iter=2, tol=0.0, labels_out=out)
    assert r["labels"] is out and np.array_equal(out, engine.fit(init, max_iter=2, tol=0.0)["labels"])


def test_unproject_detrend(engine, golden):
    g = golden("unproject_small.npz")
    n = engine.unproject(g["disparity"], g["mask"], disparity_scale=-1.0 / 16.0, detrend=True)
    ref = g["points_detrended"]
    assert n == ref.shape[0]
    cloud = engine.get_cloud(False).astype(np.float64)
    np.testing.assert_array_equal(cloud[:, :2], ref[:, :2])
    # z is the FP64 plane distance rounded to FP32 (tolerance: 1e-5 absolute on ~100 m heights)
    np.testing.assert_allclose(cloud[:, 2], ref[:, 2], rtol=0, atol=1e-4)
    hm = synth.make_stack(2, 200, 300, seed=9).numpy()
    ref = UO.unproject_stack(hm, detrend=True)
    assert engine.unproject(hm, detrend=True) == ref.shape[0]
    np.testing.assert_allclose(engine.get_cloud(False)[:, 2], ref[:, 2], rtol=0, atol=2e-5)


def test_unproject_reference_raster_files(pkg, engine, tmp_path):
    """f3: the reference's 5-out-F.tif rasters (disparity.py:213-224) go to the GPU as stored."""
    tio = importlib.import_module("3d-point-cloud-multiday-imagery_b200.tiff_io")
    rs = np.random.RandomState(5)
    paths, days = [], []
    for d, (h, w) in enumerate([(90, 130), (75, 141), (90, 141)]):
        bands = np.zeros((3, h, w), dtype=np.float32)
        bands[0] = synth.make_stack(1, h, w, seed=30 + d, n_buildings=4).numpy()[0]
        bands[0][rs.rand(h, w) < 0.02] = 18 * 16.0  # WLS sentinel: |h| > 144 (disparity.py:182-187)
        bands[2] = rs.rand(h, w) > 0.15
        p = str(tmp_path / f"{d}-5-out-F.tif")
        tio.write_tiff(p, bands, planar=(d == 1))
        paths.append(p)
        days.append(bands)
    stack = tio.load_height_rasters(paths)
    ref = UO.unproject_stack(stack[..., 0].astype(np.float64), stack[..., 2] != 0)
    n = engine.unproject(stack, raster_layout="gtiff3")
    assert n == ref.shape[0]
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), ref)
    # and with the per-day plane fit + extra mask, against the separate-plane path
    extra = rs.rand(*stack.shape[:3]) > 0.1
    n1 = engine.unproject(stack, extra, raster_layout="gtiff3", detrend=True)
    c1 = engine.get_cloud(False).copy()
    hm = np.ascontiguousarray(stack[..., 0])
    n2 = engine.unproject(hm, (stack[..., 2] != 0) & extra, detrend=True)
    assert n1 == n2
    np.testing.assert_array_equal(c1, engine.get_cloud(False))
    res = pkg.fuse_height_rasters(paths, n_clusters=4, init="k-means++", random_state=0, max_iter=10, engine=engine)
    assert res.n_points == ref.shape[0]
    np.testing.assert_array_equal(res.fused_cloud[:, ::-1].astype(np.float64), ref)


def test_streamed_cloud_matches_get_cloud(engine):
    """Slab pipeline (several slabs, day-aligned cuts when detrending) == the one-shot path."""
    import torch

    hm_dev = synth.make_stack(3, 1500, 1500, seed=11, device="cuda")
    hm = hm_dev.cpu().numpy()
    mask = np.random.RandomState(3).rand(*hm.shape) > 0.2
    for detrend in (False, True):
        for src, m in ((hm, None), (hm, mask), (hm_dev, None)):
            n, cloud = engine.unproject(src, m, detrend=detrend, stream_cloud="napari")
            engine.wait()
            ref = engine.get_cloud(napari_order=True, out=np.empty((n, 3), dtype=np.float32))
            assert n == cloud.shape[0]
            np.testing.assert_array_equal(cloud, ref)
            if not detrend:
                P = UO.unproject_stack(hm[:1], None if m is None else m[:1])
                np.testing.assert_array_equal(cloud[: P.shape[0], ::-1].astype(np.float64), P)
    n2, xyz = engine.unproject(hm, stream_cloud="xyz")
    engine.wait()
    np.testing.assert_array_equal(xyz, engine.get_cloud(False, out=np.empty((n2, 3), dtype=np.float32)))
    del hm_dev
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------
# K2+K3 single step
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["kmeans_stack_small.npz", "kmeans_c1_like.npz", "kmeans_tol.npz"])
def test_single_step_golden(engine, golden, name):
    g = golden(name)
    n = engine.unproject(g["height_maps"])
    assert n == int(g["n_points"])
    labels, sums, counts, n_ref = engine.lloyd_step(g["init"])
    P = UO.unproject_stack(g["height_maps"])
    check_labels(P, g["init"], g["step_labels"], labels, allow_near=0)
    np.testing.assert_array_equal(counts, g["step_counts"].astype(np.int64))
    c_gpu = sums / np.maximum(counts, 1)[:, None]
    check_centroids(g["step_centers"], c_gpu, P)
    print("refined in FP64:", n_ref)


@pytest.mark.parametrize("k", [1, 2, 7, 16, 64, 255, 256, 257, 300])
def test_single_step_vs_oracle_many_k(engine, k):
    hm = synth.make_stack(2, 160, 200, seed=k, n_buildings=8).numpy()
    P = UO.unproject_stack(hm)
    P32 = P.astype(np.float32)
    n = engine.unproject(hm)
    assert n == P.shape[0]
    rs = np.random.RandomState(k)
    C = P[np.sort(rs.choice(n, k, replace=False))] + rs.normal(0, 0.37, size=(k, 3))
    labels, sums, counts, _ = engine.lloyd_step(C)
    lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(P32[:, 0], P32[:, 1], P32[:, 2], C)
    check_labels(P, C, lab_ref, labels)
    agree = labels == lab_ref
    if agree.all():
        np.testing.assert_array_equal(counts, counts_ref.astype(np.int64))
        scale = np.abs(P).max(axis=0) * np.maximum(counts_ref, 1)[:, None]
        assert (np.abs(sums - sums_ref) / scale).max() < 1e-6


def test_single_step_exact_ties_lowest_index(engine):
    # points exactly equidistant from two centroids: strict '<' keeps the lower index (pyx:205-213)
    pts = np.array([[0.0, 0, 0], [2.0, 0, 0], [1.0, 0, 0], [1.0, 5, 0], [1.0, -3, 2]], dtype=np.float32)
    engine.set_points(pts)
    C = np.array([[0.0, 0, 0], [2.0, 0, 0]])
    labels, sums, counts, n_ref = engine.lloyd_step(C)
    assert labels.tolist() == [0, 1, 0, 0, 0]
    assert n_ref >= 3  # the three ties were decided by the FP64 refine
    labels, _, _, _ = engine.lloyd_step(C[::-1].copy())
    assert labels.tolist() == [1, 0, 0, 0, 0]
    # duplicated centroids: all points go to the first copy
    labels, _, counts, _ = engine.lloyd_step(np.array([[1.0, 0, 0], [1.0, 0, 0], [1.0, 0, 0]]))
    assert labels.tolist() == [0] * 5 and counts.tolist() == [5, 0, 0]


def test_long_one_sided_runs_do_not_overflow(engine):
    """Regression: boundary groups that all carry one label form long runs whose fixed-point
    coordinates share a sign; their warp totals exceed 32 bits (bug found at config-3 size)."""
    W, H = 4096, 1536
    ys, xs = np.divmod(np.arange(W * H, dtype=np.int64), W)
    X = np.stack([xs, ys, np.zeros_like(xs)], axis=1).astype(np.float32)
    engine.set_points(X)
    C = np.array([[2047.5, 255.5, 0.0], [2047.5, 767.5, 0.0], [2047.5, 1279.5, 0.0]])
    for _ in range(2):  # cold and with the cached mirror / summaries
        labels, sums, counts, _ = engine.lloyd_step(C)
        ref = np.minimum(ys // 512, 2)
        np.testing.assert_array_equal(labels, ref.astype(np.int32))
        np.testing.assert_array_equal(counts, np.bincount(ref, minlength=3))
        for d in range(2):
            np.testing.assert_array_equal(sums[:, d], np.bincount(ref, weights=X[:, d].astype(np.float64), minlength=3))
    init = C + [[0, 40, 0], [0, -30, 0], [0, 25, 0]]
    r = engine.fit(init, max_iter=6, tol=0.0)
    ref = KO.kmeans_fit(X.astype(np.float64), init, max_iter=6, tol=0.0)
    assert r["n_iter"] == ref["n_iter"] and r["n_relocations"] == 0
    np.testing.assert_array_equal(r["labels"], ref["labels"])
    np.testing.assert_allclose(r["centers"], ref["centers"], rtol=1e-7, atol=1e-6)  # z is constant: no std scale


def test_generic_float_cloud_random_order(engine):
    rs = np.random.RandomState(4)
    X32 = np.concatenate([rs.normal(c, 3.0, (5000, 3)) for c in rs.uniform(-1e3, 1e3, size=(9, 3))]).astype(np.float32)
    X32 = X32[rs.permutation(X32.shape[0])]
    X32 += np.array([2.5e4, -3.1e4, 512.0], dtype=np.float32)  # far from the origin
    X = X32.astype(np.float64)
    engine.set_points(X32)
    C = X[rs.choice(X.shape[0], 33, replace=False)]
    labels, sums, counts, _ = engine.lloyd_step(C)
    lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(X32[:, 0], X32[:, 1], X32[:, 2], C)
    check_labels(X, C, lab_ref, labels)
    if (labels == lab_ref).all():
        np.testing.assert_array_equal(counts, counts_ref.astype(np.int64))
    c_gpu = sums / np.maximum(counts, 1)[:, None]
    c_ref = sums_ref / np.maximum(counts_ref, 1)[:, None]
    check_centroids(c_ref, c_gpu, X)


# ---------------------------------------------------------------------------------------
# full fit
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["kmeans_stack_small.npz", "kmeans_c1_like.npz", "kmeans_tol.npz"])
def test_fit_golden(engine, golden, name):
    g = golden(name)
    engine.unproject(g["height_maps"])
    r = engine.fit(g["init"], max_iter=int(g["max_iter"]), tol=float(g["tol"]))
    P = UO.unproject_stack(g["height_maps"])
    assert r["n_iter"] == int(g["n_iter"])
    check_labels(P, g["centers"], g["labels"], r["labels"])
    check_centroids(g["centers"], r["centers"], P)
    np.testing.assert_allclose(r["inertia"], float(g["inertia"]), rtol=1e-6)


def test_fit_relocation_golden(engine, golden):
    g = golden("kmeans_relocate.npz")
    engine.set_points(g["X"])
    r = engine.fit(g["init"], max_iter=100, tol=0.0)
    assert r["n_relocations"] >= 1
    X = g["X"].astype(np.float64)
    assert r["n_iter"] == int(g["n_iter"])
    np.testing.assert_allclose(r["inertia"], float(g["inertia"]), rtol=1e-6)
    check_labels(X, g["centers"], g["labels"], r["labels"])
    check_centroids(g["centers"], r["centers"], X)


@pytest.mark.parametrize("max_iter", [1, 2, 3, 11])
def test_fit_relocation_at_any_iteration_count(engine, golden, max_iter):
    """An empty cluster is found by the update, and the update of an iteration is applied by the next
    launch -- or by the settle kernel when it was the last one (max_iter = 1 here: the host then
    relocates after the final pass was already enqueued and goes round again).  Against live
    scikit-learn on the relocation fixture's data."""
    from oracle import sklearn_ref

    if not sklearn_ref.available():
        pytest.skip("scikit-learn not importable")
    g = golden("kmeans_relocate.npz")
    X = g["X"].astype(np.float64)
    engine.set_points(g["X"])
    r = engine.fit(g["init"], max_iter=max_iter, tol=0.0)
    ref = sklearn_ref.fit(X, g["init"], max_iter=max_iter, tol=0.0)
    assert r["n_relocations"] >= 1
    assert r["n_iter"] == ref["n_iter"]
    check_labels(X, ref["centers"], ref["labels"], r["labels"])
    check_centroids(ref["centers"], r["centers"], X)
    np.testing.assert_allclose(r["inertia"], ref["inertia"], rtol=1e-6)  # (as test_fit_relocation_golden)


def test_fit_sklearn_known_answers(engine):
    # sklearn/cluster/tests/test_k_means.py:85-111, padded to d = 3 (unit weights)
    X = np.array([[0, 0, 0], [0.5, 0, 0], [0.5, 1, 0], [1, 1, 0]], dtype=np.float32)
    engine.set_points(X)
    r = engine.fit(np.array([[0.5, 0.5, 0], [3, 3, 0]], dtype=np.float64), max_iter=300, tol=1e-4)
    np.testing.assert_allclose(r["inertia"], 0.25, rtol=1e-9)
    assert r["n_iter"] == 3
    a = r["labels"].tolist() == [0, 0, 1, 1] and np.allclose(r["centers"], [[0.25, 0, 0], [0.75, 1, 0]])
    b = r["labels"].tolist() == [1, 1, 0, 0] and np.allclose(r["centers"], [[0.75, 1, 0], [0.25, 0, 0]])
    assert a or b


def test_fit_config1_vs_oracle(engine):
    """BASELINE.json configs[0]: 3-day 512x512 stack, k=8, 20 Lloyd iterations."""
    hm = synth.make_stack(3, 512, 512, seed=0).numpy()
    P = UO.unproject_stack(hm)
    n = engine.unproject(hm)
    assert n == P.shape[0]
    init = synth.init_from_points(P.astype(np.float32), 8, 0)
    r = engine.fit(init, max_iter=20, tol=0.0)
    ref = KO.kmeans_fit(P, init, max_iter=20, tol=0.0)
    assert r["n_iter"] == ref["n_iter"]
    check_labels(P, ref["centers"], ref["labels"], r["labels"])
    check_centroids(ref["centers"], r["centers"], P)
    np.testing.assert_allclose(r["inertia"], ref["inertia"], rtol=1e-6)
    print("n_refined", r["n_refined"], "of", n * r["n_iter"])


def test_fit_is_deterministic_and_idempotent(engine):
    hm = synth.make_stack(2, 300, 400, seed=3).numpy()
    n = engine.unproject(hm)
    P32 = engine.get_cloud(False)
    init = synth.init_from_points(P32, 12, 3)
    a = engine.fit(init, max_iter=15, tol=0.0)
    b = engine.fit(init, max_iter=15, tol=0.0)
    assert np.array_equal(a["labels"], b["labels"])
    assert a["centers"].tobytes() == b["centers"].tobytes()  # bitwise
    assert a["inertia"] == b["inertia"] and a["n_iter"] == b["n_iter"]
    # converge, then restart from the converged centroids: one iteration, nothing moves
    c = engine.fit(init, max_iter=300, tol=0.0)
    d = engine.fit(c["centers"], max_iter=300, tol=0.0)
    assert d["n_iter"] <= 2
    assert np.array_equal(c["labels"], d["labels"])
    np.testing.assert_allclose(c["centers"], d["centers"], rtol=0, atol=1e-9)
    assert n == a["labels"].shape[0]


def test_predict_matches_oracle_e_step(engine):
    hm = synth.make_stack(2, 200, 256, seed=6, n_buildings=7).numpy()
    P = UO.unproject_stack(hm)
    engine.unproject(hm)
    C = P[np.sort(np.random.RandomState(1).choice(P.shape[0], 24, replace=False))] + 0.123
    labels, inertia = engine.predict(C)
    mean = P.mean(axis=0)
    ref = KO.e_step(P - mean, C - mean)
    check_labels(P, C, ref, labels, allow_near=0)
    np.testing.assert_allclose(inertia, KO.inertia(P, C, ref), rtol=1e-9)


@pytest.mark.parametrize("shape,k", [((3, 40, 50), 5), ((2, 90, 127), 40), ((2, 64, 333), 300), ((1, 7, 1000), 70)])
def test_row_straddling_groups(engine, shape, k):
    """Groups of 128 raster-ordered points that run over the end of a row (several rows when
    the raster is narrow) are assigned in halves by the final pass; the mirror numbers its
    rows of cells in serpentine order.  Labels and inertia must not notice either."""
    D, H, W = shape
    hm = synth.make_stack(D, H, W, seed=W + k, n_buildings=6).numpy()
    P = UO.unproject_stack(hm)
    n = engine.unproject(hm)
    assert n == P.shape[0]
    rs = np.random.RandomState(W)
    C = P[np.sort(rs.choice(n, k, replace=False))] + rs.normal(0, 0.21, size=(k, 3))
    labels, inertia = engine.predict(C)
    mean = P.mean(axis=0)
    ref = KO.e_step(P - mean, C - mean)
    check_labels(P, C, ref, labels)
    if (labels == ref).all():
        np.testing.assert_allclose(inertia, KO.inertia(P, C, ref), rtol=1e-9)
    r = engine.fit(C, max_iter=8, tol=0.0)
    fit_ref = KO.kmeans_fit(P, C, max_iter=8, tol=0.0)
    assert r["n_iter"] == fit_ref["n_iter"]
    check_labels(P, fit_ref["centers"], fit_ref["labels"], r["labels"])
    check_centroids(fit_ref["centers"], r["centers"], P)
    np.testing.assert_allclose(r["inertia"], fit_ref["inertia"], rtol=1e-6)


@pytest.mark.parametrize("shape,k,rows,valid", [
    ((3, 40, 48), 5, None, 1.0), ((2, 90, 128), 40, None, 1.0), ((2, 64, 328), 16, None, 1.0),
    ((2, 64, 328), 20, None, 0.3), ((1, 7, 1000), 70, None, 1.0), ((4, 60, 64), 9, (30, 170), 1.0),
    ((4, 60, 64), 9, (50, 75), 1.0), ((3, 33, 8), 3, None, 0.5), ((6, 256, 512), 300, (100, 1400), 0.9),
])
def test_raster_mirror_equals_generic_mirror(engine, shape, k, rows, valid):
    """The run-table build of the tile-ordered mirror (raster clouds, mirror.cuh) against the
    generic histogram + scatter build and against the oracle: same labels, bitwise-equal centroids."""
    D, H, W = shape
    hm = synth.make_stack(D, H, W, seed=D * H + W + k, n_buildings=8).numpy()
    mask = None
    if valid < 1.0:
        mask = np.random.RandomState(k).rand(D, H, W) < valid
        mask[0, : H // 2] = False  # half a day without a single point
    r0, r1 = rows if rows else (0, D * H)
    sl = slice(r0 * W, r1 * W)
    kw = dict(stack_shape=shape, pix_begin=r0 * W)
    n = engine.unproject(hm.reshape(-1)[sl], None if mask is None else mask.reshape(-1)[sl], **kw)
    full = UO.unproject_stack(hm, mask)
    flat_valid = UO.valid_mask(hm, mask).reshape(-1)
    a, b = int(flat_valid[: r0 * W].sum()), int(flat_valid[: r1 * W].sum())
    P = full[a:b]
    assert n == P.shape[0]
    init = synth.init_from_points(P.astype(np.float32), k, 1)
    res = {}
    for on in (True, False):
        engine.raster_mirror(on)
        try:
            engine.drop_caches()
            res[on] = engine.fit(init, max_iter=12, tol=0.0)
            step = engine.lloyd_step(init)
        finally:
            engine.raster_mirror(True)
        if on:
            step_on = step
    assert res[True]["centers"].tobytes() == res[False]["centers"].tobytes()
    assert np.array_equal(res[True]["labels"], res[False]["labels"])
    assert res[True]["n_iter"] == res[False]["n_iter"] and res[True]["inertia"] == res[False]["inertia"]
    assert np.array_equal(step_on[0], step[0]) and np.array_equal(step_on[2], step[2])
    ref = KO.kmeans_fit(P, init, max_iter=12, tol=0.0)
    assert res[True]["n_iter"] == ref["n_iter"]
    check_labels(P, ref["centers"], ref["labels"], res[True]["labels"])
    check_centroids(ref["centers"], res[True]["centers"], P)
    np.testing.assert_array_equal(step_on[2], np.bincount(KO.lloyd_iter(P - P.mean(0), init - P.mean(0))[0], minlength=k))


@pytest.mark.parametrize("shape,k", [((3, 200, 256), 7), ((2, 300, 520), 40), ((1, 64, 2048), 300)])
def test_two_level_classification_equals_flat(engine, shape, k):
    """The classification pass walks super-groups first on large clouds and single groups on small
    ones (MDKM_OPT_TWO_LEVEL picks automatically): both must give the same fit, bit for bit."""
    cabi = importlib.import_module("3d-point-cloud-multiday-imagery_b200._cabi")
    hm = synth.make_stack(*shape, seed=k, n_buildings=10).numpy()
    n = engine.unproject(hm)
    init = synth.init_from_points(engine.get_cloud(False), k, 2)
    res = {}
    try:
        for mode in (0, 1):
            engine.set_option(cabi.OPT_TWO_LEVEL, mode)
            res[mode] = engine.fit(init, max_iter=15, tol=0.0)
            res[mode, "step"] = engine.lloyd_step(init)
    finally:
        engine.set_option(cabi.OPT_TWO_LEVEL, -1)
    assert res[0]["centers"].tobytes() == res[1]["centers"].tobytes() and res[0]["n_iter"] == res[1]["n_iter"]
    assert np.array_equal(res[0]["labels"], res[1]["labels"]) and res[0]["inertia"] == res[1]["inertia"]
    assert np.array_equal(res[0, "step"][0], res[1, "step"][0]) and np.array_equal(res[0, "step"][2], res[1, "step"][2])
    assert res[1]["worklist_groups"] == res[0]["worklist_groups"]  # the same groups reach the per-point pass
    P = UO.unproject_stack(hm)
    ref = KO.kmeans_fit(P, init, max_iter=15, tol=0.0)
    assert res[1]["n_iter"] == ref["n_iter"]
    check_labels(P, ref["centers"], ref["labels"], res[1]["labels"])


@pytest.mark.parametrize("shape,k", [((3, 200, 256), 7), ((2, 300, 520), 40), ((1, 64, 2048), 300)])
def test_dependent_launch_equals_cooperative_launch(engine, shape, k):
    """After the first (cooperative) launch of a fit the Lloyd kernels use programmatic stream
    serialisation (MDKM_OPT_DEPENDENT_LAUNCH, default on); with every launch cooperative the fit
    must be the same, bit for bit."""
    cabi = importlib.import_module("3d-point-cloud-multiday-imagery_b200._cabi")
    hm = synth.make_stack(*shape, seed=k + 1, n_buildings=10).numpy()
    engine.unproject(hm)
    init = synth.init_from_points(engine.get_cloud(False), k, 2)
    res = {}
    try:
        for mode in (0, 1):
            engine.set_option(cabi.OPT_DEPENDENT_LAUNCH, mode)
            res[mode] = engine.fit(init, max_iter=25, tol=0.0)
    finally:
        engine.set_option(cabi.OPT_DEPENDENT_LAUNCH, 1)
    assert res[0]["centers"].tobytes() == res[1]["centers"].tobytes() and res[0]["n_iter"] == res[1]["n_iter"]
    assert np.array_equal(res[0]["labels"], res[1]["labels"]) and res[0]["inertia"] == res[1]["inertia"]


@pytest.mark.parametrize("k", [5, 48, 300])
def test_deferred_update_iteration_counts(engine, k):
    """The centroid update of an iteration is applied by the NEXT launch (or by the settle kernel
    behind the last one).  Whatever max_iter is -- one iteration, the end of a batch of launches, one
    past it, convergence in the middle of a batch -- centroids, labels and n_iter must equal the
    oracle's."""
    hm = synth.make_stack(2, 160, 192, seed=11 + k, n_buildings=8).numpy()
    engine.unproject(hm)
    P = UO.unproject_stack(hm)
    init = synth.init_from_points(P.astype(np.float32), k, 4)
    for max_iter, tol in ((1, 0.0), (2, 0.0), (10, 0.0), (11, 0.0), (21, 0.0), (300, 1e-4), (300, 0.0)):
        res = engine.fit(init, max_iter=max_iter, tol=tol)
        ref = KO.kmeans_fit(P, init, max_iter=max_iter, tol=tol)
        assert res["n_iter"] == ref["n_iter"], (max_iter, tol, res["n_iter"], ref["n_iter"])
        check_labels(P, ref["centers"], ref["labels"], res["labels"])
        check_centroids(ref["centers"], res["centers"], P)


def test_fit_errors(engine):
    engine.set_points(np.zeros((3, 3), dtype=np.float32))
    with pytest.raises(Exception, match="n_samples=3 should be >= n_clusters=4"):
        engine.fit(np.zeros((4, 3)))
    with pytest.raises(Exception):
        engine.fit(np.zeros((2, 3)), max_iter=0)
    r = engine.fit(np.array([[1.0, 1, 1]]), max_iter=5, tol=0.0)  # k = 1
    assert r["labels"].tolist() == [0, 0, 0] and np.allclose(r["centers"], 0)


# ---------------------------------------------------------------------------------------
# a5: percentile ground-levelling (plugin.py:181-192), per day
# ---------------------------------------------------------------------------------------
def _check_ground_level(engine, hm, mask=None):
    P, hn_ref, lo_ref, hi_ref, off_ref = UO.reference_tail_stack(hm, mask, detrend=False)
    n = engine.unproject(hm, mask)
    assert n == P.shape[0]
    np.testing.assert_array_equal(engine.segment_offsets, off_ref)
    lo, hi, hn = engine.ground_level(True)
    # z is exact in FP32 here, so the order statistics and numpy's lerp are reproduced exactly
    np.testing.assert_array_equal(lo, lo_ref)
    np.testing.assert_array_equal(hi, hi_ref)
    cloud = engine.get_cloud(False)
    np.testing.assert_array_equal(cloud[:, :2].astype(np.float64), P[:, :2])
    np.testing.assert_array_equal(cloud[:, 2], P[:, 2].astype(np.float32))       # one rounding of z - h_min
    np.testing.assert_array_equal(hn, hn_ref.astype(np.float32))                  # idem for h_norm
    return n


def test_ground_level_per_day_exact(engine):
    hm = synth.make_stack(3, 120, 200, seed=21, n_buildings=9).numpy()
    _check_ground_level(engine, hm)
    rs = np.random.RandomState(2)
    _check_ground_level(engine, hm, rs.rand(*hm.shape) > 0.3)


def test_ground_level_tiny_and_empty_days(engine):
    hm = np.full((5, 3, 7), np.nan, dtype=np.float32)
    hm[0, 1, 2] = 3.5                       # one point
    hm[1, 0, :2] = [1.0, -2.0]              # two points
    hm[3].flat[:9] = np.arange(9) * 0.25    # day 2 and day 4 stay empty
    P, hn_ref, lo_ref, hi_ref, off_ref = UO.reference_tail_stack(hm, detrend=False)
    assert engine.unproject(hm) == 12
    np.testing.assert_array_equal(engine.segment_offsets, off_ref)
    lo, hi, hn = engine.ground_level(True)
    np.testing.assert_array_equal(lo, lo_ref)  # NaN for the empty days on both sides
    np.testing.assert_array_equal(hi, hi_ref)
    np.testing.assert_array_equal(hn, hn_ref.astype(np.float32))
    np.testing.assert_array_equal(engine.get_cloud(False)[:, 2], P[:, 2].astype(np.float32))


def test_reference_tail_golden(engine, golden):
    """Unprojection + detrend + ground level, the reference's whole per-pair tail."""
    g = golden("unproject_small.npz")
    n = engine.unproject(g["disparity"], g["mask"], disparity_scale=-1.0 / 16.0, detrend=True)
    assert n == g["tail_points"].shape[0]
    lo, hi, hn = engine.ground_level(True)
    # the detrended z is the FP64 plane distance rounded to FP32 (1e-4 absolute, see above)
    np.testing.assert_allclose(lo, g["tail_h_min"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(hi, g["tail_h_max"], rtol=0, atol=1e-4)
    cloud = engine.get_cloud(False).astype(np.float64)
    np.testing.assert_array_equal(cloud[:, :2], g["tail_points"][:, :2])
    np.testing.assert_allclose(cloud[:, 2], g["tail_points"][:, 2], rtol=0, atol=2e-4)
    np.testing.assert_allclose(hn, g["tail_height_norm"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_reference_lines_golden(engine, golden, case):
    """The CUDA path against outputs of the reference's OWN source lines (plugin.py:148-192,
    exec'd verbatim by tests/golden/make_unproject_ref.py): int16 disparity + validity mask in,
    points_coords (napari z,y,x) + the 'height' property out."""
    g = golden("unproject_ref.npz")
    disp, vm = g[f"{case}_disparity"], g[f"{case}_validity_mask"]
    # a1-a3: h = -disp/16, validity, np.where order -- bit-exact
    n = engine.unproject(disp[None], vm[None], disparity_scale=-1.0 / 16.0)
    assert n == int(g[f"{case}_valid_mask"].sum())
    np.testing.assert_array_equal(engine.get_cloud(False).astype(np.float64), g[f"{case}_P"])
    # a4: plane detrend (FP64 moments + Jacobi on the device, z stored as FP32)
    n = engine.unproject(disp[None], vm[None], disparity_scale=-1.0 / 16.0, detrend=True)
    cloud = engine.get_cloud(False).astype(np.float64)
    np.testing.assert_array_equal(cloud[:, :2], g[f"{case}_P"][:, :2])
    np.testing.assert_allclose(cloud[:, 2], g[f"{case}_height_rel"], rtol=0, atol=1e-4)
    # a5: percentile levelling and the colour property
    lo, hi, hn = engine.ground_level(True)
    np.testing.assert_allclose(lo[0], float(g[f"{case}_h_min"]), rtol=0, atol=1e-4)
    np.testing.assert_allclose(hi[0], float(g[f"{case}_h_max"]), rtol=0, atol=1e-4)
    np.testing.assert_allclose(hn, g[f"{case}_h_norm"], rtol=0, atol=1e-5)
    zyx = engine.get_cloud(napari_order=True).astype(np.float64)
    ref = g[f"{case}_points_coords"]
    np.testing.assert_array_equal(zyx[:, 1:], ref[:, 1:])
    np.testing.assert_allclose(zyx[:, 0], ref[:, 0], rtol=0, atol=2e-4)


def test_ground_level_needs_whole_days(engine):
    hm = synth.make_stack(2, 16, 16, seed=1, n_buildings=1).numpy().reshape(-1)
    engine.unproject(hm[5:300], stack_shape=(2, 16, 16), pix_begin=5)
    assert len(engine.segment_offsets) == 3
    with pytest.raises(Exception, match="whole days"):
        engine.ground_level()


# ---------------------------------------------------------------------------------------
# a14: k-means++ seeding and the reference's default call
# ---------------------------------------------------------------------------------------
def test_kmeanspp_golden(engine, golden):
    g = golden("kmeanspp.npz")
    engine.set_points(g["X"])
    for seed, k in ((0, 4), (5, 8), (42, 16)):
        centers, idx = engine.kmeans_plusplus(k, np.random.RandomState(seed))
        np.testing.assert_array_equal(idx, g[f"indices_{seed}_{k}"])
        np.testing.assert_array_equal(centers, g[f"centers_{seed}_{k}"])


@pytest.mark.parametrize("k,seed", [(1, 0), (2, 1), (32, 2), (200, 3)])
def test_kmeanspp_vs_oracle_stack(engine, k, seed):
    hm = synth.make_stack(2, 160, 200, seed=seed, n_buildings=8).numpy()
    P = UO.unproject_stack(hm)
    engine.unproject(hm)
    c_ref, i_ref = KO.kmeans_plusplus(P - P.mean(axis=0), k, np.random.RandomState(seed))
    centers, idx = engine.kmeans_plusplus(k, np.random.RandomState(seed))
    np.testing.assert_array_equal(idx, i_ref)
    np.testing.assert_array_equal(centers, P[i_ref])


def test_kmeanspp_is_deterministic(engine):
    hm = synth.make_stack(4, 300, 500, seed=8).numpy()
    engine.unproject(hm)
    a = engine.kmeans_plusplus(64, np.random.RandomState(9))
    b = engine.kmeans_plusplus(64, np.random.RandomState(9))
    np.testing.assert_array_equal(a[1], b[1])
    assert len(set(a[1].tolist())) == 64


@pytest.mark.parametrize("k,n_init", [(5, 10), (3, 1)])
def test_default_call_golden(pkg, engine, golden, k, n_init):
    """KMeans(n_clusters, random_state=42, n_init=10) -- the reference's call, core.py:227-228."""
    g = golden("kmeans_default_call.npz")
    P = UO.unproject_stack(g["height_maps"])
    for init, pre in (("k-means++", ""), ("random", "r")):
        res = pkg.fuse_multiday_kmeans(g["height_maps"], n_clusters=k, random_state=42, n_init=n_init,
                                       init=init, engine=engine)
        assert res.n_iter == int(g[f"{pre}n_iter_{k}"])
        check_labels(P, g[f"{pre}centers_{k}"], g[f"{pre}labels_{k}"], res.labels, allow_near=0)
        check_centroids(g[f"{pre}centers_{k}"], res.centroids, P)
        np.testing.assert_allclose(res.inertia, float(g[f"{pre}inertia_{k}"]), rtol=1e-6)


def test_fuse_with_reference_tail(pkg, engine):
    hm = synth.make_stack(3, 128, 160, seed=4, n_buildings=6).numpy()
    res = pkg.fuse_multiday_kmeans(hm, n_clusters=6, init="k-means++", random_state=1, max_iter=30, tol=0.0,
                                   detrend=True, ground_level=True, engine=engine)
    assert res.height_norm is not None and res.height_norm.shape == (res.n_points,)
    assert res.height_norm.min() == 0.0 and res.height_norm.max() == 1.0
    Pt, hn_ref, lo, hi, off = UO.reference_tail_stack(hm)
    np.testing.assert_array_equal(res.extra["segment_offsets"], off)
    np.testing.assert_allclose(res.extra["h_min"], lo, atol=1e-4)
    np.testing.assert_allclose(res.fused_cloud[:, 0], Pt[:, 2], rtol=0, atol=2e-4)   # napari (z,y,x)
    np.testing.assert_allclose(res.height_norm, hn_ref, rtol=0, atol=1e-5)
    # k-means parity on the cloud the device actually clustered
    X = res.fused_cloud[:, ::-1].astype(np.float64)
    ref = KO.kmeans_default_call(X, 6, 1, n_init=1, max_iter=30, tol=0.0)
    assert res.n_iter == ref["n_iter"]
    check_labels(X, ref["centers"], ref["labels"], res.labels)
    check_centroids(ref["centers"], res.centroids, X)
    layers = pkg.to_layers(res)
    assert layers[0][1]["properties"]["height"] is res.height_norm


# ---------------------------------------------------------------------------------------
# full-size properties (BASELINE.json configs[1]: 10 x 2048 x 2048, k = 16)
# ---------------------------------------------------------------------------------------
def test_config2_full_size_single_step(engine):
    import torch

    D, H, W, k = 10, 2048, 2048, 16
    hm = synth.make_stack(D, H, W, seed=0, device="cuda")
    n = engine.unproject(hm)
    hm_h = hm.cpu().numpy()
    del hm
    torch.cuda.empty_cache()
    ok = UO.valid_mask(hm_h)
    assert n == int(ok.sum())
    cloud = engine.get_cloud(False)
    # np.where order: spot-check against the oracle on the first and last day
    P0 = UO.unproject_stack(hm_h[:1])
    np.testing.assert_array_equal(cloud[: P0.shape[0]].astype(np.float64), P0)
    Pl = UO.unproject_stack(hm_h[-1:])
    np.testing.assert_array_equal(cloud[-Pl.shape[0]:].astype(np.float64), Pl)
    init = synth.init_from_points(cloud, k, 0)
    labels, sums, counts, n_ref = engine.lloyd_step(init)
    lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(
        np.ascontiguousarray(cloud[:, 0]), np.ascontiguousarray(cloud[:, 1]),
        np.ascontiguousarray(cloud[:, 2]), init)
    bad = np.nonzero(labels != lab_ref)[0]
    print("full-size step: mismatches", bad.size, "refined", n_ref, "of", n)
    r = KO.compare_labels(cloud[bad].astype(np.float64), init, lab_ref[bad], labels[bad])
    assert r["n_hard"] == 0, r
    assert counts.sum() == n
    if bad.size == 0:
        np.testing.assert_array_equal(counts, counts_ref.astype(np.int64))
    c_gpu = sums / counts[:, None]
    c_ref = sums_ref / counts_ref[:, None]
    std = np.array([cloud[:, i].std(dtype=np.float64) for i in range(3)])
    err = (np.abs(c_gpu - c_ref) / np.maximum(np.abs(c_ref), std)).max()
    print("full-size centroid rel err", err)
    assert err < CENTROID_TOL
    # checksum-of-sums property: total of the per-cluster sums equals the column sums
    col = np.array([cloud[:, i].sum(dtype=np.float64) for i in range(3)])
    np.testing.assert_allclose(sums.sum(axis=0), col, rtol=1e-9)


def test_config2_full_size_fit(engine):
    """BASELINE.json configs[1] at full size and its full 20 iterations (cold mirror / summaries,
    fused update, final pass) against the C oracle iterated step by step."""
    import torch

    D, H, W, k, iters = 10, 2048, 2048, 16, 20
    hm = synth.make_stack(D, H, W, seed=1, device="cuda")
    n = engine.unproject(hm)
    del hm
    torch.cuda.empty_cache()
    cloud = engine.get_cloud(False)
    x, y, z = (np.ascontiguousarray(cloud[:, i]) for i in range(3))
    init = synth.init_from_points(cloud, k, 1)
    C = init.copy()
    for _ in range(iters):
        _, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
        C = sums_ref / np.maximum(counts_ref, 1)[:, None]
    lab_ref, _, counts_final, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)  # final E-step (not strict)
    r = engine.fit(init, max_iter=iters, tol=0.0)
    assert r["n_iter"] == iters and r["n_relocations"] == 0
    std = np.array([cloud[:, i].std(dtype=np.float64) for i in range(3)])
    err = (np.abs(r["centers"] - C) / np.maximum(np.abs(C), std)).max()
    print("full-size fit: centroid rel err", err, "refined", r["n_refined"], "of", n * iters)
    assert err < CENTROID_TOL
    bad = np.nonzero(r["labels"] != lab_ref)[0]
    cmp = KO.compare_labels(cloud[bad].astype(np.float64), C, lab_ref[bad], r["labels"][bad])
    print("full-size fit: label mismatches", bad.size, cmp)
    assert cmp["n_hard"] == 0
    np.testing.assert_array_equal(np.bincount(r["labels"], minlength=k)[: k] if bad.size == 0 else counts_final,
                                  counts_final.astype(np.int64))
    # the same fit again with the cached structures: bit-identical
    r2 = engine.fit(init, max_iter=iters, tol=0.0)
    assert r2["centers"].tobytes() == r["centers"].tobytes() and np.array_equal(r2["labels"], r["labels"])
    d = ((cloud.astype(np.float64) - r["centers"][r["labels"]]) ** 2).sum()
    np.testing.assert_allclose(r["inertia"], d, rtol=1e-9)


# ---------------------------------------------------------------------------------------
# the other BASELINE.json configurations, each on a size the C oracle finishes in seconds
# ---------------------------------------------------------------------------------------
def _oracle_iterate(x, y, z, init, iters):
    """`iters` Lloyd iterations of the C oracle (sklearn/_k_means_lloyd.pyx restated), then the
    final E-step of _kmeans.py:742-756.  Returns (centres, labels, counts)."""
    C = init.copy()
    for _ in range(iters):
        _, sums, counts, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
        assert counts.min() > 0  # no relocation in these runs
        C = sums / counts[:, None]
    lab, _, counts, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
    return C, lab, counts


def _check_fit_vs_oracle(engine, cloud, init, iters, tag):
    x, y, z = (np.ascontiguousarray(cloud[:, i]) for i in range(3))
    k = init.shape[0]
    C, lab_ref, counts_ref = _oracle_iterate(x, y, z, init, iters)
    r = engine.fit(init, max_iter=iters, tol=0.0)
    assert r["n_iter"] == iters and r["n_relocations"] == 0
    std = np.array([cloud[:, i].std(dtype=np.float64) for i in range(3)])
    err = (np.abs(r["centers"] - C) / np.maximum(np.abs(C), std)).max()
    bad = np.nonzero(r["labels"] != lab_ref)[0]
    cmp = KO.compare_labels(cloud[bad].astype(np.float64), C, lab_ref[bad], r["labels"][bad], rel_band=LABEL_BAND)
    settled = 1.0 - r["worklist_groups"] / max(1, r["groups"] * r["n_iter"])
    print(f"{tag}: N={cloud.shape[0]} k={k} iters={iters} centroid rel err {err:.2e} label mismatches {bad.size} "
          f"near-ties {cmp['n_near_tie']} hard {cmp['n_hard']} refined {r['n_refined']} settled group-iterations {settled:.3f}")
    assert err < CENTROID_TOL
    assert cmp["n_hard"] == 0, cmp
    if bad.size == 0:
        np.testing.assert_array_equal(np.bincount(r["labels"], minlength=k), counts_ref.astype(np.int64))
    return r, C


def test_config4_k1024_step_by_step(engine):
    """BASELINE.json configs[3] code path (k = 1024: uint16 labels, bucket index, shared-memory
    atomics) on one 4096 x 4096 day: single steps against the C oracle, then a 3-iteration fit."""
    import torch

    H = W = 4096
    k = 1024
    hm = synth.make_stack(1, H, W, seed=4, device="cuda")
    n = engine.unproject(hm)
    del hm
    torch.cuda.empty_cache()
    cloud = engine.get_cloud(False)
    x, y, z = (np.ascontiguousarray(cloud[:, i]) for i in range(3))
    C = synth.init_from_points(cloud, k, 4)
    init = C.copy()
    for step in range(2):  # step 0 from k data points (exact ties possible), step 1 from real centroids
        labels, sums, counts, n_ref = engine.lloyd_step(C)
        lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
        bad = np.nonzero(labels != lab_ref)[0]
        cmp = KO.compare_labels(cloud[bad].astype(np.float64), C, lab_ref[bad], labels[bad], rel_band=LABEL_BAND)
        print(f"k=1024 step {step}: mismatches {bad.size} near-ties {cmp['n_near_tie']} hard {cmp['n_hard']} "
              f"refined {n_ref} of {n}")
        assert cmp["n_hard"] == 0, cmp
        assert counts.sum() == n
        if bad.size == 0:
            np.testing.assert_array_equal(counts, counts_ref.astype(np.int64))
            np.testing.assert_allclose(sums, sums_ref, rtol=1e-6, atol=1e-2)
        C = sums_ref / np.maximum(counts_ref, 1)[:, None]
    _check_fit_vs_oracle(engine, cloud, init, 3, "c4-path fit")
    # settling disabled (every point through the per-point pass): bit-identical results
    r1 = engine.fit(init, max_iter=3, tol=0.0)
    engine.settle_groups(False)
    try:
        r2 = engine.fit(init, max_iter=3, tol=0.0)
    finally:
        engine.settle_groups(True)
    assert r2["centers"].tobytes() == r1["centers"].tobytes() and np.array_equal(r2["labels"], r1["labels"])
    assert r2["worklist_groups"] == r2["groups"] * 3 and r1["worklist_groups"] < r2["worklist_groups"]


def test_config5_tol_run_matches_sklearn_n_iter(engine):
    """BASELINE.json configs[4] style: tol = 1e-4, max_iter = 300, k = 32 on a mid-size stack,
    against LIVE scikit-learn (the reference's implementation): same n_iter, labels, centroids."""
    from oracle import sklearn_ref

    if not sklearn_ref.available():
        pytest.skip("scikit-learn not importable")
    hm = synth.make_stack(4, 1024, 1024, seed=6).numpy()
    n = engine.unproject(hm)
    cloud = engine.get_cloud(False)
    X = cloud.astype(np.float64)
    init = synth.init_from_points(cloud, 32, 6)
    ref = sklearn_ref.fit(X, init, max_iter=300, tol=1e-4)
    r = engine.fit(init, max_iter=300, tol=1e-4)
    print(f"c5-style: N={n} sklearn n_iter {ref['n_iter']} ours {r['n_iter']} refined {r['n_refined']} "
          f"tol_scaled {r['tol_scaled']:.6e}")
    assert r["n_iter"] == ref["n_iter"]
    assert 1 < r["n_iter"] < 300
    check_labels(X, ref["centers"], ref["labels"], r["labels"])
    check_centroids(ref["centers"], r["centers"], X)
    np.testing.assert_allclose(r["inertia"], ref["inertia"], rtol=1e-9)


def test_config3_row_band_shard(engine):
    """BASELINE.json configs[2] geometry: one rank's row band of an 8192-wide frame, starting in
    the middle of a day (pix_begin > 0) and running into the next one; k = 64, 20 iterations."""
    import torch

    D, H, W, k, iters = 2, 8192, 8192, 64, 20
    rows = 1024
    pix_begin = (H - rows // 2) * W              # the last 512 rows of day 0 ...
    count = rows * W                             # ... and the first 512 rows of day 1
    hm = synth.make_stack(1, rows, W, seed=8, device="cuda").reshape(-1)
    n = engine.unproject(hm, stack_shape=(D, H, W), pix_begin=pix_begin)
    hm_h = hm.cpu().numpy()
    del hm
    torch.cuda.empty_cache()
    idx = np.flatnonzero(UO.valid_mask(hm_h)) + pix_begin  # np.where order of the global stack
    rem = idx % (H * W)
    P = np.stack([(rem % W).astype(np.float64), (rem // W).astype(np.float64), hm_h[idx - pix_begin].astype(np.float64)], axis=1)
    assert n == P.shape[0]
    cloud = engine.get_cloud(False)
    np.testing.assert_array_equal(cloud.astype(np.float64), P)
    off = engine.segment_offsets
    assert off.tolist() == [0, int((idx < H * W).sum()), n]
    init = synth.init_from_points(cloud, k, 8)
    _check_fit_vs_oracle(engine, cloud, init, iters, "c3 row band")
