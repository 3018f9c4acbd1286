"""GPU, >= 2 devices: the sharded fit (fused NVLink exchange and the NCCL exchange) gives the same
centroids (bitwise), n_iter and labels as one GPU -- runs tools/multi_gpu_parity.py under torchrun.
Skipped on single-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_invariance():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multi_gpu_parity.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-30:])
    assert res.returncode == 0, tail
    assert "multi-GPU parity: all checks passed" in res.stdout, tail


def test_in_process_device_group(pkg):
    """Two devices driven from THIS process (one host thread each, `devices=[0, 1]`): what the
    reference's in-process caller (a napari worker thread, widget.py:116-147) would use.  Same
    labels, cloud and bit-identical centroids as one device."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    hm = pkg.make_stack(4, 300, 256, seed=3).numpy()
    for kw in (dict(init="k-means++", n_init=3, random_state=7, max_iter=15, tol=0.0),
               dict(init="random", n_init=2, random_state=3, max_iter=300, tol=1e-4),
               dict(init="k-means++", n_init=1, random_state=1, max_iter=20, tol=0.0, detrend=True, ground_level=True)):
        one = pkg.fuse_multiday_kmeans(hm, n_clusters=6, device=0, **kw)
        two = pkg.fuse_multiday_kmeans(hm, n_clusters=6, devices=[0, 1], **kw)
        assert two.extra["exchange"] == "nvlink_p2p" and two.extra["devices"] == [0, 1]
        assert two.n_points == one.n_points and two.n_iter == one.n_iter
        assert two.centroids.tobytes() == one.centroids.tobytes()
        np.testing.assert_array_equal(two.labels, one.labels)
        np.testing.assert_array_equal(two.fused_cloud, one.fused_cloud)
        np.testing.assert_array_equal(two.extra["segment_offsets"], one.extra["segment_offsets"])
        assert abs(two.inertia - one.inertia) <= 1e-12 * abs(one.inertia)
        if kw.get("ground_level"):
            np.testing.assert_array_equal(two.height_norm, one.height_norm)
            np.testing.assert_array_equal(two.extra["h_min"], one.extra["h_min"])
    # a reusable group, and the plugin wrapper on top of it
    with pkg.DeviceGroup([0, 1]) as grp:
        assert grp.p2p
        a = pkg.fuse_multiday_kmeans(hm, n_clusters=5, init="k-means++", random_state=0, max_iter=10, tol=0.0, engine=grp)
        b = pkg.fuse_multiday_kmeans(hm, n_clusters=5, init="k-means++", random_state=0, max_iter=10, tol=0.0, engine=grp)
        assert a.centroids.tobytes() == b.centroids.tobytes() and np.array_equal(a.labels, b.labels)
    layers = pkg.MultiDayFusionPlugin(n_clusters=4, devices=[0, 1], random_state=0, max_iter=5).run(hm)
    assert layers[0][2] == "points" and layers[0][0].shape == (a.n_points, 3)
