"""Public Python API: the drop-in for the reference's multi-day point-cloud fusion step.

The reference has no such routine of its own; the signature below is
assembled from the three anchors it does have:

* the plugin contract ``SatellitePlugin.run(...) -> List[Layer]`` (``interface.py:10-47``,
  ``Layer = (ndarray, params, type)`` at ``interface.py:5-7``),
* the unprojection tail of ``HeightMapExtractor.run``
  (``members/rafael/disparity/plugin.py:147-233``), whose Points layer this module
  reproduces (napari (z, y, x) column order, ``plugin.py:192``; layer params ``:220-233``),
* scikit-learn's ``KMeans`` call convention as the reference uses it
  (``members/jasraj/land_use_classification/core.py:227-228``): ``n_clusters``, ``init``,
  ``n_init``, ``max_iter=300``, ``tol=1e-4``, ``random_state``; results named like
  ``labels_ / cluster_centers_ / inertia_ / n_iter_``.

Everything numeric happens in libmdkm.so (sm_100a CUDA kernels); this file only validates
arguments, draws the host-side random numbers scikit-learn would draw, and packs results.
"""
from __future__ import annotations

import traceback
from dataclasses import dataclass, field
from typing import Any, Dict, List, Literal, Optional, Tuple

import numpy as np

from .engine import Engine

LayerType = Literal["image", "labels", "points", "shapes"]
Layer = Tuple[np.ndarray, Dict[str, Any], LayerType]  # interface.py:5-7

MAX_ABS_HEIGHT = 144.0  # MAX_DISP / 2 (constants.py:54, plugin.py:151)


@dataclass
class FusionResult:
    """What the fused k-means returns (names follow scikit-learn's fitted attributes)."""

    labels: np.ndarray           # int32 [N], np.where order per day, days concatenated
    centroids: np.ndarray        # float64 [K,3] as (x, y, z)
    fused_cloud: Optional[np.ndarray]  # float32 [N,3] in napari (z, y, x) order, or None
    n_iter: int
    inertia: float
    n_points: int
    n_refined: int = 0           # point-iterations decided by the FP64 refine (near-ties)
    n_relocations: int = 0
    height_norm: Optional[np.ndarray] = None  # plugin.py:184-188 'height' property, if requested
    extra: Dict[str, Any] = field(default_factory=dict)

    # scikit-learn style aliases
    @property
    def labels_(self):
        return self.labels

    @property
    def cluster_centers_(self):
        return self.centroids

    @property
    def inertia_(self):
        return self.inertia

    @property
    def n_iter_(self):
        return self.n_iter


def _check_random_state(seed):
    if seed is None or isinstance(seed, (int, np.integer)):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError("random_state must be None, an int or a numpy RandomState")


def _label_mapping(l1, l2, k):
    """The map label-of-l1 -> label-of-l2 by first occurrence, and whether every point obeys it
    (the loop of sklearn/cluster/_k_means_common.pyx:314-330 on one shard)."""
    mapping = np.full(k, -1, dtype=np.int64)
    if l1.shape[0]:
        uniq, first = np.unique(l1, return_index=True)
        mapping[uniq] = l2[first]
    return mapping, bool(np.array_equal(mapping[l1], l2))


def _merge_label_mappings(mappings, oks) -> bool:
    """Do the shards' maps describe ONE map?  (labels equal up to a permutation, over all shards)"""
    merged = np.full(mappings[0].shape[0], -1, dtype=np.int64)
    for m in mappings:
        m = np.asarray(m, dtype=np.int64)
        both = (m >= 0) & (merged >= 0)
        if np.any(m[both] != merged[both]):
            return False
        merged = np.where(m >= 0, m, merged)
    return bool(all(oks))


def _is_same_clustering(l1, l2, k, gather=None) -> bool:
    """sklearn/cluster/_k_means_common.pyx:314-330 (labels equal up to a permutation).  With
    sharded labels ``gather(mapping, ok) -> (mappings, oks)`` collects every rank's shard map
    (a collective: all ranks call it), and the verdict is the same on every rank."""
    mapping, ok = _label_mapping(l1, l2, k)
    if gather is None:
        return ok
    return _merge_label_mappings(*gather(mapping, ok))


def _random_seeds(rs, n: int, k: int) -> np.ndarray:
    """``random_state.choice(n, size=k, replace=False, p=w / w.sum())`` with unit weights
    (sklearn/_kmeans.py:1014-1021), reproduced bit for bit for any ``n`` without the n-sized
    probability vector (``_npdraw.py``)."""
    from ._npdraw import choice_uniform_without_replacement

    return choice_uniform_without_replacement(rs, n, k)


def _run_kmeans(eng: Engine, *, n_clusters, init, n_init, max_iter, tol, random_state, want_labels=True,
                labels_out=None, gather_mappings=None):
    """KMeans.fit driver (sklearn/_kmeans.py:1436-1563): init, n_init restarts, best inertia.

    ``labels_out``: optional int32 array that receives this rank's labels of the best run.
    ``gather_mappings``: with a multi-rank engine, the collective that lets the ranks agree on
    sklearn's "same clustering up to a permutation" test (see ``_is_same_clustering``); when it is
    missing the test is skipped on multi-rank engines (the inertia comparison alone decides)."""
    n = eng.n_points_global  # with a communicator: every rank draws the same seeds
    k = int(n_clusters)
    if k < 1:
        raise ValueError("n_clusters must be >= 1")
    if n < k:
        raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
    rs = _check_random_state(random_state)
    init_is_array = not isinstance(init, str)
    if init_is_array:
        init_arr = np.ascontiguousarray(init, dtype=np.float64)
        if init_arr.shape != (k, 3):
            raise ValueError(f"The shape of the initial centers {init_arr.shape} does not match (n_clusters, 3)")
        n_init = 1  # sklearn/_kmeans.py:905-913: explicit init => single run
    elif init not in ("k-means++", "random"):
        raise ValueError("init must be 'k-means++', 'random' or an array of shape (n_clusters, 3)")
    if n_init == "auto":
        n_init = 1 if init == "k-means++" else 10  # sklearn/_kmeans.py:896-903
    if eng.n_ranks > 1 and gather_mappings is None:
        from .dist import torch_gather_mappings

        gather_mappings = torch_gather_mappings(eng)  # None unless torch.distributed spans the ranks
    single = int(n_init) == 1
    best = None
    for _ in range(int(n_init)):
        if init_is_array:
            centers0 = init_arr
        elif init == "random":
            centers0 = eng.gather_points(_random_seeds(rs, n, k)).astype(np.float64)
        else:
            centers0, _ = eng.kmeans_plusplus(k, rs)
        r = eng.fit(centers0, max_iter=max_iter, tol=tol, want_labels=want_labels,
                    labels_out=labels_out if (single and want_labels) else None)
        # sklearn/_kmeans.py:1534-1541: a run replaces the best one when its inertia is lower AND it
        # is not the same clustering up to a permutation.  The inertia is global, so every rank
        # takes the same branch and the (collective) permutation test pairs up.
        better = best is not None and r["inertia"] < best["inertia"]
        same = False
        if better and want_labels and (eng.n_ranks == 1 or gather_mappings is not None):
            same = _is_same_clustering(r["labels"], best["labels"], k, gather_mappings if eng.n_ranks > 1 else None)
        if best is None or (better and not same):
            best = r
            if not single and r["labels"] is not None:
                best = dict(r, labels=r["labels"].copy())  # result buffers are reused by the next run
    if labels_out is not None and want_labels and not single:
        labels_out[...] = best["labels"]
        best = dict(best, labels=labels_out)
    return best


def fuse_multiday_kmeans(height_maps, valid_masks=None, *, n_clusters=8, init="k-means++", n_init=1,
                         max_iter=300, tol=1e-4, random_state=None, max_abs_height=MAX_ABS_HEIGHT,
                         detrend=False, disparity_scale=None, ground_level=False, return_cloud=True,
                         device=0, devices=None, engine=None, stack_shape=None,
                         pix_begin=0, raster_layout=None) -> FusionResult:
    """Unproject a multi-day height-map stack into one XYZ cloud and cluster it (Lloyd).

    height_maps : float32 ``[D,H,W]`` (NaN = nodata), numpy or torch (CPU or CUDA); or int16
        OpenCV fixed-point disparity with ``disparity_scale=-1/16`` (plugin.py:147-148).
    valid_masks : optional bool/uint8 ``[D,H,W]`` (``final_defined`` of disparity.py:203-204).
    n_clusters, init, n_init, max_iter, tol, random_state : as sklearn.cluster.KMeans.
    detrend : apply the per-day plane fit of plugin.py:161-171 to z before clustering.
    ground_level : per day, shift z so that its 2nd percentile is 0 and return the 'height'
        colour property (plugin.py:181-192), before clustering -- with ``detrend=True`` this
        is the reference's whole per-pair tail, applied to every day of the stack.
    raster_layout : "gtiff3" when ``height_maps`` is ``[D,H,W,3]`` float32 in the pixel layout of the
        reference's ``5-out-F.tif`` (see ``fuse_height_rasters``).
    engine : reuse an existing ``Engine`` (keeps device buffers, and with
        ``Engine(pinned_results=True)`` page-locked result buffers, across calls); with a
        multi-rank engine pass this rank's flat slice plus ``stack_shape`` / ``pix_begin``.
        A ``DeviceGroup`` is accepted too (see ``devices``).
    devices : several CUDA devices used from THIS process (``[0, 1, 2, 3]``): the host stack is
        sharded over them by row bands, one host thread drives each device, the devices exchange
        their K x 4 partial sums over NVLink inside the Lloyd kernel, and the results are
        assembled in the reference's point order -- bit-identical to one device.  This is how the
        in-process caller of the reference (a napari worker thread, widget.py:116-147) gets more
        than one GPU without ``torchrun``.
    Returns a FusionResult; ``fused_cloud`` is float32 ``[N,3]`` in napari (z,y,x) order.
    """
    from .group import DeviceGroup

    if devices is not None or isinstance(engine, DeviceGroup):
        if stack_shape is not None or pix_begin:
            raise ValueError("stack_shape / pix_begin describe a pre-sharded stack; a device group shards by itself")
        return _fuse_on_group(height_maps, valid_masks, group=engine if isinstance(engine, DeviceGroup) else None,
                              devices=devices, n_clusters=n_clusters, init=init, n_init=n_init, max_iter=max_iter,
                              tol=tol, random_state=random_state, max_abs_height=max_abs_height, detrend=detrend,
                              disparity_scale=disparity_scale, ground_level=ground_level, return_cloud=return_cloud,
                              raster_layout=raster_layout)
    own = engine is None
    eng = engine or Engine(device)
    try:
        # without ground levelling the cloud is final as soon as a slab is unprojected: it is
        # streamed back while the rest of the stack is still being uploaded
        stream = "napari" if (return_cloud and not ground_level) else None
        n = eng.unproject(height_maps, valid_masks, max_abs_height=max_abs_height, detrend=detrend,
                          disparity_scale=disparity_scale, stack_shape=stack_shape, pix_begin=pix_begin,
                          stream_cloud=stream, raster_layout=raster_layout)
        cloud = None
        if stream is not None:
            n, cloud = n
        hn = None
        extra = {"segment_offsets": eng.segment_offsets}
        if ground_level:
            lo, hi, hn = eng.ground_level(True)
            extra["h_min"], extra["h_max"] = lo, hi
        # otherwise its device->host copy is issued now and overlaps the Lloyd loop
        if return_cloud and cloud is None:
            cloud = eng.get_cloud(napari_order=True, wait=False)
        r = _run_kmeans(eng, n_clusters=n_clusters, init=init, n_init=n_init, max_iter=max_iter,
                        tol=tol, random_state=random_state)
        eng.wait()
        return FusionResult(labels=r["labels"], centroids=r["centers"], fused_cloud=cloud,
                            n_iter=r["n_iter"], inertia=r["inertia"], n_points=n,
                            n_refined=r["n_refined"], n_relocations=r["n_relocations"],
                            height_norm=hn, extra=extra)
    finally:
        if own:
            eng.close()


def _fuse_on_group(height_maps, valid_masks, *, group, devices, n_clusters, init, n_init, max_iter, tol,
                   random_state, **unproject_kw) -> FusionResult:
    """``fuse_multiday_kmeans`` on several devices of this process (``group.DeviceGroup``)."""
    import threading

    from .group import DeviceGroup, clone_random_state

    own = group is None
    grp = group or DeviceGroup(devices)
    try:
        if random_state is None:  # every rank must draw the same numbers
            random_state = int(np.random.RandomState().randint(0, 2**31 - 1))
        states = clone_random_state(random_state, grp.world)
        # thread all-gather of the shards' label maps (see _is_same_clustering)
        barrier = threading.Barrier(grp.world)
        slots = [None] * grp.world

        def gather_for(rank):
            def gather(mapping, ok):
                slots[rank] = (mapping, ok)
                barrier.wait()
                got = list(slots)
                barrier.wait()  # nobody overwrites a slot before everybody has read it
                return [g[0] for g in got], [g[1] for g in got]
            return gather

        def run(eng, labels_out):
            r = eng.rank
            return _run_kmeans(eng, n_clusters=n_clusters, init=init, n_init=n_init, max_iter=max_iter, tol=tol,
                               random_state=states[r], labels_out=labels_out,
                               gather_mappings=gather_for(r) if grp.world > 1 else None)

        res, labels, cloud, hn, extra = grp.fuse(height_maps, valid_masks, run_kmeans=run, **unproject_kw)
        return FusionResult(labels=labels, centroids=res["centers"], fused_cloud=cloud, n_iter=res["n_iter"],
                            inertia=res["inertia"], n_points=int(labels.shape[0]), n_refined=res["n_refined"],
                            n_relocations=res["n_relocations"], height_norm=hn, extra=extra)
    finally:
        if own:
            grp.close()


def fuse_height_rasters(paths, **kwargs) -> FusionResult:
    """``fuse_multiday_kmeans`` on the reference's own per-pair height rasters: the
    ``5-out-F.tif`` files its stereo stage writes (disparity.py:213-224; 3-band Float32 GTiff,
    band 0 = -disparity/16, band 2 = final_defined).  The pixels go to the GPU as they are in
    the files; validity (band 2) and the |h| <= 144 test are applied there."""
    from .tiff_io import load_height_rasters

    stack = load_height_rasters(list(paths))
    return fuse_multiday_kmeans(stack, raster_layout="gtiff3", **kwargs)


def kmeans_points(points, *, n_clusters=8, init="k-means++", n_init=1, max_iter=300, tol=1e-4,
                  random_state=None, return_cloud=False, device=0,
                  engine: Optional[Engine] = None) -> FusionResult:
    """Lloyd k-means of an already unprojected ``[N,3]`` (x,y,z) float32 cloud."""
    own = engine is None
    eng = engine or Engine(device)
    try:
        n = eng.set_points(points)
        r = _run_kmeans(eng, n_clusters=n_clusters, init=init, n_init=n_init, max_iter=max_iter,
                        tol=tol, random_state=random_state)
        cloud = eng.get_cloud(napari_order=True) if return_cloud else None
        return FusionResult(labels=r["labels"], centroids=r["centers"], fused_cloud=cloud,
                            n_iter=r["n_iter"], inertia=r["inertia"], n_points=n,
                            n_refined=r["n_refined"], n_relocations=r["n_relocations"])
    finally:
        if own:
            eng.close()


def to_layers(result: FusionResult, prefix: str = "Multi-day") -> List[Layer]:
    """napari layers for a FusionResult, mirroring plugin.py:220-233.

    One Points layer with the fused cloud coloured by cluster (or by the reference's
    'height' property when it was computed) and one with the K centroids.
    """
    if result.fused_cloud is None:
        raise ValueError("result has no fused_cloud (return_cloud=False)")
    props: Dict[str, Any] = {"cluster": result.labels}
    face = "cluster"
    if result.height_norm is not None:
        props["height"] = result.height_norm  # plugin.py:186-188
        face = "height"                       # plugin.py:231: the reference colours by 'height'
    layers: List[Layer] = [(
        result.fused_cloud,
        {
            "name": f"{prefix} 3D Point Cloud",
            "size": 2,
            "properties": props,
            "scale": (1, 1, 1),
            "opacity": 0.8,
            "face_colormap": "turbo",
            "face_color": face,
        },
        "points",
    )]
    c = result.centroids
    layers.append((
        np.stack([c[:, 2], c[:, 1], c[:, 0]], axis=1),
        {"name": f"{prefix} Cluster Centroids", "size": 12, "scale": (1, 1, 1),
         "properties": {"cluster": np.arange(c.shape[0])}, "face_color": "cluster",
         "face_colormap": "turbo", "symbol": "cross"},
        "points",
    ))
    return layers


class MultiDayFusionPlugin:
    """``SatellitePlugin``-shaped wrapper (interface.py:10-47) around the fused path.

    ``run`` never raises: like ``HeightMapExtractor.run`` (plugin.py:236-241) it returns a
    single 100x100 image layer named ``"Error: ..."`` on failure, so the napari worker thread
    (widget.py:116-147) sees the same convention.
    """

    requires_image = False  # plugin.py:30

    def __init__(self, n_clusters=8, device=0, devices=None, log_path=None, **kmeans_kwargs):
        """``log_path``: where the traceback of a failed run is appended, like the reference's
        ``data/TEMP/log.txt`` (plugin.py:48, 236-240); None = only printed.
        ``devices``: several GPUs of this process (see ``fuse_multiday_kmeans``)."""
        self.n_clusters = n_clusters
        self.device = device
        self.devices = devices
        self.log_path = log_path
        self.kmeans_kwargs = kmeans_kwargs

    @property
    def name(self) -> str:
        return "Multi-day 3D Point Cloud (B200 k-means fusion)"

    @property
    def requires_viewer(self) -> bool:
        return False

    def run(self, image, viewer=None, valid_masks=None) -> List[Layer]:
        try:
            res = fuse_multiday_kmeans(image, valid_masks, n_clusters=self.n_clusters, device=self.device,
                                       devices=self.devices, **self.kmeans_kwargs)
            return to_layers(res)
        except Exception as e:  # noqa: BLE001 - reference convention, plugin.py:236-241
            traceback.print_exc()
            if self.log_path is not None:
                try:  # plugin.py:238-239: f.writelines(f"Error: {e}\n{traceback}") into TEMP/log.txt
                    with open(self.log_path, "a") as f:
                        f.writelines(f"Error: {str(e)}\n{traceback.format_exc()}")
                except OSError:
                    pass
            return [(np.ones((100, 100)), {"name": f"Error: {str(e)}"}, "image")]
