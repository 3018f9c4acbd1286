"""CPU oracle for the height-map -> XYZ unprojection tail -- TEST INFRASTRUCTURE ONLY.

Restates, in float64 numpy and with the same library calls, the tail of
``HeightMapExtractor.run`` in the reference
(``members/rafael/disparity/plugin.py:147-192``).  ``plugin.py`` itself cannot be
imported in this image (it needs osgeo / rasterio / skimage / napari) and the reference
holds no test or fixture for this stage.  **Pinned** instead by
outputs of the reference's own source lines: ``tests/golden/make_unproject_ref.py`` reads
``plugin.py:148-192`` from the reference checkout, ``exec``s the lines verbatim (with
``C.MAX_DISP`` and ``normalise_for_display`` executed from ``constants.py`` / ``utils.py``)
on synthetic int16 disparities and stores every intermediate in
``tests/golden/unproject_ref.npz``; ``tests/test_oracle.py`` requires this restatement to
reproduce them bit for bit (valid mask, P, centre, normal, height_rel, h_min, h_max,
h_norm, points_coords).

The multi-day merge (one cloud for all days, day-major order) has no reference code
(SURVEY.md F1); it is ``np.concatenate`` over what the reference's ``for pair`` loop
(``plugin.py:106``) produces per pair.
"""
from __future__ import annotations

import numpy as np

MAX_DISP = 288  # members/rafael/disparity/constants.py:54-57


def height_from_disparity(disparity: np.ndarray) -> np.ndarray:
    """plugin.py:148 -- OpenCV stores disparity as int16 fixed point, 1/16 px."""
    return -disparity.astype(float) / 16.0


def valid_mask(height_map, validity_mask=None, limit=MAX_DISP / 2):
    """plugin.py:151-152."""
    hm = np.asarray(height_map)
    with np.errstate(invalid="ignore"):
        m = np.isfinite(hm) & (np.abs(hm) <= limit)
    if validity_mask is not None:
        m &= np.asarray(validity_mask).astype(bool)
    return m


def unproject_day(height_map, validity_mask=None, limit=MAX_DISP / 2, detrend=False):
    """One day / pair.  Returns (P[N,3] = x,y,z float64, plane) following plugin.py:157-171.

    ``detrend=True`` applies the SVD plane fit of plugin.py:161-171 and replaces z by the
    signed distance to the plane; x, y stay pixel indices (plugin.py:192 uses the raw
    ``y_indices, x_indices``).
    """
    hm = np.asarray(height_map, dtype=np.float64)
    m = valid_mask(hm, validity_mask, limit)
    y_idx, x_idx = np.where(m)  # row-major order, plugin.py:157
    z = hm[m]
    P = np.stack([x_idx, y_idx, z], axis=1).astype(np.float64)  # plugin.py:160
    plane = None
    if detrend and P.shape[0] >= 3:
        center = np.mean(P, axis=0)  # plugin.py:161
        Pc = P - center
        _, _, Vh = np.linalg.svd(Pc, full_matrices=False)  # plugin.py:164
        normal = Vh[2]
        if np.dot(normal, np.array([0, 0, 1])) < 0:  # plugin.py:167-168
            normal = -normal
        P = P.copy()
        P[:, 2] = np.dot(Pc, normal)  # plugin.py:171
        plane = (center, normal)
    return P, plane


def ground_level(z_values):
    """plugin.py:181-192: percentile normalisation.  Returns (z - h_min, h_norm)."""
    h_min = np.percentile(z_values, 2)
    h_max = np.percentile(z_values, 98)
    div = h_max - h_min + 1e-6
    h_norm = np.clip((z_values - h_min) / div, 0, 1)
    return z_values - h_min, h_norm


def unproject_stack(height_maps, validity_masks=None, limit=MAX_DISP / 2, detrend=False):
    """All days merged, day-major then row-major.  Returns P[N,3] (x,y,z) float64."""
    hm = np.asarray(height_maps)
    if hm.ndim == 2:
        hm = hm[None]
        if validity_masks is not None and np.asarray(validity_masks).ndim == 2:
            validity_masks = np.asarray(validity_masks)[None]
    out = []
    for d in range(hm.shape[0]):
        vm = None if validity_masks is None else np.asarray(validity_masks)[d]
        P, _ = unproject_day(hm[d], vm, limit, detrend)
        out.append(P)
    if not out:
        return np.zeros((0, 3))
    return np.concatenate(out, axis=0)


def reference_tail_stack(height_maps, validity_masks=None, limit=MAX_DISP / 2, detrend=True):
    """The reference's whole per-pair tail (plugin.py:148-192) applied to every day, merged.

    Per day: unprojection, plane detrend (the reference always applies it), percentile
    ground-levelling.  Returns (P[N,3] = x,y,z with z levelled, h_norm[N], h_min[D], h_max[D],
    offsets[D+1]).
    """
    hm = np.asarray(height_maps)
    if hm.ndim == 2:
        hm = hm[None]
        if validity_masks is not None and np.asarray(validity_masks).ndim == 2:
            validity_masks = np.asarray(validity_masks)[None]
    pts, hns, los, his, off = [], [], [], [], [0]
    for d in range(hm.shape[0]):
        vm = None if validity_masks is None else np.asarray(validity_masks)[d]
        P, _ = unproject_day(hm[d], vm, limit, detrend)
        if P.shape[0]:
            z = P[:, 2]
            h_min = np.percentile(z, 2)   # plugin.py:181
            h_max = np.percentile(z, 98)  # plugin.py:182
            z0, hn = ground_level(z)
            P = P.copy()
            P[:, 2] = z0
        else:
            h_min = h_max = np.nan
            hn = np.zeros(0)
        pts.append(P)
        hns.append(hn)
        los.append(h_min)
        his.append(h_max)
        off.append(off[-1] + P.shape[0])
    return (np.concatenate(pts, axis=0), np.concatenate(hns), np.array(los), np.array(his),
            np.array(off, dtype=np.int64))


def to_napari_points(P):
    """plugin.py:192: ``np.stack([z, y, x], axis=1)`` -- napari axis order."""
    return np.stack([P[:, 2], P[:, 1], P[:, 0]], axis=1)
