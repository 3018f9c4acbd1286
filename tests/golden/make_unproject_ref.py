"""Golden vectors for the unprojection tail, produced by the REFERENCE'S OWN SOURCE LINES.

``members/rafael/disparity/plugin.py`` cannot be imported in this image (its module imports
need osgeo / rasterio / skimage / napari), but the tail of ``HeightMapExtractor.run`` that this
project replaces (``plugin.py:147-192``: ``height_map = -disparity/16`` ... ``points_coords =
np.stack([z, y, x])``) is plain numpy.  This script reads those lines from the reference file
where it lies (nothing is copied into the repository), dedents them, and ``exec``s them
verbatim with

  * ``disparity`` / ``validity_mask``  the synthetic inputs saved in the fixture,
  * ``C.MAX_DISP``                     executed from ``constants.py:54-57``,
  * ``normalise_for_display``          executed from ``utils.py:9-14`` (display only; its result
                                       never reaches the points),
  * ``layers``, ``PREFIX``             an empty list / the plugin's string,

and stores what the reference code left in its local variables: ``valid_mask``, ``P``,
``center``, ``normal``, ``height_rel``, ``h_min``, ``h_max``, ``h_norm``, ``points_coords``.

Run in the build container (numpy 2.3.5), where /root/reference exists:

    python tests/golden/make_unproject_ref.py      ->  tests/golden/unproject_ref.npz

The fixture pins oracle/unproject_oracle.py (tests/test_oracle.py) and, through it and directly,
the CUDA path (tests/test_gpu_parity.py) to the reference instead of to a restatement.
"""
from __future__ import annotations

import os
import re
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MDKM_REFERENCE", "/root/reference")
PLUGIN = os.path.join(REF, "members/rafael/disparity/plugin.py")
CONSTANTS = os.path.join(REF, "members/rafael/disparity/constants.py")
UTILS = os.path.join(REF, "members/rafael/disparity/utils.py")

FIRST = "height_map = -disparity.astype(float) / 16.0"   # plugin.py:148
LAST = "points_coords = np.stack([z_values, y_indices, x_indices], axis=1)"  # plugin.py:192


def reference_tail_source():
    """The literal lines of plugin.py from FIRST to LAST (inclusive), dedented."""
    lines = open(PLUGIN).read().split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip().startswith(FIRST))
    i1 = next(i for i, l in enumerate(lines) if l.strip().startswith(LAST))
    assert (i0 + 1, i1 + 1) == (148, 192), (i0 + 1, i1 + 1)  # the lines SURVEY.md section 8(a) cites
    return textwrap.dedent("\n".join(lines[i0:i1 + 1])), (i0 + 1, i1 + 1)


def reference_namespace():
    consts = open(CONSTANTS).read().split("\n")
    j0 = next(i for i, l in enumerate(consts) if re.match(r"MAX_DISP\s*=", l))
    cns = {}
    exec("\n".join(consts[j0:j0 + 4]), cns)  # constants.py:54-57
    C = types.SimpleNamespace(MAX_DISP=cns["MAX_DISP"])
    utils = open(UTILS).read().split("\n")
    u0 = next(i for i, l in enumerate(utils) if l.startswith("def normalise_for_display"))
    u1 = next(i for i in range(u0 + 1, len(utils)) if utils[i].startswith("def "))
    uns = {"np": np}
    exec("\n".join(utils[u0:u1]), uns)
    return {"np": np, "C": C, "normalise_for_display": uns["normalise_for_display"],
            "PREFIX": "[Multi-day 3D Point Cloud]"}


def run_reference_tail(disparity, validity_mask):
    """Executes plugin.py:148-192 on one pair's (disparity int16 [H,W], validity bool [H,W])."""
    src, _ = reference_tail_source()
    ns = reference_namespace()
    ns.update(disparity=disparity, validity_mask=validity_mask, layers=[])
    exec(compile(src, PLUGIN, "exec"), ns)
    keep = ("valid_mask", "P", "center", "normal", "height_rel", "h_min", "h_max", "h_norm", "points_coords")
    return {k: np.asarray(ns[k]) for k in keep}


def synth_pair(rs, H, W, tilt=(0.3, -0.2), frac_invalid=0.12, frac_sentinel=0.05, n_boxes=4, flip=False):
    """int16 fixed-point disparity (OpenCV: 16 * pixels) of a tilted ground with a few plateaus,
    sentinels outside |h| <= 144 and a ``final_defined``-style validity mask."""
    y, x = np.mgrid[0:H, 0:W]
    h = tilt[0] * x + tilt[1] * y + rs.normal(0, 0.4, (H, W))
    for _ in range(n_boxes):
        x0, y0 = rs.randint(0, W), rs.randint(0, H)
        h[y0:y0 + rs.randint(3, 12), x0:x0 + rs.randint(3, 12)] += rs.uniform(4, 30)
    if flip:
        h = -h + 40.0
    disp = np.rint(-16.0 * h).astype(np.int64)
    sent = rs.rand(H, W) < frac_sentinel
    disp[sent] = rs.choice([-16 * 145, 16 * 200, 32767, -32768, -16 * 144 - 1], size=int(sent.sum()))
    edge = rs.rand(H, W) < 0.01
    disp[edge] = rs.choice([16 * 144, -16 * 144], size=int(edge.sum()))  # exactly on the limit: valid
    disp = np.clip(disp, -32768, 32767).astype(np.int16)
    mask = rs.rand(H, W) > frac_invalid
    return disp, mask


def main():
    rs = np.random.RandomState(20260118)
    out = {}
    cases = [
        ("a", 64, 96, dict()),
        ("b", 37, 53, dict(tilt=(-0.8, 0.5), frac_invalid=0.5, frac_sentinel=0.2)),   # ragged, sparse
        ("c", 128, 160, dict(tilt=(0.05, 0.02), n_boxes=10, flip=True)),
        ("d", 5, 260, dict(tilt=(0.5, 2.0), frac_invalid=0.02, n_boxes=1)),            # wide and short
    ]
    for name, H, W, kw in cases:
        disp, mask = synth_pair(rs, H, W, **kw)
        r = run_reference_tail(disp, mask)
        out[f"{name}_disparity"] = disp
        out[f"{name}_validity_mask"] = mask
        for k, v in r.items():
            out[f"{name}_{k}"] = v
        print(name, (H, W), "valid", int(r["valid_mask"].sum()), "h_min", float(r["h_min"]), "h_max", float(r["h_max"]),
              "normal", r["normal"])
    _, rng = reference_tail_source()
    out["cases"] = np.array([c[0] for c in cases])
    out["source_lines"] = np.array(rng)
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(HERE, "unproject_ref.npz"), **out)
    print("wrote unproject_ref.npz from", PLUGIN, "lines", rng)


if __name__ == "__main__":
    sys.exit(main())
