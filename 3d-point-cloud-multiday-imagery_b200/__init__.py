"""B200-native multi-day height-map -> XYZ -> Lloyd k-means fusion.

Drop-in for the one hot path of rafael-alani/3d-point-cloud-multiday-imagery described in
SURVEY.md section 8.  The heavy lifting is in ``libmdkm.so`` (hand-written sm_100a CUDA behind
the C ABI of ``include/mdkm.h``); there is no CPU fallback.
"""
from .api import (FusionResult, Layer, MultiDayFusionPlugin, fuse_height_rasters, fuse_multiday_kmeans,
                  kmeans_points, to_layers)
from .build import build_library
from .dist import init_engine_comm, shard_range
from .engine import Engine
from .group import DeviceGroup
from .synth import init_from_points, make_stack, make_stack_range

__all__ = [
    "DeviceGroup", "Engine", "FusionResult", "Layer", "MultiDayFusionPlugin", "build_library",
    "fuse_height_rasters", "fuse_multiday_kmeans", "init_engine_comm", "init_from_points", "kmeans_points",
    "make_stack", "make_stack_range", "shard_range", "to_layers",
]
