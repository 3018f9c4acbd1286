#!/usr/bin/env python
"""Tuning aid: cold-fit time, phases and worklist size for different cell shapes of the raster mirror.

    python tools/sweep_cells.py c2 [c3 ...]
"""
import importlib, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
C = importlib.import_module("3d-point-cloud-multiday-imagery_b200._cabi")
import bench  # noqa: E402  (CONFIGS)

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = pkg.Engine(0, stream=stream)
for cfg in sys.argv[1:] or ["c2"]:
    D, H, W, k, iters = bench.CONFIGS[cfg]
    hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
    n = eng.unproject(hm)
    del hm
    torch.cuda.empty_cache()
    idx = np.sort(np.random.RandomState(0).choice(n, k, replace=False))
    init = eng.gather_points(idx).astype(np.float64)
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    iters = min(iters, 20)
    for px, rows in [(0, 0), (16, 1), (16, 2), (16, 3), (16, 4), (8, 1), (8, 2), (8, 3), (8, 4), (8, 6), (8, 8), (8, 12), (8, 16)]:
        eng.set_option(C.OPT_CELL_PX, px)
        eng.set_option(C.OPT_CELL_ROWS, rows)
        for _ in range(2):
            eng.drop_caches()
            eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
        eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        reps = 5
        for _ in range(reps):
            eng.drop_caches()
            r = eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
        e1.record(stream)
        torch.cuda.synchronize()
        ph = eng.profile_phases()
        eng.profile(False)
        print(f"{cfg} cell {px:2d}x{rows:2d}: fit {e0.elapsed_time(e1)/reps:7.3f} ms  build {ph['build'][0]/reps:6.3f}  "
              f"step {1e3*ph['step'][0]/max(1,ph['step'][1]):7.2f} us  final {ph['final'][0]/reps:6.3f}  "
              f"worklist/iter {r['worklist_groups']/r['n_iter']:9.0f} of {r['groups']}", flush=True)
eng.close()
