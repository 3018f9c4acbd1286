#!/usr/bin/env python
"""Tuning aid: cold-fit time and phases for values of one library option.

    python tools/sweep_opts.py ITERS_PER_LAUNCH 1,2,5,10 c2 c3 c4
"""
import importlib, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
C = importlib.import_module("3d-point-cloud-multiday-imagery_b200._cabi")
import bench  # noqa: E402  (CONFIGS)

opt = getattr(C, "OPT_" + sys.argv[1])
values = [int(v) for v in sys.argv[2].split(",")]
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng = pkg.Engine(0, stream=stream)
for cfg in sys.argv[3:] or ["c2"]:
    D, H, W, k, iters = bench.CONFIGS[cfg]
    hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
    n = eng.unproject(hm)
    del hm
    torch.cuda.empty_cache()
    idx = np.sort(np.random.RandomState(0).choice(n, k, replace=False))
    init = eng.gather_points(idx).astype(np.float64)
    labels = torch.empty(n, dtype=torch.int32, device="cuda")
    iters = min(iters, 20)
    for v in values:
        eng.set_option(opt, v)
        for _ in range(2):
            eng.drop_caches()
            eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
        eng.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        reps = 8
        for _ in range(reps):
            eng.drop_caches()
            r = eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
        e1.record(stream)
        torch.cuda.synchronize()
        ph = eng.profile_phases()
        eng.profile(False)
        print(f"{cfg} {sys.argv[1]}={v:3d}: fit {e0.elapsed_time(e1)/reps:7.3f} ms  build {ph['build'][0]/reps:6.3f}  "
              f"step {1e3*ph['step'][0]/max(1,ph['step'][1]):7.2f} us  final {ph['final'][0]/reps:6.3f}  "
              f"worklist/iter {r['worklist_groups']/r['n_iter']:9.0f} of {r['groups']}", flush=True)
eng.close()
