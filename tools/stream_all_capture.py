#!/usr/bin/env python
"""Cold fits of one config with settling disabled (MDKM_OPT_SETTLE_GROUPS = 0: every point is fetched
and assigned in every iteration) -- the workload of bench.py's `roofline.stream_all`, on its own so
that ncu can capture a steady-state iteration of it:

    ncu --set full --clock-control none --cache-control none -k regex:lloyd_step -s 25 -c 1 \
        -o gpurun_out/prof_stream_all python tools/stream_all_capture.py c2
"""
import importlib, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
C = importlib.import_module("3d-point-cloud-multiday-imagery_b200._cabi")
import bench  # noqa: E402  (CONFIGS)

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
D, H, W, k, iters = bench.CONFIGS[cfg]
eng = pkg.Engine(0)
hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
n = eng.unproject(hm)
del hm
torch.cuda.empty_cache()
init = eng.gather_points(np.sort(np.random.RandomState(0).choice(n, k, replace=False))).astype(np.float64)
labels = torch.empty(n, dtype=torch.int32, device="cuda")
eng.set_option(C.OPT_SETTLE_GROUPS, 0)
iters = min(iters, 20)
for _ in range(2):
    eng.drop_caches()
    eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
eng.profile(True)
for _ in range(3):
    eng.drop_caches()
    r = eng.fit(init, max_iter=iters, tol=0.0, labels_out=labels)
ph = eng.profile_phases()
us = 1e3 * ph["step"][0] / max(1, ph["step"][1])
print(f"{cfg} stream-all: {us:.1f} us per iteration, {16.0 * n / us / 1e3:.0f} GB/s of algorithmic bytes, "
      f"worklist {r['worklist_groups'] / r['n_iter']:.0f} of {r['groups']} groups per iteration")
eng.close()
