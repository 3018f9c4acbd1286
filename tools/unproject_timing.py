#!/usr/bin/env python
"""Kernel time of the unprojection on device-resident rasters (CUDA-event spans of the library's own
profiling hooks), for one or more builds of the library:

    python tools/unproject_timing.py [c2] [lib.so ...]
"""
import importlib, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2:  # one child process per library (the library is bound at import time)
    for lib in sys.argv[2:]:
        env = dict(os.environ)
        if lib != "default":
            env["MDKM_LIB"] = os.path.join(ROOT, lib)
        subprocess.run([sys.executable, __file__, sys.argv[1]], env=env)
    sys.exit(0)

import torch
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
import bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
D, H, W, k, iters = bench.CONFIGS[cfg]
eng = pkg.Engine(0)
hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
for detrend in (False, True):
    n = eng.unproject(hm, detrend=detrend)
    eng.profile(True)
    res = []
    for _ in range(3):
        for _ in range(5):
            n = eng.unproject(hm, detrend=detrend)
        ms, px = eng.profile_phases()["unproject"]
        res.append(ms / 5)
    eng.profile(False)
    best = min(res)
    print(f"{os.environ.get('MDKM_LIB', 'default')} {cfg} detrend={detrend}: unproject kernel "
          + " ".join(f"{r:.4f}" for r in res) + f" ms per call, best {16.0 * D * H * W / best / 1e6:.0f} GB/s (n={n})",
          flush=True)
eng.close()
