#!/usr/bin/env python
"""In-kernel phase stamps of the last Lloyd iteration of a fit (diagnostic build, -DMDKM_TIMING).

Builds tools/_build/libmdkm_timing.so when it is missing or stale (here, on the CPU box: nvcc
cross-compiles) and, on a GPU, runs cold fits of the given configs with it; the library prints
one "[mdkm timing]" line per fit to stderr.

    python tools/phase_timing.py build
    python tools/phase_timing.py c2 c3
"""
import importlib, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "3d-point-cloud-multiday-imagery_b200")
OUT = os.path.join(ROOT, "tools", "_build", "libmdkm_timing.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = os.path.join(PKG, "csrc", "mdkm.cu")
    newest = max(os.path.getmtime(os.path.join(PKG, "csrc", f)) for f in os.listdir(os.path.join(PKG, "csrc")))
    if os.path.exists(OUT) and os.path.getmtime(OUT) > newest:
        return
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "--shared", "-Xcompiler", "-fPIC", "-DMDKM_TIMING", "-o", OUT, src, "-ldl"])


if __name__ == "__main__":
    build()
    if sys.argv[1:] == ["build"]:
        sys.exit(0)
    os.environ["MDKM_LIB"] = OUT
    import numpy as np
    import torch
    pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
    import bench
    eng = pkg.Engine(0)
    for cfg in sys.argv[1:] or ["c2"]:
        D, H, W, k, iters = bench.CONFIGS[cfg]
        hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
        n = eng.unproject(hm)
        del hm
        torch.cuda.empty_cache()
        init = eng.gather_points(np.sort(np.random.RandomState(0).choice(n, k, replace=False))).astype(np.float64)
        for it in (3, 10, min(iters, 20)):
            print(f"{cfg}: fit of {it} iterations", file=sys.stderr, flush=True)
            for _ in range(3):
                eng.fit(init, max_iter=it, tol=0.0, want_labels=False)
    eng.close()
