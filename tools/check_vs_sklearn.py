#!/usr/bin/env python
"""Full-fit parity against the real scikit-learn on mid-size stacks (diagnostic)."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
from oracle import sklearn_ref, kmeans_oracle as KO, unproject_oracle as UO

eng = pkg.Engine(0)
for (D, H, W, k, iters) in [(4, 256, 2048, 64, 20), (20, 256, 2048, 64, 20), (6, 512, 512, 300, 10), (6, 1024, 1024, 1024, 8)]:
    hm = pkg.make_stack(D, H, W, seed=1).numpy()
    P = UO.unproject_stack(hm)
    n = eng.unproject(hm)
    init = pkg.init_from_points(P.astype(np.float32), k, 1)
    ref = sklearn_ref.fit(P, init, max_iter=iters, tol=0.0)
    for rep in range(3):
        eng.drop_caches()
        r = eng.fit(init, max_iter=iters, tol=0.0)
        cmp = KO.compare_labels(P, ref["centers"], ref["labels"], r["labels"])
        err = KO.centroid_rel_err(ref["centers"], r["centers"], P.std(axis=0))
        print(f"D={D} {H}x{W} k={k} n={n}: n_iter {r['n_iter']} vs {ref['n_iter']}, reloc {r['n_relocations']}, "
              f"labels {cmp}, centroid err {err:.2e}, inertia rel {abs(r['inertia']-ref['inertia'])/ref['inertia']:.2e}", flush=True)
