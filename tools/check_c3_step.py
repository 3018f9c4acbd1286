#!/usr/bin/env python
"""Config-3-shard size (163 M points, k = 64): single step and short fits against the C oracle."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")
from oracle import c_oracle, kmeans_oracle as KO
D, H, W, k = 20, 1024, 8192, 64
eng = pkg.Engine(0)
hm = pkg.make_stack(D, H, W, seed=0, device="cuda")
n = eng.unproject(hm)
del hm
torch.cuda.empty_cache()
cloud = eng.get_cloud(False)
x, y, z = (np.ascontiguousarray(cloud[:, i]) for i in range(3))
idx = np.sort(np.random.RandomState(0).choice(n, k, replace=False))
C = eng.gather_points(idx).astype(np.float64)
for it in range(6):
    eng.drop_caches()
    lab, sums, counts, nref = eng.lloyd_step(C)
    lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
    bad = int((lab != lab_ref).sum())
    print(f"iter {it}: label mismatches {bad}, counts equal {np.array_equal(counts, counts_ref.astype(np.int64))}, "
          f"min count {counts.min()} / {int(counts_ref.min())}, refined {nref}", flush=True)
    cnew = sums_ref / np.maximum(counts_ref, 1)[:, None]
    C = cnew
r = eng.fit(eng.gather_points(idx).astype(np.float64), max_iter=20, tol=0.0, want_labels=False)
print("fit:", r["n_iter"], r["n_relocations"], r["inertia"])
C0 = eng.gather_points(idx).astype(np.float64)
Cs = [C0]
C = C0
for it in range(6):
    lab_ref, sums_ref, counts_ref, _ = c_oracle.lloyd_step_f32soa(x, y, z, C)
    C = sums_ref / np.maximum(counts_ref, 1)[:, None]
    Cs.append(C)
for m in range(1, 7):
    r = eng.fit(C0, max_iter=m, tol=0.0, want_labels=False)
    err = np.abs(r["centers"] - Cs[m]).max()
    print(f"max_iter={m}: n_iter {r['n_iter']} reloc {r['n_relocations']} max |dc| {err:.3e}", flush=True)
