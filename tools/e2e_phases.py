#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end call (diagnostic, not a bench number)."""
import importlib, os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")

D, H, W, k, iters = 10, 2048, 2048, 16, 20
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
eng = pkg.Engine(0, stream=stream, pinned_results=True)
hm = pkg.make_stack(D, H, W, seed=0, device="cuda:0")
hm_host = torch.empty(hm.shape, dtype=torch.float32, pin_memory=True)
hm_host.copy_(hm)
torch.cuda.synchronize()

def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3

dev = torch.empty(hm.numel(), dtype=torch.float32, device="cuda:0")
print("H2D 168MB pinned: %.2f ms" % t(lambda: dev.copy_(hm_host.view(-1), non_blocking=True)))
big = torch.empty(123 * 1000 * 1000, dtype=torch.float32, device="cuda:0")
big_h = torch.empty(big.numel(), dtype=torch.float32, pin_memory=True)
ms = t(lambda: big_h.copy_(big, non_blocking=True))
print("D2H 492MB pinned: %.2f ms (%.1f GB/s)" % (ms, big.numel() * 4 / ms / 1e6))
s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2):
        dev.copy_(hm_host.view(-1), non_blocking=True)
    big_h.copy_(big, non_blocking=True)
print("H2D 168MB || D2H 492MB: %.2f ms" % t(both))

n = eng.unproject(hm_host)
init = pkg.init_from_points(eng.get_cloud(False)[:100000].copy(), k, 0)
print("unproject(host) no stream: %.2f ms" % t(lambda: eng.unproject(hm_host)))
def u_s():
    r = eng.unproject(hm_host, stream_cloud="napari"); return r
print("unproject(host) + streamed cloud, returns: %.2f ms" % t(u_s))
def u_sw():
    r = eng.unproject(hm_host, stream_cloud="napari"); eng.wait(); return r
print("unproject(host) + streamed cloud + wait: %.2f ms" % t(u_sw))
print("unproject(device): %.2f ms" % t(lambda: eng.unproject(hm)))
print("get_cloud sync: %.2f ms" % t(lambda: eng.get_cloud(True)))
print("fit no labels: %.2f ms" % t(lambda: eng.fit(init, max_iter=iters, tol=0.0, want_labels=False)))
print("fit + labels: %.2f ms" % t(lambda: eng.fit(init, max_iter=iters, tol=0.0)))
def e2e():
    return pkg.fuse_multiday_kmeans(hm_host, n_clusters=k, init=init, max_iter=iters, tol=0.0, engine=eng)
print("e2e: %.2f ms" % t(e2e))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); e2e(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
