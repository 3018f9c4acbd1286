#!/usr/bin/env python
"""Multi-GPU invariance check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_parity.py

The sharded fit (fused NVLink exchange, and the NCCL exchange) must give the SAME centroids
(bitwise), n_iter and labels as the single-GPU fit of the whole stack, because the exchanged
sums are integers.  Also times one Lloyd iteration on both exchange paths.
"""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3d-point-cloud-multiday-imagery_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    D, H, W = 2 * world, 600, 800
    hm = pkg.make_stack(D, H, W, seed=3).numpy()  # identical on every rank (CPU generator)
    failures = []

    def check(name, cond):
        if not cond:
            failures.append(name)
        if rank == 0:
            print(("ok   " if cond else "FAIL ") + name, flush=True)

    # single-GPU reference of the whole stack (every rank computes it on its own GPU)
    single = pkg.Engine(local)
    n_all = single.unproject(hm)
    cases = [("k16_tol0", 16, 25, 0.0), ("k64_tol", 64, 300, 1e-4), ("k300_tol0", 300, 8, 0.0)]
    inits, refs = {}, {}
    for name, k, it, tol in cases:
        inits[name] = pkg.init_from_points(single.get_cloud(False), k, 5)
        refs[name] = single.fit(inits[name], max_iter=it, tol=tol)
    # a case that needs empty-cluster relocation: two far-away initial centroids
    bad = inits["k16_tol0"].copy()
    bad[3] = [1e5, 1e5, 1e5]
    bad[9] = [-1e5, 3e4, 50]
    inits["reloc"] = bad
    cases.append(("reloc", 16, 12, 0.0))
    refs["reloc"] = single.fit(bad, max_iter=12, tol=0.0)
    check("single-GPU relocation happened", refs["reloc"]["n_relocations"] >= 1)
    off = single.segment_offsets
    kpp_ref = {(k, seed): single.kmeans_plusplus(k, np.random.RandomState(seed)) for k, seed in ((8, 0), (40, 1))}
    rand_ref = single.gather_points(np.array([0, 5, n_all - 1, n_all // 2], dtype=np.int64))
    full_ref = pkg.fuse_multiday_kmeans(hm, n_clusters=6, init="k-means++", n_init=3, random_state=7, max_iter=15,
                                        tol=0.0, engine=single)
    full_ref_labels = full_ref.labels.copy()
    single.close()

    # sharded: rank r owns days [2r, 2r+2)
    b, e = pkg.shard_range(D, rank, world)
    for p2p in (True, False):
        eng = pkg.Engine(local)
        pkg.init_engine_comm(eng, rank, world, p2p=p2p)
        check(f"p2p={p2p}: exchange path is {'fused NVLink' if p2p else 'NCCL'}", eng.p2p == p2p)
        n_loc = eng.unproject(hm[b:e].reshape(-1), stack_shape=(D, H, W), pix_begin=b * H * W)
        check(f"p2p={p2p}: shard sizes", n_loc == int(off[e] - off[b]))
        for name, k, it, tol in cases:
            r = eng.fit(inits[name], max_iter=it, tol=tol)
            ref = refs[name]
            tag = f"p2p={p2p} {name}: "
            check(tag + f"n_iter {r['n_iter']} == {ref['n_iter']}", r["n_iter"] == ref["n_iter"])
            check(tag + "centroids bitwise", r["centers"].tobytes() == ref["centers"].tobytes())
            check(tag + "labels", np.array_equal(r["labels"], ref["labels"][off[b]:off[e]]))
            check(tag + "inertia", abs(r["inertia"] - ref["inertia"]) <= 1e-12 * ref["inertia"])
            check(tag + "relocations", r["n_relocations"] == ref["n_relocations"])
        # sharded k-means++ / global gather / the public API with restarts
        for (k, seed), (c_ref, i_ref) in kpp_ref.items():
            c, i = eng.kmeans_plusplus(k, np.random.RandomState(seed))
            check(f"p2p={p2p} kmeans++ k={k}: indices", np.array_equal(i, i_ref))
            check(f"p2p={p2p} kmeans++ k={k}: centres", np.array_equal(c, c_ref))
        got = eng.gather_points(np.array([0, 5, n_all - 1, n_all // 2], dtype=np.int64))
        check(f"p2p={p2p} global gather", np.array_equal(got, rand_ref))
        res = pkg.fuse_multiday_kmeans(hm[b:e].reshape(-1), n_clusters=6, init="k-means++", n_init=3, random_state=7,
                                       max_iter=15, tol=0.0, engine=eng, stack_shape=(D, H, W), pix_begin=b * H * W)
        check(f"p2p={p2p} API k-means++ n_init=3: centroids", res.centroids.tobytes() == full_ref.centroids.tobytes())
        check(f"p2p={p2p} API k-means++ n_init=3: labels", np.array_equal(res.labels, full_ref_labels[off[b]:off[e]]))
        check(f"p2p={p2p} API k-means++ n_init=3: n_iter", res.n_iter == full_ref.n_iter)
        # time per Lloyd iteration (k=16, 40 iterations, no convergence)
        name, k = "k16_tol0", 16
        for _ in range(3):
            eng.fit(inits[name], max_iter=40, tol=0.0, want_labels=False)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        its = 0
        for _ in range(10):
            its += eng.fit(inits[name], max_iter=40, tol=0.0, want_labels=False)["n_iter"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            print(f"p2p={p2p}: {1e6 * dt / its:.1f} us per Lloyd iteration ({n_loc} pts/rank, k={k}, {world} ranks)",
                  flush=True)
        eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        print(f"rank {rank}: FAILED {failures}", flush=True)
        sys.exit(1)
    if rank == 0:
        print("multi-GPU parity: all checks passed", flush=True)


if __name__ == "__main__":
    main()
