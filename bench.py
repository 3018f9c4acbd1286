#!/usr/bin/env python
"""bench.py -- k-means points*iterations/s of the multi-day fusion path on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

A "step" is one complete Lloyd fit (BASELINE.json configs[1]: a 10-day 2048x2048 height-map
stack per GPU, ~41.9 M points, k = 16, 20 iterations, tol = 0) over points already resident
in HBM; `value` = points * iterations of all ranks / device time (CUDA events, max over
ranks).  `e2e` is the same metric through the public Python API with pinned HOST rasters in
and host labels / centroids / fused cloud out.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3d-point-cloud-multiday-imagery_b200"

METRIC = "kmeans_points_iters_per_sec"
UNIT = "points*iters/s"
ALGO_BYTES_PER_POINT_ITER = 16.0  # 12 B xyz read + 4 B label write

CONFIGS = {
    # name: (days per GPU, H, W, k, iters)
    "c1": (3, 512, 512, 8, 20),
    "c2": (10, 2048, 2048, 16, 20),
    # c3 = 20-day 8192x8192 stack (1.34 G pts) over 8 GPUs: 2.5 days = 20 x 1024 rows bands per GPU
    "c3": (20, 1024, 8192, 64, 20),
    "c4": (12, 4096, 4096, 1024, 10),
    "c5": (30, 4096, 4096, 32, 300),
}
# relative tolerance of the convergence test (sklearn's tol); config 5 is the time-to-solution run
TOL = {"c5": 1e-4}


def workload_name(cfg, n_gpus):
    D, H, W, k, it = CONFIGS[cfg]
    tol = TOL.get(cfg, 0.0)
    iters = f"{it} Lloyd iters, tol=0" if tol == 0 else f"max_iter={it}, tol={tol:g} (time to solution)"
    return (f"{cfg}: synthetic {D}-day {H}x{W} height-map stack per GPU (~{D*H*W*0.975/1e6:.1f}M pts/GPU), "
            f"k={k}, {iters}, {n_gpus} GPU(s)")


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_ev.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_traffic(cfg):
    """DRAM bytes per Lloyd iteration from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))[cfg]
        return float(t["dram_bytes_per_iteration"]), t.get("source", "")
    except Exception:
        return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# reference arm: scikit-learn, the implementation the reference calls (core.py:227-228)
# ---------------------------------------------------------------------------------------
def cpu_sample(pkg, cfg, seed=0):
    """A bounded sample of the workload: ONE day of the stack, same generator, same k / iters."""
    from oracle import unproject_oracle as UO

    D, H, W, k, iters = CONFIGS[cfg]
    hm = pkg.make_stack(1, H, W, seed=seed).numpy()
    P = UO.unproject_stack(hm)
    init = pkg.init_from_points(P.astype(np.float32), k, seed)
    return P, init, k, iters, f"1 of {D} days of the {H}x{W} stack ({P.shape[0]} pts), k={k}, {iters} iters, float64"


def time_cpu_reference(P, init, iters, repeats, tol=0.0):
    """Returns (best pts*it/s, kind, cores).  sklearn when importable, else the C port."""
    from oracle import sklearn_ref

    if sklearn_ref.available():
        best = 0.0
        for _ in range(repeats):
            r = sklearn_ref.fit(P, init, max_iter=iters, tol=tol)
            best = max(best, P.shape[0] * r["n_iter"] / r["wall_s"])
        return best, "reference", sklearn_ref.n_threads()
    from oracle import c_oracle, kmeans_oracle as KO

    P32 = P.astype(np.float32)
    x, y, z = (np.ascontiguousarray(P32[:, i]) for i in range(3))
    best = 0.0
    for _ in range(repeats):
        c = init.copy()
        t0 = time.perf_counter()
        for _it in range(iters):
            _, sums, counts, _ = c_oracle.lloyd_step_f32soa(x, y, z, c)
            KO.average_centers(sums, counts)
            c = sums
        best = max(best, P.shape[0] * iters / (time.perf_counter() - t0))
    return best, "port", c_oracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = importlib.import_module(PKG)
    P, init, k, iters, sample = cpu_sample(pkg, args.config)
    from oracle import sklearn_ref

    kind = "reference" if sklearn_ref.available() else "port"
    for _ in range(args.warmup):
        time_cpu_reference(P, init, iters, 1, TOL.get(args.config, 0.0))
    t0 = time.perf_counter()
    total = 0.0
    cores = 1
    for _ in range(args.steps):
        v, kind, cores = time_cpu_reference(P, init, iters, 1, TOL.get(args.config, 0.0))
        total += P.shape[0] * iters / v
    wall = time.perf_counter() - t0
    value = P.shape[0] * iters * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.gpus), "sample": sample,
                   "implementation": "sklearn.cluster.KMeans(algorithm='lloyd', n_init=1, init=array) float64"
                   if kind == "reference" else "oracle/lloyd_oracle.c (C port, OpenMP)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------
def run_mine(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module(PKG)

    D, H, W, k, iters = CONFIGS[args.config]
    stream = torch.cuda.Stream()
    eng = pkg.Engine(local, stream=stream, pinned_results=True)
    pkg.init_engine_comm(eng, rank, world)

    # synthetic stack of this rank's days, generated on the device (seed differs per rank)
    hm = pkg.make_stack(D, H, W, seed=args.seed + rank, device=f"cuda:{local}")
    torch.cuda.synchronize()
    hm_host = torch.empty(hm.shape, dtype=torch.float32, pin_memory=True)
    hm_host.copy_(hm)
    torch.cuda.synchronize()
    pix0 = rank * D * H * W
    with torch.cuda.stream(stream):
        n_local = eng.unproject(hm, stack_shape=(D * world, H, W), pix_begin=pix0)
    del hm
    torch.cuda.empty_cache()

    # identical initial centroids on every rank: k points of the global cloud (gather_points is
    # a collective with global indices, so every rank makes the same call)
    n_tot_t = torch.tensor([n_local], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(n_tot_t)
    n_total = int(n_tot_t.item())
    idx = np.sort(np.random.RandomState(args.seed).choice(n_total, k, replace=False))
    init_np = eng.gather_points(idx).astype(np.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_fit():
        # every timed fit starts cold: the per-cloud acceleration structures (tile-ordered mirror,
        # group summaries) are rebuilt inside the timed region, nothing is carried over
        if not args.keep_caches:
            eng.drop_caches()
        return eng.fit(init_np, max_iter=iters, tol=TOL.get(args.config, 0.0), want_labels=False)

    # ---- kernel-resident number: K fits over resident points --------------------------------
    for _ in range(args.warmup):
        r = one_fit()
    eng.profile_read()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n_iter_sum = 0
    for _ in range(args.steps):
        r = one_fit()
        n_iter_sum += r["n_iter"]
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    _, _, launches = eng.profile_read()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = n_total * n_iter_sum / (ms_total * 1e-3)

    # ---- dominant kernel (assignment + accumulate) timed alone with CUDA events ---------------
    eng.profile(True)
    for _ in range(2):
        one_fit()
    step_ms, n_steps, _ = eng.profile_read()
    eng.profile(False)
    peak, peak_src = measured_peak()
    step_avg_ms = step_ms / max(n_steps, 1)
    achieved = ALGO_BYTES_PER_POINT_ITER * n_local / (step_avg_ms * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(args.config)
    # the DRAM rate the kernel actually sustains: ncu's bytes per launch over the live duration
    dram_gbs = (traffic / (step_avg_ms * 1e-3) / 1e9) if traffic else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "dram_gbs": dram_gbs, "dram_frac": (dram_gbs / peak) if dram_gbs else None,
                "kernel": "lloyd_step_kernel = one Lloyd iteration (classification pass, grid barrier, per-point pass, fused update)",
                "avg_launch_ms": step_avg_ms,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_POINT_ITER * n_local, "peak_source": peak_src,
                "note": "achieved = 16 B x points / measured iteration time; the kernels move fewer bytes than "
                        "that (cached group summaries), so frac can exceed what a 16 B/point stream allows",
                "fma_bound_points_iters_per_s": 148 * 128 * 1.965e9 / (3 * k)}

    # ---- end to end through the public API: pinned host rasters in, host results out ------------
    def one_e2e():
        return pkg.fuse_multiday_kmeans(hm_host, n_clusters=k, init=init_np, max_iter=iters,
                                        tol=TOL.get(args.config, 0.0),
                                        engine=eng, stack_shape=(D * world, H, W), pix_begin=pix0)

    e2e = None
    if not args.no_e2e:
        for _ in range(max(1, args.warmup - 1)):
            res = one_e2e()
        barrier()
        t0 = time.perf_counter()
        its = 0
        for _ in range(args.steps):
            res = one_e2e()
            its += res.n_iter
        barrier()
        wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * its / float(wall.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(hm_host.numel() * 4 + init_np.nbytes),
               "d2h_bytes_per_step": int(res.labels.nbytes + res.centroids.nbytes + res.fused_cloud.nbytes),
               "ms_per_step": 1e3 * float(wall.item()) / args.steps,
               "returns": "labels int32[N], centroids f64[K,3], fused cloud f32[N,3]"}

    # ---- CPU baseline (rank 0, N = 1 only): the reference's sklearn path on a bounded sample ---
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        P, init_c, _, it_c, sample = cpu_sample(pkg, args.config, args.seed)
        v, kind, cores = time_cpu_reference(P, init_c, it_c, 2, TOL.get(args.config, 0.0))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, world), "points_total": n_total,
                       "iters_per_step": n_iter_sum / args.steps, "l2_policy": "inputs_exceed_l2 "
                       f"({n_local * 12 / 1e6:.0f} MB of xyz per GPU vs 126 MB L2)",
                       "per_cloud_structures": "kept across fits" if args.keep_caches else
                       "rebuilt inside every timed fit (mirror + group summaries)",
                       "exchange": ("none" if world == 1 else
                                    "in-kernel NVLink peer exchange of K*4+8 int64 per iteration (CUDA IPC), no NCCL"
                                    if eng.p2p else "ncclAllReduce of K*4+8 int64 per iteration")},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches,
        }
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mine", choices=["mine", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--keep-caches", action="store_true",
                    help="re-use the mirror / group summaries across fits (default: rebuilt by every fit)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "mine":
        args.warmup = 3  # timing rule: at least three warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    return run_mine(args)


if __name__ == "__main__":
    sys.exit(main())
