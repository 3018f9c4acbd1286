#!/usr/bin/env python
"""Turns what tools/final_check.sh <tag> left in gpurun_out/ into the tracked summaries of profiles/r2/.

    python tools/collect_profiles.py <tag> [prefix]
"""
import collections, csv, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else "final"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles", "r2")
os.makedirs(dst, exist_ok=True)


def launch_summary(path, out):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, mi, vi, idi, gi = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    d, names, grids = collections.defaultdict(dict), {}, {}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        d[r[idi]][r[mi]] = float(r[vi].replace(",", ""))
        names[r[idi]] = r[ki].split("(")[0][-60:]
        grids[r[idi]] = r[gi]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, ""])
    for i, m in d.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0)
        a[2] += m.get("dram__bytes_read.sum", 0)
        a[3] += m.get("dram__bytes_write.sum", 0)
        a[4] = grids[i]
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("per kernel: launches, mean duration (ncu: cold caches between kernels, serialised -- compare SHARES), "
                "share of the summed device time, mean DRAM read / write per launch\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:60s} n={a[0]:4d} mean_us={a[1]/a[0]/1e3:9.2f} share={a[1]/tot:6.3f} "
                    f"rd_MB={a[2]/a[0]/1e6:8.2f} wr_MB={a[3]/a[0]/1e6:8.2f} grid={a[4]}\n")


def copy(name, new):
    p = os.path.join(src, name)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, new))
        return True
    return False


for c in ("c2", "c3", "c4", "c5"):
    copy(f"bench_{tag}_{c}.json", f"{pre}_bench_n1_{c}.json")
copy(f"pytest_gpu_{tag}.log", f"{pre}_pytest_gpu.log")
copy(f"smoke_{tag}.log", f"{pre}_smoke.log")
if copy(f"launches_{tag}.csv", f"{pre}_launches_c2.csv"):
    launch_summary(os.path.join(src, f"launches_{tag}.csv"), os.path.join(dst, f"{pre}_launches_c2_summary.txt"))
for rep, name in ((f"prof_step_{tag}.ncu-rep", f"{pre}_step_kernel_ncu_k16.txt"),
                  (f"prof_build_{tag}.ncu-rep", f"{pre}_build_unproject_final_kernels_ncu_c2.txt"),
                  (f"prof_step_c4_{tag}.ncu-rep", f"{pre}_step_kernel_ncu_k1024.txt")):
    p = os.path.join(src, rep)
    if os.path.exists(p):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "ncu_summary.py"), p],
                             capture_output=True, text=True).stdout
        open(os.path.join(dst, name), "w").write(out)
print(sorted(os.listdir(dst)))
