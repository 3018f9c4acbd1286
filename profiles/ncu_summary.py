#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`): headline metrics, instruction mix, stall mix.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-index]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        print("==", r[hdr.index("Kernel Name")][:80])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]} {units[hdr.index(k)]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.2f} warps/issue")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-kernel-base", "function"]))))
    hdr = src[1]
    isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > iex:
            body.append(r)
    tot = sum(int(r[iex]) for r in body)
    stot = sum(int(r[ist]) for r in body)
    op, st = collections.Counter(), collections.Counter()
    for r in body:
        m = r[isrc].split()
        name = (m[1] if m[0].startswith("@") else m[0]).split(".")[0]
        op[name] += int(r[iex])
        st[name] += int(r[ist])
    print(f"== SASS mix (first kernel): {len(body)} instrs, {tot} warp-instr executed, {stot} stall samples")
    for k, v in op.most_common(24):
        print(f"  {k:10s} {v:12d} {100*v/tot:5.1f}%   stall samples {100*st[k]/max(stot,1):5.1f}%")


if __name__ == "__main__":
    main()
