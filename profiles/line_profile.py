#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

    python profiles/line_profile.py <report.ncu-rep> <libmdkm.so> <mangled-kernel-substring>

Uses `nvdisasm -g` line info of the cubin inside the .so (compiled with -lineinfo) and the
`--page source` CSV of the report; instructions are matched by order within the function.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, kern = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines = sass.splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
    cur = ("?", 0)
    instr_lines = []
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .text"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            inl = re.search(r"inlined at \"([^\"]+)\", line (\d+)", l)
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            instr_lines.append(cur)
    src = list(csv.reader(io.StringIO(subprocess.run(
        ["ncu", "-i", rep, "--page", "source", "--csv", "--print-kernel-base", "function"],
        capture_output=True, text=True).stdout)))
    hdr = src[1]
    iex, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    body = []
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) > iex:
            body.append(r)
    n = min(len(body), len(instr_lines))
    print(f"sass instrs: report {len(body)}, disasm {len(instr_lines)}")
    ex, st = collections.Counter(), collections.Counter()
    for r, key in zip(body[:n], instr_lines[:n]):
        ex[key] += int(r[iex])
        st[key] += int(r[ist])
    tot, stot = sum(ex.values()), sum(st.values())
    srcs = {}
    for (f, ln), v in (sorted(ex.items()) if os.environ.get("ALL_LINES") else ex.most_common(45)):
        path = None
        for root in ("3d-point-cloud-multiday-imagery_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                path = p
                break
        text = ""
        if path:
            srcs.setdefault(path, open(path).read().splitlines())
            if 0 < ln <= len(srcs[path]):
                text = srcs[path][ln - 1].strip()[:90]
        print(f"{100*v/tot:5.1f}% exec {100*st[(f, ln)]/max(stot,1):5.1f}% stall  {f}:{ln:<5d} {text}")


if __name__ == "__main__":
    main()
