#!/usr/bin/env python
"""Prints one short line per bench JSON line read from the given files (diagnostics)."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print(f"{d['config']['workload'][:40]:40s} value {d['value']:.4g} ms/step {d['ms_per_step']:.3f} "
              f"iter_ms {r.get('avg_launch_ms', 0):.4f} e2e_ms {e.get('ms_per_step', 0) or 0:.2f}")
