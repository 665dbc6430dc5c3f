#!/usr/bin/env python
"""Merge a partial per-kernel capture (ncu_kernels.py output for a subset of kernels, e.g. `ncu -k regex:...`) into a committed
profiles/kernel_rooflines_<config>.json: entries of the new capture replace the old ones by kernel name.
usage: ncu_merge.py <base.json> <new.json> <out.json>"""
import json
import sys

base, new, out = sys.argv[1:4]
b, n = json.load(open(base)), json.load(open(new))
for k, v in n.get("kernels", {}).items():
    v["source"] = n.get("source")
    b["kernels"][k] = v
b["source"] = f"{b.get('source')}; kernels re-captured later: {sorted(n.get('kernels', {}))} from {n.get('source')}"
b["step_ms_under_ncu"] = round(sum(v.get("ms", 0.0) for v in b["kernels"].values()), 3)
json.dump(b, open(out, "w"), indent=1)
