#!/usr/bin/env python
"""Record DRAM traffic of one kernel launch from an ncu --set full capture into profiles/ncu_traffic.json.
usage: ncu_traffic.py <rep.ncu-rep> <kernel-key> <workload-key e.g. c2_s1>"""
import csv, io, json, os, subprocess, sys
rep, kern, wl = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
d = {k: (x, uu) for k, uu, x in zip(h, u, v)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
tot = sum(float(d[k][0]) * scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
ms = float(d["gpu__time_duration.sum"][0]) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[d["gpu__time_duration.sum"][1]]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
j = json.load(open(path)) if os.path.exists(path) else {}
j.setdefault(wl, {})[kern] = {"dram_bytes": tot, "kernel_ms_under_ncu": ms, "source": os.path.basename(rep)}
json.dump(j, open(path, "w"), indent=1, sort_keys=True)
print(wl, kern, tot / 1e9, "GB", ms, "ms")
