#!/usr/bin/env python
"""ncu raw CSV (ncu -i x.ncu-rep --page raw --csv) of one profiled step -> per-kernel roofline facts.

usage: ncu_kernels.py <raw.csv> <out.json> <config> <pairs> [source-note]
For every kernel name (template arguments stripped, launches summed): time, warp instructions, DRAM bytes, active threads per warp,
achieved occupancy, issue-slot utilisation, and the two fractions bench.py reports next to its live stage times:
  frac_issue = warp instructions / (time x SMs x 4 schedulers x SM clock)   — the bound of the integer DP kernels
  frac_hbm   = DRAM bytes / time / measured HBM peak (MEASURED_PEAKS.json)   — the bound of the FM-index walks when the index >> L2
"""
import csv
import json
import os
import re
import sys

raw, out, config, pairs = sys.argv[1:5]
note = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(raw)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(raw, errors="ignore")))
# the header is the first row that starts with "ID"
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[h], rows[h + 1]
col = {n: i for i, n in enumerate(names)}
scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
scale_t = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "nsecond": 1e-6, "second": 1e3}


def num(r, k, scale=None):
    if k not in col:
        return None
    try:
        v = float(r[col[k]].replace(",", ""))
    except ValueError:
        return None
    if v != v:      # "nan": the metric was not collected for this launch
        return None
    if scale is not None:
        v *= scale.get(units[col[k]], 1)
    return v


try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    peaks = {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
kern = {}
n_sm = 148
for r in rows[h + 2:]:
    if len(r) < len(names):
        continue
    name = re.sub(r"<.*", "", r[col["Kernel Name"]]).strip()
    name = re.sub(r"\(.*", "", name)
    if name.startswith("void "):
        name = name[5:]
    k = kern.setdefault(name, {"launches": 0, "ms": 0.0, "warp_instr": 0.0, "dram_bytes": 0.0, "_at": 0.0, "_occ": 0.0, "_issue": 0.0, "_clk": 0.0, "_l2hit": 0.0, "regs": 0})
    ms = num(r, "gpu__time_duration.sum", scale_t) or 0.0
    k["launches"] += 1
    k["ms"] += ms
    wi = num(r, "sm__inst_executed.sum")
    if wi is None:      # the full set reports the rate: warp instructions per cycle per SM x elapsed cycles x SMs
        ipc, cyc = num(r, "sm__inst_executed.avg.per_cycle_elapsed"), num(r, "sm__cycles_elapsed.avg")
        wi = ipc * cyc * n_sm if ipc is not None and cyc is not None else 0.0
    k["warp_instr"] += wi
    k["dram_bytes"] += (num(r, "dram__bytes_read.sum", scale_b) or 0.0) + (num(r, "dram__bytes_write.sum", scale_b) or 0.0)
    k["_at"] += ms * (num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") or 0.0)
    k["_occ"] += ms * (num(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0.0)
    k["_issue"] += ms * (num(r, "smsp__issue_active.avg.pct") or num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or 0.0)
    clk = num(r, "sm__cycles_elapsed.avg.per_second")
    if clk is not None:
        u = units[col["sm__cycles_elapsed.avg.per_second"]]
        clk *= {"Ghz": 1e3, "GHz": 1e3, "Mhz": 1, "MHz": 1, "hz": 1e-6, "Hz": 1e-6, "cycle/nsecond": 1e3, "cycle/usecond": 1, "cycle/second": 1e-6}.get(u, 1)
    k["_clk"] += ms * (clk or 0.0)
    k["_l2hit"] += ms * (num(r, "lts__t_sector_hit_rate.pct") or 0.0)
    k["regs"] = max(k["regs"], int(num(r, "launch__registers_per_thread") or 0))
res = {}
for name, k in kern.items():
    ms = k["ms"]
    if ms <= 0:
        continue
    sm_mhz = k["_clk"] / ms if k["_clk"] else float(peaks.get("sm_max_mhz", 1965.0))
    fi = k["warp_instr"] / (ms * 1e-3 * n_sm * 4 * sm_mhz * 1e6) if k["warp_instr"] else None
    fh = k["dram_bytes"] / (ms * 1e-3) / 1e9 / hbm
    res[name] = {"launches": k["launches"], "ms": round(ms, 4), "warp_instr": k["warp_instr"], "dram_bytes": k["dram_bytes"],
                 "active_threads": round(k["_at"] / ms, 2), "achieved_occupancy_pct": round(k["_occ"] / ms, 1), "issue_active_pct": round(k["_issue"] / ms, 1),
                 "l2_hit_pct": round(k["_l2hit"] / ms, 1), "sm_mhz": round(sm_mhz, 1), "registers": k["regs"],
                 "frac_issue": None if fi is None else round(fi, 4), "frac_hbm": round(fh, 4),
                 "bound": "hbm" if fh > (fi or 0) else "issue"}
tot = sum(v["ms"] for v in res.values())
for v in res.values():
    v["share_of_step"] = round(v["ms"] / tot, 4)
json.dump({"source": note, "workload": {"config": config, "pairs": int(pairs)}, "n_sm": n_sm, "hbm_peak_gbs": hbm,
           "note": "per-launch times under ncu are cold-cache and serialised: shares, instruction and byte counts are the facts; live times are in the bench line",
           "step_ms_under_ncu": round(tot, 3), "kernels": dict(sorted(res.items(), key=lambda kv: -kv[1]["ms"]))}, open(out, "w"), indent=1)
for n, v in sorted(res.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{n:40s} {v['ms']:10.3f} ms {100 * v['share_of_step']:5.1f}%  issue {v['frac_issue']}  hbm {v['frac_hbm']}  thr/warp {v['active_threads']}  occ {v['achieved_occupancy_pct']}%  regs {v['registers']}")
