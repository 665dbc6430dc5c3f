"""`bench.py --mode percall`: the drop-in regime of the boundary.

Lariat's worker goroutines hand BWA one barcode group (~100 read pairs) per call (reference: go/src/inference/lariat.go:348-376,
461-547).  This arm measures exactly that against libcuda_lariat.so with the C driver lariat_b200/csrc/lc_percall_driver.c (plain C
over the C ABI, pthreads, pageable input, every worker keeps its previous result alive while it makes the next call): read-pairs/s
and per-call latency for each host thread count, next to the big-batch figure of the default mode.  The driver's output for the
first `--cpu-sample` pairs is compared byte for byte with the CPU reference inside the run.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(_HERE, "lc_percall_driver")


def run_driver(fa, pairs, out, threads, call_pairs, n_pairs, repeat=1, clones=1, hold=1, device=0, env=None, devices=1):
    cmd = [DRIVER, fa, pairs, out, "-t", str(threads), "-c", str(call_pairs), "-n", str(n_pairs), "-r", str(repeat), "-k", str(clones),
           "-h", str(hold), "-g", str(device), "-G", str(devices)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    if r.returncode != 0:
        raise SystemExit(f"lc_percall_driver failed: {r.stderr[-2000:]}")
    return json.loads(r.stderr.strip().splitlines()[-1])


def run(args, lib, ref, bases, off, n_pairs, cfg, metric, ncores, fa, pairs, cpu_reference):
    if not os.path.exists(DRIVER):
        raise SystemExit(f"{DRIVER} not built (make -C lariat_b200/csrc)")
    ref.close()      # the driver processes own the GPU from here
    device = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    verified = 0
    if cpu_reference is not None:
        sample = min(n_pairs, args.cpu_sample)
        with tempfile.TemporaryDirectory() as td:
            want_p, got_p = os.path.join(td, "want.out"), os.path.join(td, "got.out")
            kind, j = cpu_reference(fa, pairs, sample, ncores, out=want_p)
            run_driver(fa, pairs, got_p, 32, args.call_pairs, sample, device=device)
            if open(got_p, "rb").read() != open(want_p, "rb").read():
                raise SystemExit("VERIFICATION FAILED: per-call results differ from the CPU reference — no number is reported")
        verified = sample
        cpu = {"value": j["pairs_per_s"], "unit": "read-pairs/s", "cores": ncores, "kind": kind,
               "sample": f"first {sample} pairs of the workload, index preloaded, {ncores} threads, wall clock {j['seconds']:.2f}s"}
    sweep = []
    devices = int(os.environ.get("LC_PERCALL_DEVICES", "1"))      # > 1: ONE driver process over that many GPUs (the driver's -G), as a Go binary would
    for t in [int(x) for x in str(args.threads).split(",") if x]:
        # enough work per point for a stable rate: at least ~2 s at 1 M pairs/s, bounded by the workload
        j = run_driver(fa, pairs, "-", t, args.call_pairs, n_pairs, repeat=max(1, args.steps // 2), device=device, devices=devices)
        sweep.append(j)
    best = max(sweep, key=lambda x: x["pairs_per_s"])
    first = sweep[0]
    return {"metric": metric, "mode": "percall", "value": first["pairs_per_s"], "unit": "read-pairs/s", "n_gpus": devices, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": first["call_ms_p50"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": cfg,
            "what": f"{args.call_pairs}-pair calls (one barcode group each) from T host threads through lc_align_packed, pageable input, results held "
                    "across the next call then freed; value = the first thread count of --threads; wall clock of the driver's timed section",
            "percall": sweep, "best": {"threads": best["threads"], "pairs_per_s": best["pairs_per_s"]},
            "verified_pairs": verified, "cpu_baseline": cpu,
            "e2e": {"value": first["pairs_per_s"], "unit": "read-pairs/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None}}
