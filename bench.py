#!/usr/bin/env python
"""bench.py — read-pairs/s of the BWA-MEM candidate-generation path on B200 (BASELINE.json metric).

One "step" = one pass of the whole hot path (seeding -> SA -> chaining -> extension -> dedup + mate rescue -> CIGARs)
over one batch of barcode-sorted synthetic read pairs.  Workload at N=1: BASELINE.json configs[1] — synthetic 100 Mb
reference with 5 % segmental duplications, 1 M read pairs / 10 k barcodes, 50 kb molecules (`--config c2`).
`--config c3|c4` = the 3.1 Gb reference (index built on the box by lc_index_build) with `--pairs` read pairs.
  value : whole-job read-pairs/s with the reads already resident in HBM (lc_batch_run only)
  e2e   : the same through the C-ABI call a Lariat shim makes (lc_align_packed: pinned host reads -> H2D -> kernels -> D2H results);
          e2e.pageable = the same call fed from ordinary (pageable) host memory, which is what a cgo caller passes
  verified_pairs : the leading pairs of the timed workload whose results (all 88 bytes of every region, pos/strand/NM/CIGAR) were
          compared byte for byte, inside this run, with the CPU reference's output on the same pairs; a mismatch aborts the run
  roofline : dominant kernel (smem_kernel) algorithmic FM-index bytes / its CUDA-event time vs the measured HBM peak, plus one
          entry per kernel of the step (profiles/kernel_rooflines.json: ncu instruction / DRAM counters of the same workload)
  cpu_baseline : the reference's own CPU path (oracle/_ref/replay_ref, else the oracle port) on a bounded sample, all host cores
N>1 (torchrun, one rank per GPU): barcode groups shard across ranks, index replicated per GPU, no data-path collective.
  --scaling weak  (default): every rank owns its own `--pairs` pairs (per-GPU work fixed)
  --scaling strong: ONE read set of `--pairs` pairs, barcode groups cut by sharding.shard_groups (total work fixed)
`--mode percall`: the drop-in regime — T host threads each submitting one ~100-pair barcode group per call.
`--impl reference` times only the CPU reference arm.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from lariat_b200 import iofmt, sharding, synth  # noqa: E402

CACHE = os.path.join(ROOT, "data", "cache")
METRIC = "read-pairs/s aligned (bit-exact vs BWA)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


PAIRS_OVERRIDE = 0


def config_sizes(name, scale):
    g, r = synth.config_specs(name)
    if PAIRS_OVERRIDE:
        r.n_barcodes = max(20, int(r.n_barcodes * PAIRS_OVERRIDE / r.n_pairs))
        r.n_pairs = PAIRS_OVERRIDE
    if scale != 1.0:
        g.total_bp = int(g.total_bp * scale)
        r.n_pairs = max(2000, int(r.n_pairs * scale))
        r.n_barcodes = max(20, int(r.n_barcodes * scale))
    return g, r


def prepare_data(name, scale, rank_seed=0, need_index=True, device=0, need_pairs=True):
    """Genome + index (lc_index_build on the GPU, byte-identical to `bwa index`) + read pairs; cached under data/cache."""
    g, r = config_sizes(name, scale)
    r.seed += 1000 * rank_seed
    gname = "c3" if name == "c4" else name      # c4 runs on the c3 reference
    tag = f"{gname}_s{scale:g}" + (f"_bp{g.total_bp}" if gname == "c3" else "")
    d = os.path.join(CACHE, tag)
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, f"{gname}.fa")
    pairs = os.path.join(d, f"{name}.r{rank_seed}.n{r.n_pairs}.pairs")
    contigs = loci = None
    if need_index and not all(os.path.exists(fa + e) for e in (".bwt", ".sa", ".pac", ".ann", ".amb")):
        t = time.time()
        contigs, loci = synth.make_genome(g)
        t1 = time.time()
        from lariat_b200 import index_build
        index_build.build_index(fa, contigs, device)
        log(f"[bench] genome of {g.total_bp} bp in {t1 - t:.1f}s, index (lc_index_build) in {time.time() - t1:.1f}s")
    if need_pairs and not os.path.exists(pairs):
        t = time.time()
        if contigs is None:
            contigs, loci = synth.make_genome(g)
        sim = synth.simulate_pairs(contigs, loci, r)
        iofmt.write_pairs(pairs, sim["r1"], sim["r2"])
        log(f"[bench] simulated {r.n_pairs} pairs in {time.time() - t:.1f}s")
    return fa, pairs, g, r


def load_pairs(path):
    raw = np.fromfile(path, dtype=np.uint8)
    n = int(raw[4:8].view("<i4")[0])
    lens = raw[8:8 + 8 * n].view("<i4")
    off = np.zeros(2 * n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    bases = raw[8 + 8 * n:8 + 8 * n + int(off[-1])]
    return n, bases, off


def write_pairs_slice(path, bases, off, p0, p1):
    """LRP1 file of pairs [p0, p1) of a packed read set (the checker's input format)."""
    lens = np.diff(off[2 * p0:2 * p1 + 1]).astype("<i4")
    with open(path, "wb") as f:
        f.write(b"LRP1" + np.int32(p1 - p0).tobytes() + lens.tobytes() + bases[off[2 * p0]:off[2 * p1]].tobytes())


class ClockSampler(threading.Thread):
    def __init__(self, idx):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag = idx, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append([x.strip() for x in o])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = [x for x in self.samples if len(x) >= 6 and x[0].isdigit()]
        if not s:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for x in s for i in range(4) if x[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median([int(x[0]) for x in s])), "sm_max_mhz": float(s[0][1]), "reasons": reasons, "samples": len(s)}


def cpu_reference(fa, pairs, sample_pairs, threads, repeat=1, out="-"):
    """Times the reference CPU implementation (unmodified libbwa driven by oracle/replay.c) or, if it was not built here, the oracle
    port.  out != "-": also writes the LRO1 dump of the sample (the run-time verification of the CUDA arm compares with it)."""
    exe, kind = os.path.join(ROOT, "oracle", "_ref", "replay_ref"), "reference"
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
        exe, kind = os.path.join(ROOT, "oracle", "replay_port"), "port"
    res = subprocess.run([exe, fa, pairs, out, "-t", str(threads), "-n", str(sample_pairs), "-r", str(repeat)],
                         capture_output=True, text=True, check=True).stderr.strip().splitlines()[-1]
    j = json.loads(res)
    return kind, j


def workload_config(args):
    """The `config` object: identical in the CUDA arm and the reference arm (the driver compares them)."""
    workload = {"c1": "synthetic 10 Mb reference, 20k 2x150 pairs / 200 barcodes",
                "c2": "synthetic 100 Mb reference with 5% segmental duplications, 1M 2x150 read pairs / 10k barcodes, 50 kb molecules",
                "c3": "synthetic 3.1 Gb reference (24 contigs, random + 1M SINE-like / 100k LINE-like repeat copies + tandem arrays + 2% segdups), "
                      "2x150 read pairs over barcodes of ~100 pairs, 50 kb molecules",
                "c4": "long-molecule repeat stress on the 3.1 Gb reference: 200 kb molecules, 30% degraded mates, reads sampled from repeat copies"}.get(args.config, args.config)
    g, r = config_sizes(args.config, args.scale)
    return {"workload": workload, "config": args.config, "scale": args.scale, "reference_bp": g.total_bp, "pairs_per_gpu": r.n_pairs, "reads": "2x150",
            "batch": "one launch sequence per step over all barcode groups of the rank",
            "l2_policy": "inputs (FM-index + dense SA + per-step pools, all > 1 GB) exceed the 126 MB L2",
            "sharding": f"barcode groups split over {args.gpus} GPU(s), index replicated per GPU, {args.scaling} scaling"}


def kernel_rooflines(stage_ms, counters, clocks, peaks, n_pairs_scale, config="c2"):
    """One entry per kernel of the step.  Static part: profiles/kernel_rooflines.json, written by scripts/ncu_kernels.py from an
    `ncu --set full` capture of this workload (warp instructions, DRAM bytes, active threads per launch).  Live part: the CUDA-event
    stage times and work counters of THIS run.  frac_issue = warp-instructions / (time x SMs x 4 schedulers x SM clock): the share of the
    chip's instruction-issue slots the kernel used (the bound of the integer DP kernels); frac_hbm = DRAM bytes / time / HBM peak."""
    prof = None
    for name in (f"kernel_rooflines_{config}.json", "kernel_rooflines.json"):      # the capture of THIS workload if one is committed
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            break
        except Exception:
            continue
    if prof is None:
        return None
    out = {"source": prof.get("source"), "workload": prof.get("workload"), "kernels": {}}
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    n_sm = prof.get("n_sm", 148)
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for k, v in prof.get("kernels", {}).items():
        ms = v.get("ms")
        if not ms:
            continue
        e = {"ms_ncu": ms, "warp_instr": v.get("warp_instr"), "dram_bytes": v.get("dram_bytes"), "active_threads_per_warp": v.get("active_threads"),
             "frac_issue": v["warp_instr"] / (ms * 1e-3 * n_sm * 4 * v.get("sm_mhz", sm_mhz) * 1e6) if v.get("warp_instr") else None,
             "frac_hbm": v["dram_bytes"] / (ms * 1e-3) / 1e9 / hbm if v.get("dram_bytes") is not None else None,
             "bound": v.get("bound")}
        out["kernels"][k] = e
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--scale", type=float, default=float(os.environ.get("LARIAT_BENCH_SCALE", "1.0")))
    ap.add_argument("--pairs", type=int, default=0, help="override the config's number of read pairs (barcodes scale with it)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--cpu-sample", type=int, default=200000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline AND the run-time verification against it (diagnostic runs only)")
    ap.add_argument("--chunk-pairs", type=int, default=-1, help="lc_set_chunk_pairs for the e2e arm (-1: library default)")
    ap.add_argument("--e2e-sweep", default="", help="comma list of chunk sizes: time the e2e arm for each (stderr), then go on with --chunk-pairs")
    ap.add_argument("--mode", default="batch", choices=["batch", "percall"])
    ap.add_argument("--threads", default="32", help="percall mode: comma list of host thread counts")
    ap.add_argument("--call-pairs", type=int, default=100, help="percall mode: pairs per call (one barcode group)")
    ap.add_argument("--profile-pass", action="store_true", help="upload + one warm pass + one pass between cudaProfilerStart/Stop, then exit (for ncu)")
    args = ap.parse_args()
    global PAIRS_OVERRIDE
    PAIRS_OVERRIDE = args.pairs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    cfg = workload_config(args)

    if args.impl == "reference":
        if rank != 0:
            return
        fa, pairs, g, r = prepare_data(args.config, args.scale, 0, need_index=True, device=0)
        sample = min(r.n_pairs, args.cpu_sample)
        vals = []
        for _ in range(args.warmup):
            cpu_reference(fa, pairs, min(sample, 2000), ncores)
        for _ in range(args.steps):
            kind, j = cpu_reference(fa, pairs, sample, ncores)
            vals.append(j["pairs_per_s"])
        v = float(np.mean(vals))
        print(json.dumps({"metric": METRIC, "value": v, "unit": "read-pairs/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample / v, "higher_is_better": True,
                          "scaling": args.scaling, "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": cfg, "impl": "reference",
                          "cpu_baseline": {"value": v, "unit": "read-pairs/s", "cores": ncores, "kind": kind,
                                           "sample": f"first {sample} pairs of the workload per step, index preloaded, {ncores} threads"},
                          "e2e": {"value": v, "unit": "read-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # stdout carries exactly one JSON line: whatever native libraries write to fd 1 (NCCL prints its version banner there when the box sets
    # NCCL_DEBUG) goes to stderr instead; the JSON line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    import torch
    import torch.distributed as dist
    from lariat_b200 import gobwa
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl cuda needs a GPU: libcuda_lariat has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # data: rank 0 builds the shared index; weak scaling: every rank simulates its own barcode groups; strong: one shared read set
    if rank == 0:
        prepare_data(args.config, args.scale, 0, device=local, need_pairs=(args.scaling == "strong" or world == 1))
    if world > 1:
        dist.barrier()
    data_rank = 0 if args.scaling == "strong" else rank
    fa, pairs, g, r = prepare_data(args.config, args.scale, data_rank, device=local)
    n_pairs, bases, off = load_pairs(pairs)
    per_group = -(-r.n_pairs // r.n_barcodes)
    if args.scaling == "strong":
        # ONE read set: rank k takes the contiguous range of barcode groups shard_groups gives it
        sizes = [min(per_group, n_pairs - i * per_group) for i in range(-(-n_pairs // per_group))]
        g_lo, g_hi = sharding.shard_groups(sizes, world)[rank]
        p0, p1 = sharding.pair_range(sizes, g_lo, g_hi)
        bases = bases[off[2 * p0]:off[2 * p1]].copy()
        off = (off[2 * p0:2 * p1 + 1] - off[2 * p0]).copy()
        n_pairs = p1 - p0
        shard = {"total_groups": len(sizes), "this_rank_groups": [int(g_lo), int(g_hi)], "this_rank_pairs": [int(p0), int(p1)]}
    else:
        g_lo, g_hi = sharding.shard_groups([per_group] * (r.n_barcodes * world), world)[rank]
        shard = {"total_groups": r.n_barcodes * world, "this_rank_groups": [int(g_lo), int(g_hi)]}
    lib = gobwa.load_library()
    t0 = time.time()
    ref = gobwa.GoBwaLoadReference(fa, device=local)
    if ref is None:
        raise SystemExit("index load failed: " + gobwa.last_error())
    idx_bytes = sum(os.path.getsize(fa + e) for e in (".bwt", ".sa", ".pac"))
    log(f"[bench] rank {rank}: index ({idx_bytes / 1e9:.2f} GB of files) resident in HBM after {time.time() - t0:.1f}s; {n_pairs} pairs")
    # pinned host staging for the e2e arm
    hb = torch.from_numpy(bases.copy()).pin_memory()
    ho = torch.from_numpy(off).pin_memory()
    ctx = ref.BWTData
    res = C.POINTER(gobwa.lc_result_t)()

    def check(rc):
        if rc != 0:
            raise SystemExit(f"libcuda_lariat error {rc}: {gobwa.last_error()}")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.profile_pass:
        check(lib.lc_batch_upload(ctx, n_pairs, hb.data_ptr(), ho.data_ptr()))
        check(lib.lc_batch_run(ctx, 25))
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        check(lib.lc_batch_run(ctx, 25))
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    if args.mode == "percall":
        from lariat_b200 import percall
        out = percall.run(args, lib, ref, bases, off, n_pairs, cfg, METRIC, ncores, fa, pairs, cpu_reference if not args.no_cpu else None)
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
        return

    # ---- one counted pass (outside the timed region): algorithmic-work figures + run-time verification against the CPU reference
    lib.lc_set_counting(ctx, 1)
    check(lib.lc_align_packed(ctx, n_pairs, hb.data_ptr(), ho.data_ptr(), 25, C.byref(res)))
    br = gobwa.BatchResult(res.contents, copy=False)
    counters = br.counters
    n_cand = int(res.contents.n_regs)
    cpu = None
    verified = 0
    if rank == 0 and not args.no_cpu:
        sample = min(n_pairs, args.cpu_sample)
        got = br.dump_fast(sample)
        with tempfile.TemporaryDirectory() as td:
            sp = pairs
            if args.scaling == "strong":
                sp = os.path.join(td, "shard.pairs")
                write_pairs_slice(sp, bases, off, 0, sample)
            want_path = os.path.join(td, "want.out")
            kind, j = cpu_reference(fa, sp, sample, ncores, out=want_path)
            want = open(want_path, "rb").read()
        if got != want:
            raise SystemExit(f"VERIFICATION FAILED: the CUDA results of the first {sample} pairs differ from the CPU reference ({kind}) — no number is reported")
        verified = sample
        cpu = {"value": j["pairs_per_s"], "unit": "read-pairs/s", "cores": ncores, "kind": kind,
               "sample": f"first {sample} pairs of the workload, index preloaded, {ncores} threads, wall clock {j['seconds']:.2f}s; its output is what verified_pairs was compared with"}
        log(f"[bench] verified {sample} pairs bit-exact vs the CPU {kind} ({hashlib.md5(want).hexdigest()}); CPU {j['pairs_per_s']:.0f} pairs/s on {ncores} threads")
    lib.lc_set_counting(ctx, 0)
    del br
    lib.lc_free_result(ctx, res)      # results belong to the caller until it frees them (the reference: arena.Free(), lariat.go:546)

    # ---- device-resident arm: value
    # (the device-resident arm keeps one batch and its pools in HBM: on repeat-rich data the pools of 1 M pairs may not fit, then the
    # arm runs on the largest leading part of the workload that does — lc_align_packed, the e2e arm, chunks such calls by itself)
    n_res = n_pairs
    while True:
        check(lib.lc_batch_upload(ctx, n_res, hb.data_ptr(), ho.data_ptr()))
        rc = lib.lc_batch_run(ctx, 25)
        if rc == -5 and n_res > 20000:      # LC_ENOMEM
            n_res //= 2
            log(f"[bench] rank {rank}: device-resident batch reduced to {n_res} pairs ({gobwa.last_error()})")
            continue
        check(rc)
        break
    for _ in range(max(0, args.warmup - 1)):
        check(lib.lc_batch_run(ctx, 25))
    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    # timed on the device: CUDA events on the library's own stream around the K steps (torch.cuda.Event sees only torch's stream)
    t0 = time.perf_counter()
    check(lib.lc_timer_start(ctx))
    for _ in range(args.steps):
        check(lib.lc_batch_run(ctx, 25))
    ms_dev = C.c_float(0)
    check(lib.lc_timer_stop(ctx, C.byref(ms_dev)))
    sync_all()
    dt_wall = time.perf_counter() - t0
    dt = ms_dev.value * 1e-3
    # per-stage CUDA-event times of the last timed step (events are recorded on the library's own stream)
    check(lib.lc_batch_fetch(ctx, C.byref(res)))
    rr = res.contents
    stage_ms = {k: float(getattr(rr, k)) for k in ("ms_seed", "ms_sa", "ms_chain", "ms_extend", "ms_pair", "ms_cigar", "ms_total")}
    launches = int(rr.gpu_launches)
    dt_max = sharding.max_over_ranks(dt, device="cuda")
    total_pairs = sharding.sum_over_ranks(float(n_pairs), device="cuda")
    value = sharding.sum_over_ranks(float(n_res), device="cuda") * args.steps / dt_max

    # ---- end-to-end arm: host buffers in, host results out, every step
    all_keys = ("ms_h2d", "ms_seed", "ms_sa", "ms_chain", "ms_extend", "ms_pair", "ms_cigar", "ms_d2h", "ms_total")
    swept_env = set()
    for ent in [x for x in args.e2e_sweep.split(",") if x]:      # "N" or "N:KEY=VAL+KEY=VAL" (chunk-plan knobs of lc_api.cu, read per call)
        cp = int(ent.split(":")[0])
        for k in swept_env:
            os.environ.pop(k, None)
        for kv in (ent.split(":")[1].split("+") if ":" in ent else []):
            k, v = kv.split("=")
            os.environ[k] = v
            swept_env.add(k)
        check(lib.lc_set_chunk_pairs(ctx, cp))
        for _ in range(2):
            check(lib.lc_align_packed(ctx, n_pairs, hb.data_ptr(), ho.data_ptr(), 25, C.byref(res)))
            lib.lc_free_result(ctx, res)
        tw = time.perf_counter()
        for _ in range(args.steps):
            check(lib.lc_align_packed(ctx, n_pairs, hb.data_ptr(), ho.data_ptr(), 25, C.byref(res)))
            last = {k: round(float(getattr(res.contents, k)), 2) for k in all_keys}
            lib.lc_free_result(ctx, res)
        tw = (time.perf_counter() - tw) / args.steps
        log(f"[bench] e2e sweep chunk_pairs={ent}: {1e3 * tw:.2f} ms/step wall, {n_pairs / tw:.0f} pairs/s; last call " + json.dumps(last))
    if args.e2e_sweep:
        for k in swept_env:
            os.environ.pop(k, None)
        check(lib.lc_set_chunk_pairs(ctx, 500000))      # the library's default
    if args.chunk_pairs >= 0:
        check(lib.lc_set_chunk_pairs(ctx, args.chunk_pairs))

    def time_e2e(b_ptr, o_ptr):
        """K calls, each followed by lc_free_result (the result is the caller's until then); returns (seconds, facts of the last call)"""
        for _ in range(max(1, args.warmup - 1)):
            check(lib.lc_align_packed(ctx, n_pairs, b_ptr, o_ptr, 25, C.byref(res)))
            lib.lc_free_result(ctx, res)
        sync_all()
        t0 = time.perf_counter()
        check(lib.lc_timer_start(ctx))
        facts = None
        for _ in range(args.steps):
            check(lib.lc_align_packed(ctx, n_pairs, b_ptr, o_ptr, 25, C.byref(res)))
            rr = res.contents
            facts = (int(rr.n_regs), int(rr.n_cigar_words), int(rr.n_reads), {k: float(getattr(rr, k)) for k in all_keys})
            lib.lc_free_result(ctx, res)
        check(lib.lc_timer_stop(ctx, C.byref(ms_dev)))
        sync_all()
        wall = time.perf_counter() - t0
        # the call returns with the results on the host, so device events and the host clock bracket the same work; report the larger
        return max(ms_dev.value * 1e-3, wall), facts

    dte, (n_regs, n_cw, n_rd, e2e_ms) = time_e2e(hb.data_ptr(), ho.data_ptr())      # stage times are sums over the chunks of the call (chunks of the two lanes overlap)
    e2e = total_pairs * args.steps / sharding.max_over_ranks(dte, device="cuda")
    h2d = int(bases.nbytes + off.nbytes)
    d2h = int(n_regs * (88 + 32) + n_cw * 4 + (n_rd + 1) * 8)
    # the same call fed from pageable memory (what a cgo caller hands over: a Go slice)
    pb, po = np.array(bases, copy=True), np.array(off, copy=True)
    dtp, _ = time_e2e(pb.ctypes.data, po.ctypes.data)
    e2e_pageable = total_pairs * args.steps / sharding.max_over_ranks(dtp, device="cuda")
    sampler.stop_flag = True

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (smem_kernel): algorithmic FM-index bytes the reference would touch / CUDA-event time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    res_frac = n_res / float(n_pairs)      # the counted pass covered all n_pairs pairs, the timed stage n_res of them
    fm_bytes = 64.0 * counters["occ_blocks"] * res_frac
    ach = fm_bytes / (stage_ms["ms_seed"] * 1e-3) / 1e9 if stage_ms["ms_seed"] > 0 else 0.0
    clocks = sampler.summary()
    kr = kernel_rooflines(stage_ms, counters, clocks, peaks, res_frac, "c3" if args.config == "c4" else args.config)
    traffic = None
    limiter = None
    if kr and cfg["config"] == (kr.get("workload") or {}).get("config") and "smem_kernel" in kr["kernels"]:
        sk = kr["kernels"]["smem_kernel"]
        traffic = sk.get("dram_bytes")
        if sk.get("frac_hbm") is not None and sk["frac_hbm"] < 0.3:
            limiter = "L2 random-sector rate + issue slots: the index is largely L2-resident at this reference size (DRAM throughput %.0f%% of peak)" % (100 * sk["frac_hbm"])
    roof = {"kernel": "smem_kernel (+ seed_finish/redo: the seeding stage)", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "limiter": limiter, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
            "definition": "achieved = 64 B x the occ blocks the REFERENCE layout would touch (SURVEY 8d; counted in the kernel) / the stage's CUDA-event time; "
                          "traffic = ncu dram bytes of one launch (profiles/kernel_rooflines.json), null when no capture of this workload is committed",
            "algorithmic_bytes_per_launch": fm_bytes, "kernel_ms": stage_ms["ms_seed"],
            "sa_lookup_kernel": {"algorithmic_bytes": (64.0 * counters["lf_steps"] + 8.0 * counters["sa_calls"]) * res_frac, "ms": stage_ms["ms_sa"],
                                 "note": "the dense suffix array answers bwt_sa inside the seeding stage; algorithmic bytes are the reference's LF walks"},
            "extend_kernel": {"cells": counters["ext_cells"] * res_frac, "ms": stage_ms["ms_extend"],
                              "gcups": counters["ext_cells"] * res_frac / max(stage_ms["ms_extend"], 1e-9) / 1e6},
            "matesw_kernel": {"cells": counters["sw_cells"] * res_frac, "ms": stage_ms["ms_pair"],
                              "gcups": counters["sw_cells"] * res_frac / max(stage_ms["ms_pair"], 1e-9) / 1e6},
            "cigar_kernels": {"cells": counters["glb_cells"] * res_frac, "ms": stage_ms["ms_cigar"],
                              "gcups": counters["glb_cells"] * res_frac / max(stage_ms["ms_cigar"], 1e-9) / 1e6},
            "per_kernel": kr}
    out = {"metric": METRIC, "value": value, "unit": "read-pairs/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt_max / args.steps, "higher_is_better": True,
           "timing": "CUDA events on the library's stream around the K steps, max over ranks", "ms_per_step_wall": 1000.0 * dt_wall / args.steps,
           "scaling": args.scaling, "vs_baseline": None, "dtype": "int32",
           "data": "synthetic (seeded numpy genome + simulated linked reads; index built on the GPU by lc_index_build, byte-identical to bwa index)",
           "config": cfg, "shard": shard, "resident_batch_pairs": n_res, "index_file_bytes": idx_bytes, "candidates_per_step": n_cand,
           "verified_pairs": verified,
           "e2e": {"value": e2e, "unit": "read-pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "stage_ms": e2e_ms,
                   "input": "pinned host memory", "pageable": {"value": e2e_pageable, "unit": "read-pairs/s", "input": "pageable host memory (numpy arrays)"}},
           "gpu_launches": launches * args.steps, "stage_ms_last_step": stage_ms, "counters_per_step": counters,
           "roofline": roof, "cpu_baseline": cpu, "clocks": clocks}
    os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
