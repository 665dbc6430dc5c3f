#!/usr/bin/env python
"""BASELINE.json configs[4]: ksw_extend2 / SMEM sweep — read length 100-250, band width 100, mismatch+indel rate 0-10 %.
For every (read length, error rate) the whole path runs on 500 k simulated pairs (10^6 reads) over a C2-sized 100 Mb reference; the extension and seeding
kernels are timed in place (CUDA events of the library's own stages) and divided into the algorithmic work the reference would do,
counted on the device in a separate pass (lc_set_counting): cells of ksw_extend2 as it iterates them (B/ksw.c:424) -> GCUPS;
64 B x the occ blocks bwt_2occ4 would touch (B/bwt.c:194) -> FM-index GB/s.  One JSON line per point on stdout.
usage: python scripts/bench_micro.py [--pairs N] [--genome BP]"""
import argparse, ctypes as C, json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lariat_b200 import gobwa, index_build, iofmt, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=500_000)
ap.add_argument("--genome", type=int, default=100_000_000)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
import torch  # noqa: E402
assert torch.cuda.is_available(), "needs a GPU: libcuda_lariat has no CPU path"
lib = gobwa.load_library()
with tempfile.TemporaryDirectory() as td:
    g = synth.GenomeSpec(args.genome, 8, 4242, segdup_frac=0.05, name="c5chr")
    contigs, loci = synth.make_genome(g)
    fa = os.path.join(td, "c5.fa")
    synth.write_fasta(fa, contigs)
    index_build.build_index(fa, contigs, "cuda")
    ref = gobwa.GoBwaLoadReference(fa)
    assert ref is not None, gobwa.last_error()
    ctx = ref.BWTData
    res = C.POINTER(gobwa.lc_result_t)()
    for L in (100, 125, 150, 200, 250):
        for err in (0.0, 0.01, 0.02, 0.05, 0.10):
            r = synth.ReadSpec(args.pairs, max(20, args.pairs // 100), read_len=L, sub_rate=err * 0.8, indel_rate=err * 0.2, seed=900 + L)
            sim = synth.simulate_pairs(contigs, loci, r)
            p = os.path.join(td, "p.pairs")
            iofmt.write_pairs(p, sim["r1"], sim["r2"])
            raw = np.fromfile(p, dtype=np.uint8)
            n = int(raw[4:8].view("<i4")[0])
            lens = raw[8:8 + 8 * n].view("<i4")
            off = np.zeros(2 * n + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            bases = np.ascontiguousarray(raw[8 + 8 * n:8 + 8 * n + int(off[-1])])
            lib.lc_set_counting(ctx, 1)
            assert lib.lc_align_packed(ctx, n, bases.ctypes.data, off.ctypes.data, 25, C.byref(res)) == 0, gobwa.last_error()
            cnt = gobwa.BatchResult(res.contents, copy=False).counters
            lib.lc_free_result(ctx, res)
            lib.lc_set_counting(ctx, 0)
            assert lib.lc_batch_upload(ctx, n, bases.ctypes.data, off.ctypes.data) == 0
            ms = {"ms_seed": [], "ms_extend": [], "ms_total": []}
            for s in range(args.steps + 2):
                assert lib.lc_batch_run(ctx, 25) == 0, gobwa.last_error()
                assert lib.lc_batch_fetch(ctx, C.byref(res)) == 0
                if s >= 2:
                    for k in ms:
                        ms[k].append(float(getattr(res.contents, k)))
            m = {k: float(np.median(v)) for k, v in ms.items()}
            print(json.dumps({"read_len": L, "error_rate": err, "band_w": 100, "pairs": n,
                              "ksw_extend2_gcups": cnt["ext_cells"] / m["ms_extend"] / 1e6, "ext_calls": cnt["ext_calls"], "ext_cells": cnt["ext_cells"],
                              "fm_index_gbs": 64.0 * cnt["occ_blocks"] / m["ms_seed"] / 1e6, "bwt_extend_calls": cnt["extend_calls"],
                              "pairs_per_s": n / m["ms_total"] * 1e3, **m}), flush=True)
    ref.close()
