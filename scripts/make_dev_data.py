#!/usr/bin/env python
"""Small C2-like / C4-like datasets for CPU-side development (hostsim, oracle): genome + index + pairs under data/dev/.
usage: make_dev_data.py <name: c2s|c4s> [genome_bp] [n_pairs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lariat_b200 import synth, iofmt, index_build
name = sys.argv[1]
bp = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20_000
if name == "c2s":
    g = synth.GenomeSpec(bp, 4, 2345, segdup_frac=0.05, name="c2chr")
    r = synth.ReadSpec(n, max(20, n // 100), seed=778)
else:
    scale = bp / 3_100_000_000
    g = synth.GenomeSpec(bp, 4, 3345, segdup_frac=0.02, families=[(300, max(1, int(1_000_000 * scale)), 0.10, 0.15), (6000, max(1, int(100_000 * scale)), 0.02, 0.08)],
                         n_tandem=max(1, int(20_000 * scale)), name="c3chr")
    r = synth.ReadSpec(n, max(20, n // 100), mol_len=200_000, mols_per_barcode=50, degraded_frac=0.30, repeat_bias=0.8, seed=780)
d = os.path.join(ROOT, "data", "dev")
os.makedirs(d, exist_ok=True)
t = time.time()
contigs, loci = synth.make_genome(g)
fa = os.path.join(d, name + ".fa")
synth.write_fasta(fa, contigs)
index_build.build_index_torch(fa, contigs)
sim = synth.simulate_pairs(contigs, loci, r)
iofmt.write_pairs(os.path.join(d, name + ".pairs"), sim["r1"], sim["r2"])
print(f"{name}: {bp} bp, {n} pairs in {time.time() - t:.1f}s -> {fa}")
