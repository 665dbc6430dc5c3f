#!/usr/bin/env python
"""Randomised CPU campaign for the kernels' stage bodies (tests/hostsim) against the oracle port: read lengths 76-250, substitution
rates 0.5-6 %, an indel in 30-90 % of the reads, up to 30 % degraded mates, repeat families and tandem arrays — every case must be the
same bytes.  The pytest suite runs two such cases (tests/test_hostsim.py::test_cigar_bands_on_indel_rich_reads); this runs fifteen
(about four minutes) and prints how the CIGAR jobs spread over the band classes.  usage: hostsim_campaign.py [workdir]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lariat_b200 import index_build, iofmt, synth  # noqa: E402

HS = os.path.join(ROOT, "tests", "_build", "hostsim")      # built by the hostsim tests (tests/conftest.py)
RP = os.path.join(ROOT, "oracle", "replay_port")
work = sys.argv[1] if len(sys.argv) > 1 else "/tmp/hostsim_campaign"
bad, seed = 0, 1000
for read_len in (76, 100, 150, 200, 250):
    for sub, ind, deg in ((0.005, 0.3, 0.0), (0.03, 0.9, 0.1), (0.06, 0.9, 0.3)):
        seed += 7
        d = os.path.join(work, f"c{seed}")
        os.makedirs(d, exist_ok=True)
        g = synth.GenomeSpec(300_000, 3, seed, segdup_frac=0.05, segdup_len=(500, 3000), families=[(300, 80, 0.03, 0.15), (1500, 20, 0.02, 0.08)],
                             n_tandem=10, name="k")
        r = synth.ReadSpec(600, 10, read_len=read_len, mol_len=40_000, mols_per_barcode=8, sub_rate=sub, indel_rate=ind / read_len,
                           degraded_frac=deg, repeat_bias=0.3, insert_mean=2.2 * read_len + 50, insert_clip=(read_len + 20, 4 * read_len + 100),
                           seed=seed + 1)
        contigs, loci = synth.make_genome(g)
        fa, pairs = os.path.join(d, "k.fa"), os.path.join(d, "k.pairs")
        synth.write_fasta(fa, contigs)
        index_build.build_index_torch(fa, contigs)
        sim = synth.simulate_pairs(contigs, loci, r)
        iofmt.write_pairs(pairs, sim["r1"], sim["r2"])
        t = time.time()
        subprocess.check_call([RP, fa, pairs, os.path.join(d, "want.out"), "-t", "4"], stderr=subprocess.DEVNULL)
        out = subprocess.run([HS, fa, pairs, os.path.join(d, "got.out")], stderr=subprocess.PIPE, text=True, check=True).stderr
        same = open(os.path.join(d, "want.out"), "rb").read() == open(os.path.join(d, "got.out"), "rb").read()
        bad += not same
        print(read_len, sub, ind, deg, seed, "OK" if same else "MISMATCH", out.strip().split("\n")[-1][:200], f"{time.time() - t:.0f}s", flush=True)
print("mismatches", bad)
sys.exit(1 if bad else 0)
