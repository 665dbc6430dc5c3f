"""CPU check of the product's per-read stage bodies and host orchestration (tests/hostsim: the same
__host__ __device__ code the CUDA kernels run, executed serially) against the reference golden.
This is test scaffolding only; the product library has no CPU execution path."""
import os
import subprocess

from conftest import GOLDEN


def test_stage_bodies_match_reference_golden(hostsim_bin, tmp_path):
    out = tmp_path / "sim.out"
    subprocess.check_call([hostsim_bin, os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out)],
                          stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()


def test_result_independent_of_worker_count_and_delta(hostsim_bin, tmp_path):
    a, b = tmp_path / "a.out", tmp_path / "b.out"
    args = [os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs")]
    subprocess.check_call([hostsim_bin] + args + [str(a), "-w", "1", "-n", "200"], stderr=subprocess.DEVNULL)
    subprocess.check_call([hostsim_bin] + args + [str(b), "-w", "17", "-n", "200"], stderr=subprocess.DEVNULL)
    assert a.read_bytes() == b.read_bytes()


import pytest


@pytest.mark.parametrize("extra", [["-S"], ["-W"], ["-C", "6"], ["-C", "3", "-M", "4"], ["-S", "-C", "12", "-w", "5", "-W"], ["-I", "3"], ["-I", "2", "-W", "-C", "7"],
                                   ["-C", "6", "-D", "9", "-M", "4", "-E", "7"], ["-C", "3", "-D", "4", "-W"], ["-J", "2"], ["-J", "0"],
                                   ["-Z", "100000"], ["-Z", "100000", "-W", "-S"], ["-P"], ["-P", "-J", "2"]])
def test_seeding_state_machine_fallbacks(hostsim_bin, tmp_path, extra):
    """The seeding state machine (lc_seed.cuh) with tiny list capacities (most reads go to the second pass; with -D/-E tiny
    too, on to the one-thread-per-read body), with idle ticks before every housekeeping step (-I), and with the sampled suffix
    array instead of the dense one; -Z: the low-latency route of small batches (per-read seeding and extension bodies for every
    read); -P: the cooperative form of mem_align1_core's dedup (stage 5a, long lists): results must not change."""
    out = tmp_path / "sim.out"
    subprocess.check_call([hostsim_bin, os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out)] + extra,
                          stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()


@pytest.mark.parametrize("chunk,extra", [("64", []), ("101", []), ("600", []), ("64", ["-L"]), ("97", ["-L", "-x"])])
def test_chunked_call_equals_one_batch(hostsim_bin, tmp_path, chunk, extra):
    """lc_align_packed cuts large calls into chunks of pairs (slices of the packed arrays; candidate and CIGAR indices offset by the
    totals of the chunks before): the pipeline pieces it uses for that (rebased upload, run_front, run_pack_regs, the regions' fetch,
    run_cigar, run_pack_aln, the alignments' fetch), in both orders a lane can take (-L: every second chunk runs its CIGAR stage
    before its regions are packed)."""
    out = tmp_path / "sim.out"
    if "-x" in extra:
        extra = ["-L", "-x", str(tmp_path / "sim.x")]
    subprocess.check_call([hostsim_bin, os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out), "-K", chunk] + extra,
                          stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()


def test_stage_bodies_on_fresh_repeat_rich_data(hostsim_bin, oracle_port, tmp_path):
    """Fresh seeded data every run (not only the committed golden set): a small c4-shaped workload — repeat families, tandem
    arrays, long molecules, degraded mates, repeat-biased sampling — through the kernels' stage bodies on the CPU, with idle ticks
    before every housekeeping step and 64-bit interval arithmetic, against the oracle replay."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lariat_b200 import index_build, iofmt, synth
    g = synth.GenomeSpec(1_200_000, 6, 4711, segdup_frac=0.03, segdup_len=(1000, 6000),
                         families=[(300, 400, 0.10, 0.15), (3000, 40, 0.02, 0.08)], n_tandem=30, name="fchr")
    r = synth.ReadSpec(1500, 15, mol_len=200_000, mols_per_barcode=20, degraded_frac=0.30, repeat_bias=0.8, seed=4712)
    contigs, loci = synth.make_genome(g)
    fa = str(tmp_path / "f.fa")
    synth.write_fasta(fa, contigs)
    index_build.build_index_torch(fa, contigs)
    sim = synth.simulate_pairs(contigs, loci, r)
    pairs = str(tmp_path / "f.pairs")
    iofmt.write_pairs(pairs, sim["r1"], sim["r2"])
    want, got = tmp_path / "want.out", tmp_path / "got.out"
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), fa, pairs, str(want), "-t", "4"], stderr=subprocess.DEVNULL)
    subprocess.check_call([hostsim_bin, fa, pairs, str(got), "-I", "2", "-W", "-J", "6"], stderr=subprocess.DEVNULL)
    assert got.read_bytes() == want.read_bytes()


@pytest.mark.parametrize("read_len,sub_rate", [(100, 0.03), (250, 0.02)])
def test_cigar_bands_on_indel_rich_reads(hostsim_bin, oracle_port, tmp_path, read_len, sub_rate):
    """The 16-bit band kernels of the CIGAR stage (lc_cig_band: live-slot masks, F chain, direction bits, traceback) on reads where
    most candidates need a real banded alignment: an indel in most reads, 2-3 % substitutions (wide bands), short and long reads."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lariat_b200 import index_build, iofmt, synth
    g = synth.GenomeSpec(400_000, 3, 911 + read_len, segdup_frac=0.05, segdup_len=(500, 3000), families=[(300, 60, 0.05, 0.12)], n_tandem=8, name="ichr")
    r = synth.ReadSpec(700, 10, read_len=read_len, mol_len=40_000, mols_per_barcode=8, sub_rate=sub_rate, indel_rate=0.9 / read_len,
                       degraded_frac=0.10, insert_mean=2.2 * read_len + 50, insert_clip=(read_len + 20, 4 * read_len + 100), seed=913 + read_len)
    contigs, loci = synth.make_genome(g)
    fa = str(tmp_path / "i.fa")
    synth.write_fasta(fa, contigs)
    index_build.build_index_torch(fa, contigs)
    sim = synth.simulate_pairs(contigs, loci, r)
    pairs = str(tmp_path / "i.pairs")
    iofmt.write_pairs(pairs, sim["r1"], sim["r2"])
    want, got = tmp_path / "want.out", tmp_path / "got.out"
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), fa, pairs, str(want), "-t", "4"], stderr=subprocess.DEVNULL)
    out = subprocess.run([hostsim_bin, fa, pairs, str(got)], stderr=subprocess.PIPE, text=True, check=True).stderr
    assert got.read_bytes() == want.read_bytes()
    import re
    m = re.search(r"band class (\d+)/(\d+)/(\d+)/(\d+)/(\d+), fallback (\d+)", out)
    bands = [int(x) for x in m.groups()[:5]]
    assert sum(bands) > 500 and sum(1 for x in bands if x > 0) >= 3, out      # the band kernels did the work, in several classes


@pytest.mark.parametrize("extra", [["-W"], ["-W", "-S"], ["-W", "-C", "6", "-M", "4", "-D", "9", "-E", "7"], ["-W", "-I", "2", "-K", "101"]])
def test_multi_superblock_and_high_interval_bits(hostsim_sb_bin, tmp_path, extra):
    """What an index of >= 2^32 symbols exercises and a small one never does: superblock index >= 1 in every occ lookup and LF step,
    non-zero high bits in the packed 64-bit interval entries (LcPkOps / LcEntOps), the 64-bit dense suffix array.  The harness is
    rebuilt with 2^12-block superblocks and an 18-bit low word, so the tiny golden index has 4 superblocks and 20-bit intervals."""
    out = tmp_path / "sim.out"
    subprocess.check_call([hostsim_sb_bin, os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out)] + extra,
                          stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()


@pytest.mark.parametrize("extra", [[], ["-K", "101", "-W"], ["-w", "7"]])
def test_getalignments_stage_matches_reference_golden(hostsim_bin, tmp_path, extra):
    """Stage 6b (lc_post.cuh: the CIGAR walk of inference.GetAlignments, lariat.go:1573-1650) against the golden records the
    compiled reference + the C restatement of the Go loop wrote (tests/golden/tiny/tiny.ref.x), whole and chunked."""
    out, x = tmp_path / "sim.out", tmp_path / "sim.x"
    subprocess.check_call([hostsim_bin, os.path.join(GOLDEN, "tiny", "tiny.fa"), os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out), "-x", str(x)] + extra,
                          stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()
    assert x.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.x"), "rb").read()
