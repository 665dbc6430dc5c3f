"""Parity tests proper: the CUDA path, called through the C ABI (ctypes over libcuda_lariat.so), against the oracle.
Bit-exact: candidate lists (all 88 bytes of every region, and their order), positions, strands, NM, CIGARs."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, have_ref
from lariat_b200 import gobwa, iofmt, synth

pytestmark = pytest.mark.gpu


def _load_pairs(path, n_use=None):
    raw = open(path, "rb").read()
    n = int(np.frombuffer(raw, dtype="<i4", count=1, offset=4)[0])
    n_use = n if n_use is None else min(n, n_use)
    lens = np.frombuffer(raw, dtype="<i4", count=2 * n, offset=8)[:2 * n_use]
    off = np.zeros(2 * n_use + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    bases = np.frombuffer(raw, dtype=np.uint8, offset=8 + 8 * n, count=int(off[-1]))
    return bases, off


@pytest.fixture(scope="module")
def tiny_ref(lib_path):
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "tiny", "tiny.fa"))
    assert ref is not None, gobwa.last_error()
    yield ref
    ref.close()


def test_golden_tiny_bit_exact(tiny_ref):
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"))
    br = tiny_ref.align_packed(bases, off, 25)
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()
    assert br.dump() == want
    assert br.gpu_launches >= 6


@pytest.mark.parametrize("env", [{"LC_DENSE_SA": "0"}, {"LC_SEED_LST_CAP": "10", "LC_SEED_PK_CAP": "6"}, {"LC_SEED_LST_CAP": "6", "LC_SEED_PK_CAP": "4", "LC_SEED_LST_CAP2": "9", "LC_SEED_PK_CAP2": "7", "LC_SEED_LST_CAP3": "12", "LC_SEED_PK_CAP3": "9"}, {"LC_DENSE_SA": "0", "LC_SEED_LST_CAP": "16"}, {"LC_FORCE_WIDE": "1"}, {"LC_EXT_HEAVY_SEEDS": "2", "LC_EXT_HEAVY_CHAINS": "1"}, {"LC_EXT_HEAVY_SEEDS": "0", "LC_EXT_HEAVY_CHAINS": "0"},
                                 {"LC_CHAIN_HEAVY": "3"}, {"LC_CHAIN_HEAVY": "40", "LC_FORCE_WIDE": "1"}, {"LC_CHAIN_HEAVY": "0"},
                                 {"LC_SMALL_READS": "4096"}, {"LC_SMALL_READS": "4096", "LC_FORCE_WIDE": "1", "LC_DENSE_SA": "0"}])
def test_golden_tiny_fallback_paths(env, lib_path, monkeypatch):
    """Same golden through the optional paths: sampled suffix array + LF walks instead of the dense array, and the
    one-thread-per-read redo kernel for reads that outgrow the state machine's fixed-size lists; LC_CHAIN_HEAVY: nearly every read
    through the warp-per-read chaining kernel (32 lanes sharing the filter's kept-chain scan), or none; LC_SMALL_READS: the low-latency
    route of small batches (warp-per-read seeding and extension for every read), which is the default for calls of <= 16384 reads."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "tiny", "tiny.fa"))
    assert ref is not None, gobwa.last_error()
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"))
    br = ref.align_packed(bases, off, 25)
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()
    assert br.dump() == want
    ref.close()


@pytest.mark.parametrize("chunk,lanes", [("40", "2"), ("97", "2"), ("0", "2"), ("61", "3"), ("33", "4"), ("50", "1")])
def test_golden_tiny_chunked_two_lanes(chunk, lanes, lib_path, monkeypatch):
    """Large calls are cut into chunks of pairs that rotate over LC_LANES streams (H2D / D2H of one chunk under the kernels of
    the others); the call's flat result arrays must be the same bytes.  LC_CHUNK_PAIRS=0 turns chunking off."""
    monkeypatch.setenv("LC_CHUNK_PAIRS", chunk)
    monkeypatch.setenv("LC_LANES", lanes)
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "tiny", "tiny.fa"))
    assert ref is not None, gobwa.last_error()
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"))
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()
    for _ in range(3):          # repeated calls reuse the pools and the pinned result arrays
        br = ref.align_packed(bases, off, 25)
        assert br.dump() == want
    assert int(br.reg_off[-1]) == len(br.regs)
    ref.close()


@pytest.mark.parametrize("env", [{"LC_CHUNK_TAPER": "1"}, {"LC_LATE_REGS": "1"}, {"LC_CHUNK_TAPER": "1", "LC_LATE_REGS": "1", "LC_LANES": "3"},
                                 {"LC_CHUNK_FIRST": "40"}, {"LC_CHUNK_FIRST": "100"}])
def test_golden_tiny_chunk_plans(env, lib_path, monkeypatch):
    """The other orders a chunked call can take: a smaller first chunk (LC_CHUNK_FIRST), a half-size first and last chunk (LC_CHUNK_TAPER), and a chunk whose predecessor is not
    through its pair stage yet, so that its CIGAR stage runs before its regions are packed and copied out (forced by LC_LATE_REGS)."""
    monkeypatch.setenv("LC_CHUNK_PAIRS", "70")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "tiny", "tiny.fa"))
    assert ref is not None, gobwa.last_error()
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"))
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()
    for _ in range(3):
        br = ref.align_packed(bases, off, 25)
        assert br.dump() == want
    ref.close()


def test_batch_split_invariance(tiny_ref):
    """Barcode groups are independent: aligning a sub-batch gives the same per-pair results as the big batch."""
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"))
    full = tiny_ref.align_packed(bases, off, 25)
    lo, hi = 100, 260
    sub = tiny_ref.align_packed(bases[off[2 * lo]:off[2 * hi]], off[2 * lo:2 * hi + 1] - off[2 * lo], 25)
    for r in range(2 * (hi - lo)):
        a = full.regs[full.reg_off[2 * lo + r]:full.reg_off[2 * lo + r + 1]]
        b = sub.regs[sub.reg_off[r]:sub.reg_off[r + 1]]
        assert a.tobytes() == b.tobytes()


def test_phix_known_answers(lib_path):
    """go/src/test/gobwa_test.go:11-26, through the gobwa mirror."""
    arena = gobwa.NewArena()
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "phix", "PhiX.fa"))
    assert ref is not None
    settings = gobwa.GoBwaAllocSettings()
    algns = gobwa.GoBwaAlign(ref, settings, "TCAAAAACTGACGCGTTGGATGAGGAGAAGTGGCTTAATATGCTTGGCACGTTCGTCAAGGACTGGTTTA", arena)
    assert algns[0].Offset == 210
    assert algns[0].Contig == "PhiX"
    algns = gobwa.GoBwaAlign(ref, settings, "TATGACCAGTGTTTCCAGTCCGTTCAGTTGTTGCAGTGGAATAGTCAGGTTAAATTTAATGTGACCGCTT", arena)
    assert len(algns) == 1
    names, lens = ref.GetReferenceContigsInfo()
    assert names == ["PhiX"] and lens == [5386]
    ref.close()


def test_zero_length_and_short_reads(tiny_ref):
    """go/src/test/lariat_test.go:12-24 (zero-length R1 after trimming must not crash) and reads below min_seed_len."""
    settings = gobwa.GoBwaAllocSettings()
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"), 4)
    r2 = bytes(bases[off[1]:off[2]])
    a1, a2 = gobwa.GoBwaMemMateSW(tiny_ref, settings, b"", r2, gobwa.NewArena(), 25)
    assert a1 == [] and len(a2) >= 1
    a1, a2 = gobwa.GoBwaMemMateSW(tiny_ref, settings, b"ACGTACGTAC", b"", gobwa.NewArena(), 25)
    assert a1 == [] and a2 == []
    out, br = gobwa.GoBwaMemMateSWBatch(tiny_ref, settings, [], 25)
    assert out == [] and br.n_reads == 0


def test_single_read_forms(tiny_ref):
    """lc_mem_reg2aln and lc_mem_matesw (source-compatible forms) agree with the batched call."""
    bases, off = _load_pairs(os.path.join(GOLDEN, "tiny", "tiny.pairs"), 60)
    br = tiny_ref.align_packed(bases, off, 25)
    settings = gobwa.GoBwaAllocSettings()
    checked = 0
    for r in range(br.n_reads):
        seq = bytes(bases[off[r]:off[r + 1]])
        for i in range(int(br.reg_off[r]), int(br.reg_off[r + 1])):
            s = gobwa.GoBwaSmithWaterman(tiny_ref, settings, seq, br.regs[i], gobwa.NewArena())
            a = br.aln[i]
            assert (s.Pos, s.Reversed, s.EditDistance) == (int(a["pos"]), bool(a["is_rev"]), int(a["NM"]))
            cig = br.cigar_of(i)
            assert s.Cigar == [v for c in cig for v in (int(c) & 0xf, int(c) >> 4)]
            checked += 1
            if checked > 40:
                return


def _fresh_dataset(tmp_path, name, gspec, rspec):
    contigs, loci = synth.make_genome(gspec)
    fa = str(tmp_path / f"{name}.fa")
    synth.write_fasta(fa, contigs)
    from lariat_b200 import index_build
    index_build.build_index(fa, contigs)
    sim = synth.simulate_pairs(contigs, loci, rspec)
    pairs = str(tmp_path / f"{name}.pairs")
    iofmt.write_pairs(pairs, sim["r1"], sim["r2"])
    return fa, pairs


def _oracle_replay(oracle_port):
    """The compiled reference's replay when it travelled to this box, else the plain-C port (both are the checker, never the product)."""
    ref = os.path.join(oracle_port, "_ref", "replay_ref")
    return ref if os.path.exists(ref) else os.path.join(oracle_port, "replay_port")


@pytest.mark.parametrize("kind", ["c1like", "repeat", "c2like", "c3like", "c4like"])
def test_fresh_synthetic_vs_oracle(kind, lib_path, oracle_port, tmp_path, monkeypatch):
    """Seeded synthetic data of every BASELINE config's shape (c1 uniform, c2 segmental duplications, c3 repeat families + tandem arrays
    on 24 contigs with 64-bit interval arithmetic forced as for the 3.1 Gb index, c4 long molecules / degraded mates / repeat-biased
    sampling) at sizes the oracle finishes in seconds, compared byte for byte with the oracle replay."""
    if kind == "c1like":
        g = synth.GenomeSpec(2_000_000, 4, 12345, name="c1chr")
        r = synth.ReadSpec(6000, 60, seed=777)
    elif kind == "repeat":
        g = synth.GenomeSpec(1_500_000, 3, 2024, segdup_frac=0.10, segdup_len=(1000, 8000), segdup_div=0.005,
                             families=[(300, 700, 0.03, 0.10), (2500, 40, 0.01, 0.03)], n_tandem=50, name="rchr")
        r = synth.ReadSpec(2500, 25, degraded_frac=0.3, repeat_bias=0.8, indel_rate=0.002, seed=9)
    elif kind == "c2like":
        g = synth.GenomeSpec(8_000_000, 8, 2345, segdup_frac=0.05, name="c2chr")
        r = synth.ReadSpec(20_000, 200, seed=778)
    else:
        scale = 12_000_000 / 3_100_000_000
        g = synth.GenomeSpec(12_000_000, 24, 3345, segdup_frac=0.02,
                             families=[(300, int(1_000_000 * scale), 0.10, 0.15), (6000, int(100_000 * scale), 0.02, 0.08)],
                             n_tandem=int(20_000 * scale), name="c3chr")
        if kind == "c3like":
            r = synth.ReadSpec(12_000, 120, seed=779)
            monkeypatch.setenv("LC_FORCE_WIDE", "1")
        else:
            r = synth.ReadSpec(8_000, 80, mol_len=200_000, mols_per_barcode=50, degraded_frac=0.30, repeat_bias=0.8, seed=780)
    if kind in ("repeat", "c4like"):
        monkeypatch.setenv("LC_CHAIN_HEAVY", "24")      # most repeat reads through the warp-per-read chaining kernel
    fa, pairs = _fresh_dataset(tmp_path, kind, g, r)
    want = str(tmp_path / "want.out")
    subprocess.check_call([_oracle_replay(oracle_port), fa, pairs, want, "-t", "8"], stderr=subprocess.DEVNULL)
    ref = gobwa.GoBwaLoadReference(fa)
    assert ref is not None, gobwa.last_error()
    bases, off = _load_pairs(pairs)
    br = ref.align_packed(bases, off, 25)
    assert br.dump() == open(want, "rb").read()
    ref.close()


def test_ksw_extend_batch_vs_oracle(tiny_ref, oracle_port):
    lib = gobwa.load_library()
    port = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["extend"]
    mat = (C.c_int8 * 25)(*[(-1 if (i == 4 or j == 4) else (1 if i == j else -4)) for i in range(5) for j in range(5)])
    for w in (100, 200, 5, 30):
        jobs = [k for k in kat if k["w"] == w]
        rng = np.random.default_rng(w)
        for _ in range(300):   # extra seeded jobs, answers from the oracle
            ql = int(rng.integers(1, 250)); tl = ql + int(rng.integers(0, 200))
            t = rng.integers(0, 4, size=tl, dtype=np.uint8); q = t[:ql].copy()
            m = rng.random(ql) < rng.choice([0, 0.01, 0.05, 0.1]); q[m] = rng.integers(0, 5, size=int(m.sum()), dtype=np.uint8)
            h0 = int(rng.integers(1, 150))
            o = [C.c_int() for _ in range(5)]
            sc = port.port_extend2(ql, q.tobytes(), tl, t.tobytes(), 5, mat, 6, 1, 6, 1, w, 5, 100, h0, *[C.byref(x) for x in o])
            jobs.append(dict(q=q.tolist(), t=t.tolist(), h0=h0, w=w, out=[sc] + [x.value for x in o]))
        qoff = np.zeros(len(jobs) + 1, dtype=np.int64); toff = np.zeros(len(jobs) + 1, dtype=np.int64)
        np.cumsum([len(j["q"]) for j in jobs], out=qoff[1:]); np.cumsum([len(j["t"]) for j in jobs], out=toff[1:])
        q = np.concatenate([np.array(j["q"], dtype=np.uint8) for j in jobs]); t = np.concatenate([np.array(j["t"], dtype=np.uint8) for j in jobs])
        h0 = np.array([j["h0"] for j in jobs], dtype=np.int32)
        out = np.zeros((len(jobs), 6), dtype=np.int32)
        ms = C.c_float(); cells = C.c_uint64()
        rc = lib.lc_ksw_extend_batch(tiny_ref.BWTData, len(jobs), q.ctypes.data, qoff.ctypes.data, t.ctypes.data, toff.ctypes.data,
                                     h0.ctypes.data, w, 5, 100, out.ctypes.data, C.byref(ms), C.byref(cells))
        assert rc == 0, gobwa.last_error()
        assert out.tolist() == [j["out"] for j in jobs]
        assert cells.value > 0


def test_ksw_align_batch_vs_reference_kat(tiny_ref):
    lib = gobwa.load_library()
    jobs = json.load(open(os.path.join(GOLDEN, "kat.json")))["align"]
    qoff = np.zeros(len(jobs) + 1, dtype=np.int64); toff = np.zeros(len(jobs) + 1, dtype=np.int64)
    np.cumsum([len(j["q"]) for j in jobs], out=qoff[1:]); np.cumsum([len(j["t"]) for j in jobs], out=toff[1:])
    q = np.concatenate([np.array(j["q"], dtype=np.uint8) for j in jobs]); t = np.concatenate([np.array(j["t"], dtype=np.uint8) for j in jobs])
    out = np.zeros((len(jobs), 7), dtype=np.int32)
    ms = C.c_float(); cells = C.c_uint64()
    rc = lib.lc_ksw_align_batch(tiny_ref.BWTData, len(jobs), q.ctypes.data, qoff.ctypes.data, t.ctypes.data, toff.ctypes.data,
                                out.ctypes.data, C.byref(ms), C.byref(cells))
    assert rc == 0, gobwa.last_error()
    assert out.tolist() == [j["out"] for j in jobs]


def test_sa_batch_is_a_permutation_consistent_with_oracle(tiny_ref, oracle_port):
    lib = gobwa.load_library()
    port = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    port.port_idx_load.restype = C.c_void_p
    port.port_sa.restype = C.c_uint64
    port.port_sa.argtypes = [C.c_void_p, C.c_uint64]
    ix = port.port_idx_load(os.path.join(GOLDEN, "tiny", "tiny.fa").encode())
    n = 2 * tiny_ref.l_pac
    rows = np.arange(1, n + 1, dtype=np.uint64)          # every BWT row but the sentinel row 0
    out = np.zeros(len(rows), dtype=np.uint64)
    ms = C.c_float()
    assert lib.lc_sa_batch(tiny_ref.BWTData, len(rows), rows.ctypes.data, out.ctypes.data, C.byref(ms)) == 0
    assert np.array_equal(np.sort(out), np.arange(n, dtype=np.uint64))      # SA is a permutation of 0..n-1
    rng = np.random.default_rng(1)
    for k in rng.integers(1, n + 1, size=2000):
        assert int(out[k - 1]) == port.port_sa(ix, int(k))


def test_get_seq_matches_oracle(tiny_ref, oracle_port):
    port = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    port.port_idx_load.restype = C.c_void_p
    port.port_get_seq.restype = C.POINTER(C.c_uint8)
    port.port_get_seq.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
    ix = port.port_idx_load(os.path.join(GOLDEN, "tiny", "tiny.fa").encode())
    rng = np.random.default_rng(3)
    for _ in range(50):
        tid = int(rng.integers(0, len(tiny_ref.names)))
        s = int(rng.integers(0, tiny_ref.lens[tid] - 300)); e = s + int(rng.integers(1, 300))
        ln = C.c_int64()
        p = port.port_get_seq(ix, tiny_ref.offsets[tid] + s, tiny_ref.offsets[tid] + e, C.byref(ln))
        want = bytes(b"ACGT"[p[i]] for i in range(ln.value))
        assert tiny_ref.GetSeq(tiny_ref.names[tid], s, e, False) == want
        comp = bytes(b"TGCA"[p[i]] for i in range(ln.value))[::-1]
        assert tiny_ref.GetSeq(tiny_ref.names[tid], s, e, True) == comp
