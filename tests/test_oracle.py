"""The oracle (oracle/port_*.c) pinned against the reference: golden dumps of the UNMODIFIED reference library,
per-function known answers generated from it, and the reference's own PhiX assertions (go/src/test/gobwa_test.go)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, have_ref
from lariat_b200 import iofmt

MAT = (C.c_int8 * 25)(*[(-1 if (i == 4 or j == 4) else (1 if i == j else -4)) for i in range(5) for j in range(5)])


def test_port_replay_matches_reference_golden(oracle_port, tmp_path):
    out = tmp_path / "port.out"
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), os.path.join(GOLDEN, "tiny", "tiny.fa"),
                           os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out)], stderr=subprocess.DEVNULL)
    assert out.read_bytes() == open(os.path.join(GOLDEN, "tiny", "tiny.ref.out"), "rb").read()


def test_getalignments_restatement_port_equals_reference_golden(oracle_port, tmp_path):
    """oracle/replay.c -x: the CIGAR walk of inference.GetAlignments (lariat.go:1573-1650) restated in C; the golden file was written by
    the build linked against the unmodified reference library, the port build must reproduce it."""
    out, x = tmp_path / "port.out", tmp_path / "port.x"
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), os.path.join(GOLDEN, "tiny", "tiny.fa"),
                           os.path.join(GOLDEN, "tiny", "tiny.pairs"), str(out), "-x", str(x)], stderr=subprocess.DEVNULL)
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.x"), "rb").read()
    assert x.read_bytes() == want
    recs = iofmt.parse_alnx(want)
    flat = [a for r in recs for a in r]
    assert len(flat) == 5684
    assert max(a["n_mm"] for a in flat) == 150          # the golden set has a lower-case read: the Go loop compares bytes
    assert sum(1 for a in flat if a["indels"]) > 20 and sum(1 for a in flat if a["soft_clipped"]) > 20
    assert all(a["pos"] < a["aend"] for a in flat)                   # after the reverse-strand swap (+1) of lariat.go:1645-1650


def test_golden_is_not_trivial():
    res = iofmt.read_results(os.path.join(GOLDEN, "tiny", "tiny.ref.out"))
    n = [len(r["regs"]) for r in res]
    assert len(res) == 2 * 605 and max(n) >= 10 and min(n) == 0
    rescued = sum(int((r["regs"]["secondary"] == -1).sum()) for r in res)   # mem_matesw marks rescued hits with secondary = -1
    gapped = sum(1 for r in res for a in r["aln"] if len(a[4]) > 2)
    assert rescued > 20 and gapped > 20


def _port(oracle_port):
    lib = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    return lib


def test_port_extend2_known_answers(oracle_port):
    lib = _port(oracle_port)
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for k in kat["extend"]:
        q = bytes(k["q"]); t = bytes(k["t"])
        o = [C.c_int() for _ in range(5)]
        sc = lib.port_extend2(len(q), q, len(t), t, 5, MAT, 6, 1, 6, 1, k["w"], 5, 100, k["h0"], *[C.byref(x) for x in o])
        assert [sc] + [x.value for x in o] == k["out"]


def test_port_sw_align2_known_answers(oracle_port):
    lib = _port(oracle_port)

    class swr(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")]
    lib.port_sw_align2.restype = swr
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for k in kat["align"]:
        ql, tl = len(k["q"]), len(k["t"])
        q = (C.c_uint8 * ql)(*k["q"]); t = (C.c_uint8 * tl)(*k["t"])
        xtra = 0x40000 | 0x80000 | (0x10000 if ql < 250 else 0) | 19
        r = lib.port_sw_align2(ql, q, tl, t, 5, MAT, 6, 1, 6, 1, xtra)
        assert [r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb] == k["out"]


def test_port_phix_known_answers(oracle_port, tmp_path):
    """go/src/test/gobwa_test.go:18-25: read 1 -> Offset 210 on contig PhiX; read 2 -> exactly one hit."""
    r1 = b"TCAAAAACTGACGCGTTGGATGAGGAGAAGTGGCTTAATATGCTTGGCACGTTCGTCAAGGACTGGTTTA"
    r2 = b"TATGACCAGTGTTTCCAGTCCGTTCAGTTGTTGCAGTGGAATAGTCAGGTTAAATTTAATGTGACCGCTT"
    pairs = tmp_path / "p.bin"
    iofmt.write_pairs(str(pairs), [r1, r2], [b"", b""])
    out = tmp_path / "o.bin"
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), os.path.join(GOLDEN, "phix", "PhiX.fa"), str(pairs), str(out)],
                          stderr=subprocess.DEVNULL)
    res = iofmt.read_results(str(out))
    a = res[0]["regs"][0]
    assert int(a["rb"]) == 210 and int(a["rid"]) == 0      # single contig at offset 0: Offset == rb
    assert open(os.path.join(GOLDEN, "phix", "PhiX.fa.ann")).read().split()[4] == "PhiX"
    assert len(res[2]["regs"]) == 1


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_equals_reference_on_fresh_data(oracle_port, tmp_path):
    """Fresh seeded data every run of the build container: port and unmodified reference must agree byte for byte."""
    from lariat_b200 import synth
    g, r = synth.config_specs("tiny")
    g.seed, r.seed, r.n_pairs, g.total_bp = 31337, 4242, 400, 300_000
    contigs, loci = synth.make_genome(g)
    fa = tmp_path / "x.fa"
    synth.write_fasta(str(fa), contigs)
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bwa"), "index", str(fa)], stderr=subprocess.DEVNULL)
    sim = synth.simulate_pairs(contigs, loci, r)
    iofmt.write_pairs(str(tmp_path / "x.pairs"), sim["r1"], sim["r2"])
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "replay_ref"), str(fa), str(tmp_path / "x.pairs"), str(tmp_path / "ref.out")],
                          stderr=subprocess.DEVNULL)
    subprocess.check_call([os.path.join(oracle_port, "replay_port"), str(fa), str(tmp_path / "x.pairs"), str(tmp_path / "port.out")],
                          stderr=subprocess.DEVNULL)
    assert (tmp_path / "ref.out").read_bytes() == (tmp_path / "port.out").read_bytes()
