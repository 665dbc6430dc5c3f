"""The 9-line barcode-sorted FASTQ reader: lc_reader_* of libcuda_lariat.so (host-side C++, no GPU needed)
against oracle/reader_port.py, a plain-Python restatement of go/src/fastqreader/reader.go that is itself pinned by the reference's own
test assertions (go/src/test/fastq_reader_test.go) on the reference's own fixtures (tests/golden/reader/)."""
import ctypes as C
import gzip
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import reader_port  # noqa: E402

FQ1 = os.path.join(GOLDEN, "reader", "1.fq")
FQ0 = os.path.join(GOLDEN, "reader", "zero_length_read_test.fastq.gz")
NF = 9


class lc_readset_t(C.Structure):
    _fields_ = [("n_records", C.c_int32), ("complete", C.c_int32), ("bases", C.POINTER(C.c_uint8)), ("quals", C.POINTER(C.c_uint8)),
                ("off", C.POINTER(C.c_int64)), ("text", C.POINTER(C.c_char)), ("text_off", C.POINTER(C.c_int64))]


def product_sets(lib_path, path, trim):
    lib = C.CDLL(lib_path)
    lib.lc_reader_open.restype = C.c_void_p
    lib.lc_reader_open.argtypes = [C.c_char_p]
    lib.lc_reader_next_set.argtypes = [C.c_void_p, C.c_int, C.POINTER(lc_readset_t)]
    lib.lc_reader_close.argtypes = [C.c_void_p]
    rd = lib.lc_reader_open(path.encode())
    assert rd
    out = []
    while True:
        s = lc_readset_t()
        rc = lib.lc_reader_next_set(rd, trim, C.byref(s))
        assert rc >= 0
        if rc == 0:
            break
        n = s.n_records
        off = np.ctypeslib.as_array(s.off, shape=(2 * n + 1,)).copy()
        nb = int(off[-1])
        bases = bytes(np.ctypeslib.as_array(s.bases, shape=(max(nb, 1),))[:nb])
        quals = bytes(np.ctypeslib.as_array(s.quals, shape=(max(nb, 1),))[:nb])
        toff = np.ctypeslib.as_array(s.text_off, shape=(n * NF + 1,)).copy()
        text = C.string_at(s.text, int(toff[-1]))
        recs = []
        for i in range(n):
            f = [text[toff[i * NF + k]:toff[i * NF + k + 1]] for k in range(NF)]
            recs.append(dict(Read1=bases[off[2 * i]:off[2 * i + 1]], ReadQual1=quals[off[2 * i]:off[2 * i + 1]], Read2=bases[off[2 * i + 1]:off[2 * i + 2]],
                             ReadQual2=quals[off[2 * i + 1]:off[2 * i + 2]], ReadInfo=f[0], ReadGroupId=f[1], Barcode10X=f[2], RawBarcode10X=f[3],
                             Barcode10XQual=f[4], Barcode=f[5], BarcodeQual=f[6], TrimBases=f[7], TrimQuals=f[8]))
        out.append((recs, bool(s.complete)))
    lib.lc_reader_close(rd)
    return out


def oracle_sets(path, trim):
    r = reader_port.FastQReader(path)
    out = []
    while True:
        recs, err, full = r.ReadBarcodeSet(trim)
        if err is not None:
            break
        out.append(([{k: getattr(x, k) for k in reader_port.FastQRecord.__slots__} for x in recs], full))
    return out


def test_oracle_reader_reproduces_the_references_own_assertions():
    """go/src/test/fastq_reader_test.go:8-61, on the reference's fixtures."""
    a = reader_port.FastQReader(FQ1)
    for _ in range(3):
        ff = a.ReadOneLine(2)
    assert ff.Read1 == b"CCGCCCTAGCCAGGAGAGAAGCACTTCTTACCTGGGTTTCTTAGAGGCTTTGGCTGGCAATATTGTCAGCACCAGAGAGGACTTCTCGATGGCTGA"
    assert ff.ReadQual1 == b"BFFFFFFFFFFIIIIIFFIIIIIIIIFIIIIIFIFIFFIIFIIIIIIIIIIIIIIIFFFFFFFFFFFFFFFFFFFBFFFFFFFFFFFFFFFFFFFF"
    assert ff.Read2 == b"GTGGTAGTCTCCTGTTCAGCCATCGAGAAGTCCTCTCTGGTGCTGACAATATTGCCAGCCAAAGCCTCTAAGAAACCCAGGTAAGAAGTGCTTCTCTC"
    assert ff.ReadQual2 == b"BBBFFFFFFFFFFIIFIIFIIIIIIIIIIIIIIIIFIIIFFFIIIIIIIIIIIIIIIIIIIIFIIIIIIIIIIFFFFFFFFBBFFFFFBBFFFFFFFF"
    assert (ff.Barcode10X, ff.Barcode10XQual, ff.Barcode, ff.BarcodeQual) == (b"AAACAGAGAAAGAT", b"BBBFFFFFFFFFFI", b"CCGAACGC", b"BBBFFFFF")
    assert ff.ReadInfo == b"HWI-D00684:80:HFCKCADXX:2:2113:9410:56703"
    a = reader_port.FastQReader(FQ1)
    set1, _, _ = a.ReadBarcodeSet(2)
    assert set1[0].ReadInfo == b"HWI-D00684:80:HFCKCADXX:2:2113:17628:14813"
    assert set1[1].Read1 == b"CTGCTGCTCTCTCCATGTTTTTCCTGCACTCCTTGCAGGGACCTGAATAGCATGAACTGACTTTTCCTTGACGTAGTTGCTTCGTAGGATACTTCT"
    set2, _, _ = a.ReadBarcodeSet(2)
    assert set2[0].ReadInfo == b"HWI-D00684:80:HFCKCADXX:2:2112:14227:100270"
    assert set2[1].Read1 == b"CGGGCAGCAGCCATGGGATGCAGGACCTGCAGTCCACACATGTCACATGAATCTCCATGGAGAGGCACACAGTTCTCCCCATCTCAGCACTCTCTC"
    set1, _, _ = reader_port.FastQReader(FQ0).ReadBarcodeSet(7)
    assert len(set1) > 0 and any(len(x.Read1) == 0 for x in set1)      # R1 of length <= 7 trimmed to nothing (lariat_test.go:12-24)
    n = sum(len(s) for s, _ in oracle_sets(FQ1, 0))
    assert n == 666 and len(oracle_sets(FQ1, 0)) == 89                  # 666 records, 89 barcodes


@pytest.mark.parametrize("path,trim", [(FQ1, 0), (FQ1, 2), (FQ1, 200), (FQ0, 7)])
def test_product_reader_equals_oracle_on_reference_fixtures(lib_path, path, trim):
    assert product_sets(lib_path, path, trim) == oracle_sets(path, trim)


def test_product_reader_grouping_rules_on_synthetic_files(lib_path, tmp_path):
    """Set boundaries: barcode change, the 200-record chunks of non-whitelisted barcodes, the 'abnormal break' continuation rule,
    junk lines between records, read-group fields, comma-separated barcode lines, plain and gzip input, no trailing newline."""
    rng = np.random.default_rng(5)

    def rec(name, bc, l1=30, l2=25):
        s = lambda n: bytes(rng.choice(list(b"ACGTN"), size=n).astype(np.uint8))
        q = lambda n: bytes(rng.integers(33, 74, size=n).astype(np.uint8))
        return b"@" + name + b"\n" + s(l1) + b"\n" + q(l1) + b"\n" + s(l2) + b"\n" + q(l2) + b"\n" + bc + b"\n" + q(len(bc)) + b"\nACGTACGT\nIIIIIIII\n"

    lines = []
    for i in range(450):      # a whitelisted barcode spread over more than 200 records: continuation sets
        lines.append(rec(b"r%d some RG:x:y" % i, b"AAAACCCCGGGGTT-1,AAAACCCCGGGGTA"))
    lines.append(b"junk line that is not a record\n")
    for i in range(520):      # not on the whitelist (no '-'): chunks of 200
        lines.append(rec(b"n%d" % i, b"ACGTACGTACGTAC", l1=5))
    for b_ in range(40):
        for i in range(int(rng.integers(1, 9))):
            lines.append(rec(b"x%d_%d\tfoo bar" % (b_, i), b"C%013d-1" % b_, l1=int(rng.integers(0, 12))))
    data = b"".join(lines)
    for name, payload in (("a.fq", data), ("a.fq.gz", gzip.compress(data)), ("trunc.fq", data[:-1]), ("trunc2.fq", data[:len(data) // 2])):
        p = str(tmp_path / name)
        open(p, "wb").write(payload)
        for trim in (0, 7):
            got, want = product_sets(lib_path, p, trim), oracle_sets(p, trim)
            assert got == want, (name, trim)
    sets = product_sets(lib_path, str(tmp_path / "a.fq"), 0)
    sizes = [len(s) for s, _ in sets]
    assert sizes[:3] == [450, 200, 200] and sizes[3] == 120 and len(sets) == 3 + 1 + 40


def test_reader_feeds_lc_align_packed_layout(lib_path):
    """The packed arrays of a set are what lc_align_packed takes: read 2i / 2i+1 of record i, offsets non-decreasing, trimmed R1."""
    for recs, _ in product_sets(lib_path, FQ1, 2)[:5]:
        assert all(len(r["Read1"]) == 96 and len(r["TrimBases"]) == 2 for r in recs)
