"""Small binary formats shared with the checker-side replay harnesses, tests and bench.

pairs file ("LRP1"): int32 n_pairs, int32 len[2n] (r1,r2 interleaved), ASCII bases back to back.
alnx file ("LRX1"): int32 n_pairs, then per read end: int32 n; n x {int64 pos, aend; int32 matches, mismatches, indels,
    indel_length, soft_clipped, soft_clipped_length, n_mm; int32 mismatchLocs[n_mm]; int32 mismatchReadLocs[n_mm]}
    (what inference.GetAlignments computes per candidate after mem_reg2aln, lariat.go:1573-1650).
result file ("LRO1"): int32 n_pairs, then per read end: int32 n; n x 88-byte alnreg; n x
    {int64 pos; int32 rid, is_rev, NM, n_cigar; uint32 cigar[n_cigar]}.
"""
from __future__ import annotations

import numpy as np

_ACGTN = np.frombuffer(b"ACGTN", dtype=np.uint8)

# layout of mem_alnreg_t (reference: go/src/gobwa/bwa/bwamem.h:59-77; sizeof == 88)
ALNREG_DTYPE = np.dtype([
    ("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
    ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"),
    ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"),
    ("n_comp_is_alt", "<u4"), ("frac_rep", "<f4"), ("hash", "<u8")])
assert ALNREG_DTYPE.itemsize == 88


def write_pairs(path, r1, r2):
    """r1/r2: uint8 code arrays [n, L] (0..4) or lists of ASCII bytes (variable length)."""
    with open(path, "wb") as f:
        if isinstance(r1, np.ndarray):
            n, L1 = r1.shape
            L2 = r2.shape[1]
            f.write(b"LRP1" + np.int32(n).tobytes())
            lens = np.empty(2 * n, dtype=np.int32)
            lens[0::2], lens[1::2] = L1, L2
            f.write(lens.tobytes())
            both = np.concatenate([_ACGTN[r1], _ACGTN[r2]], axis=1)
            f.write(both.tobytes())
        else:
            n = len(r1)
            f.write(b"LRP1" + np.int32(n).tobytes())
            lens = np.empty(2 * n, dtype=np.int32)
            lens[0::2] = [len(x) for x in r1]
            lens[1::2] = [len(x) for x in r2]
            f.write(lens.tobytes())
            f.write(b"".join(a + b for a, b in zip(r1, r2)))


def read_results(path):
    """-> list (len 2*n_pairs) of dicts {regs: structured array, aln: list of (pos, rid, is_rev, NM, cigar ndarray)}."""
    buf = open(path, "rb").read()
    return parse_results(buf)


def parse_results(buf):
    assert buf[:4] == b"LRO1"
    n = int(np.frombuffer(buf, dtype="<i4", count=1, offset=4)[0])
    off = 8
    out = []
    for _ in range(2 * n):
        k = int(np.frombuffer(buf, dtype="<i4", count=1, offset=off)[0])
        off += 4
        regs = np.frombuffer(buf, dtype=ALNREG_DTYPE, count=k, offset=off)
        off += 88 * k
        alns = []
        for _j in range(k):
            pos = int(np.frombuffer(buf, dtype="<i8", count=1, offset=off)[0])
            rid, is_rev, nm, nc = (int(x) for x in np.frombuffer(buf, dtype="<i4", count=4, offset=off + 8))
            off += 24
            cig = np.frombuffer(buf, dtype="<u4", count=nc, offset=off)
            off += 4 * nc
            alns.append((pos, rid, is_rev, nm, cig))
        out.append(dict(regs=regs, aln=alns))
    assert off == len(buf)
    return out


def cigar_str(cig):
    return "".join(f"{int(c) >> 4}{'MIDSH'[int(c) & 15]}" for c in cig)


def parse_alnx(buf):
    """-> list (len 2*n_pairs) of lists of dicts, one per candidate (LRX1)."""
    assert buf[:4] == b"LRX1"
    n = int(np.frombuffer(buf, dtype="<i4", count=1, offset=4)[0])
    off = 8
    out = []
    names = ("matches", "mismatches", "indels", "indel_length", "soft_clipped", "soft_clipped_length", "n_mm")
    for _ in range(2 * n):
        k = int(np.frombuffer(buf, dtype="<i4", count=1, offset=off)[0])
        off += 4
        recs = []
        for _j in range(k):
            pos, aend = (int(x) for x in np.frombuffer(buf, dtype="<i8", count=2, offset=off))
            h = [int(x) for x in np.frombuffer(buf, dtype="<i4", count=7, offset=off + 16)]
            off += 44
            nm = h[6]
            d = dict(zip(names, h), pos=pos, aend=aend, mm_ref=np.frombuffer(buf, dtype="<i4", count=nm, offset=off).tolist(),
                     mm_read=np.frombuffer(buf, dtype="<i4", count=nm, offset=off + 4 * nm).tolist())
            off += 8 * nm
            recs.append(d)
        out.append(recs)
    assert off == len(buf)
    return out
