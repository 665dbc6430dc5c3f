"""Synthetic references and barcode-sorted linked reads (BASELINE.json configs).

Everything is seeded numpy; nothing here touches the GPU or the oracle.  Reads are emitted both as
arrays (for the C-ABI) and in the 9-line barcode-sorted text format Lariat's reader consumes
(reference: go/src/fastqreader/reader.go:91-149 — name, R1 seq, R1 qual, R2 seq, R2 qual,
barcode seq, barcode qual, sample-index seq, sample-index qual), with qnames following the
simulated-read convention of go/check.py:55-59 (``mol:<bc>:<chr>:<molstart>:<molend>:<pos1>:<pos2>``).
"""
from __future__ import annotations

import gzip
import os
from dataclasses import dataclass, field

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


@dataclass
class GenomeSpec:
    """A synthetic reference: iid ACGT contigs plus optional repeat structure."""
    total_bp: int = 10_000_000
    n_contigs: int = 4
    seed: int = 12345
    segdup_frac: float = 0.0          # fraction of bases overwritten by diverged copies of other blocks
    segdup_len: tuple = (5_000, 50_000)
    segdup_div: float = 0.01
    # repeat families: list of (unit_len, n_copies, divergence_lo, divergence_hi)
    families: list = field(default_factory=list)
    n_tandem: int = 0                 # tandem arrays (period 40-170, 5-30 copies)
    name: str = "syn"


@dataclass
class ReadSpec:
    n_pairs: int = 20_000
    n_barcodes: int = 200
    read_len: int = 150
    mol_len: int = 50_000
    mols_per_barcode: int = 10
    insert_mean: float = 350.0
    insert_sd: float = 50.0
    insert_clip: tuple = (200, 700)
    sub_rate: float = 0.01
    indel_rate: float = 0.0005
    degraded_frac: float = 0.0        # fraction of pairs with one mate at >=10% error (forces mate rescue)
    degraded_err: float = 0.12
    repeat_bias: float = 0.0          # fraction of pairs whose molecule is placed on a repeat copy
    seed: int = 777


def make_genome(spec: GenomeSpec):
    """Return (list of (name, uint8 code array 0..3), list of repeat loci (contig, start, len))."""
    rng = np.random.default_rng(spec.seed)
    per = spec.total_bp // spec.n_contigs
    contigs = [rng.integers(0, 4, size=per, dtype=np.uint8) for _ in range(spec.n_contigs)]
    loci = []

    def mutate(block, div):
        block = block.copy()
        n = len(block)
        k = rng.binomial(n, div)
        if k:
            pos = rng.integers(0, n, size=k)
            block[pos] = (block[pos] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
        return block

    if spec.segdup_frac > 0:
        target = int(spec.total_bp * spec.segdup_frac)
        done = 0
        while done < target:
            L = int(rng.integers(spec.segdup_len[0], spec.segdup_len[1] + 1))
            sc, dc = rng.integers(0, spec.n_contigs, size=2)
            if L >= per // 2:
                L = per // 4
            s = int(rng.integers(0, per - L))
            d = int(rng.integers(0, per - L))
            blk = mutate(contigs[sc][s:s + L], spec.segdup_div)
            if rng.random() < 0.5:
                blk = (3 - blk[::-1]).astype(np.uint8)
            contigs[dc][d:d + L] = blk
            loci.append((int(dc), d, L))
            loci.append((int(sc), s, L))
            done += L
    for (ulen, ncopy, dlo, dhi) in spec.families:
        unit = rng.integers(0, 4, size=ulen, dtype=np.uint8)
        for _ in range(ncopy):
            c = int(rng.integers(0, spec.n_contigs))
            d = int(rng.integers(0, per - ulen))
            blk = mutate(unit, float(rng.uniform(dlo, dhi)))
            if rng.random() < 0.5:
                blk = (3 - blk[::-1]).astype(np.uint8)
            contigs[c][d:d + ulen] = blk
            loci.append((c, d, ulen))
    for _ in range(spec.n_tandem):
        period = int(rng.integers(40, 171))
        ncopy = int(rng.integers(5, 31))
        unit = rng.integers(0, 4, size=period, dtype=np.uint8)
        arr = np.concatenate([mutate(unit, 0.02) for _ in range(ncopy)])
        c = int(rng.integers(0, spec.n_contigs))
        d = int(rng.integers(0, per - len(arr)))
        contigs[c][d:d + len(arr)] = arr
        loci.append((c, d, len(arr)))
    named = [(f"{spec.name}{i + 1}", contigs[i]) for i in range(spec.n_contigs)]
    return named, loci


def write_fasta(path: str, contigs, width: int = 80) -> None:
    with open(path, "wb") as f:
        for name, codes in contigs:
            f.write(b">" + name.encode() + b"\n")
            s = _ACGT[codes]
            n = len(s)
            full = (n // width) * width
            if full:
                body = s[:full].reshape(-1, width)
                out = np.empty((body.shape[0], width + 1), dtype=np.uint8)
                out[:, :width] = body
                out[:, width] = 10
                f.write(out.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def simulate_pairs(contigs, loci, spec: ReadSpec, chunk: int = 100_000):
    """Simulate barcode-sorted linked read pairs (vectorised; chunked gathers).

    Returns dict with: r1, r2 (uint8 [n_pairs, read_len] codes 0..3), barcode_id [n_pairs] (sorted),
    barcodes (list of str), qname fields (contig, molstart, molend, pos1, pos2), strand.
    """
    rng = np.random.default_rng(spec.seed)
    n, L = spec.n_pairs, spec.read_len
    pad = 8
    G = np.concatenate([c for _, c in contigs])
    coff = np.concatenate([[0], np.cumsum([len(c) for _, c in contigs])]).astype(np.int64)
    clen = np.diff(coff)
    nb, mpb = spec.n_barcodes, spec.mols_per_barcode
    per_bc = -(-n // nb)
    # barcodes (sorted, unique)
    raw = rng.integers(0, 4, size=(int(nb * 1.2) + 8, 14))
    bcs = sorted({_ACGT[r].tobytes().decode() + "-1" for r in raw})[:nb]
    assert len(bcs) == nb
    # molecules
    m_c = rng.integers(0, len(contigs), size=(nb, mpb))
    mlen = np.minimum(spec.mol_len, clen[m_c] - 2 * pad - 2)
    m_s = pad + (rng.random((nb, mpb)) * (clen[m_c] - mlen - 2 * pad)).astype(np.int64)
    m_rep_s = np.full((nb, mpb), -1, dtype=np.int64)
    m_rep_l = np.zeros((nb, mpb), dtype=np.int64)
    if loci and spec.repeat_bias > 0:
        la = np.asarray(loci, dtype=np.int64)
        pick = rng.random((nb, mpb)) < spec.repeat_bias
        li = rng.integers(0, len(la), size=(nb, mpb))
        lc, ls, ll = la[li, 0], la[li, 1], la[li, 2]
        ms2 = np.clip(ls - (rng.random((nb, mpb)) * np.maximum(1, mlen - ll)).astype(np.int64), pad,
                      np.maximum(pad, clen[lc] - np.minimum(spec.mol_len, clen[lc] - 2 * pad - 2) - pad))
        m_c = np.where(pick, lc, m_c)
        mlen = np.minimum(spec.mol_len, clen[m_c] - 2 * pad - 2)
        m_s = np.where(pick, ms2, m_s)
        m_rep_s = np.where(pick, ls, -1)
        m_rep_l = np.where(pick, ll, 0)
    bc_id = np.minimum(np.arange(n) // per_bc, nb - 1).astype(np.int32)
    mi = rng.integers(0, mpb, size=n)
    c = m_c[bc_id, mi]
    ms = m_s[bc_id, mi]
    ml = mlen[bc_id, mi]
    ins = np.clip(rng.normal(spec.insert_mean, spec.insert_sd, size=n), *spec.insert_clip).astype(np.int64)
    ins = np.maximum(ins, L)
    ins = np.minimum(ins, ml - 1)
    p = ms + (rng.random(n) * np.maximum(1, ml - ins)).astype(np.int64)
    rs, rl = m_rep_s[bc_id, mi], m_rep_l[bc_id, mi]
    near = (rs >= 0) & (rng.random(n) < 0.8)
    lo = np.maximum(ms, rs - ins + L // 2)
    hi = np.minimum(ms + ml - ins, rs + rl - L // 2)
    p2 = lo + (rng.random(n) * np.maximum(1, hi - lo)).astype(np.int64)
    p = np.where(near & (hi > lo), p2, p)
    p = np.clip(p, pad, clen[c] - ins - pad - 1)
    gp = coff[c] + p
    r1 = np.empty((n, L), dtype=np.uint8)
    r2 = np.empty((n, L), dtype=np.uint8)
    deg = rng.random(n) < spec.degraded_frac
    which = rng.integers(0, 2, size=n)
    strand = (rng.random(n) < 0.5).astype(np.int8)
    ar = np.arange(L + pad, dtype=np.int64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        k = e - s
        fwd = G[gp[s:e, None] + ar[None, :]]
        rev = (3 - G[(gp[s:e] + ins[s:e] - 1)[:, None] - ar[None, :]]).astype(np.uint8)
        for t, w in ((fwd, 0), (rev, 1)):
            d = deg[s:e] & (which[s:e] == w)
            # small indels on a subset of reads (template has `pad` spare bases so length is kept)
            pind = np.where(d, 0.9, spec.indel_rate * L)
            for r in np.nonzero(rng.random(k) < pind)[0]:
                q = int(rng.integers(10, L - 10))
                g_ = int(rng.integers(1, 4))
                row = t[r]
                if rng.random() < 0.5:
                    row = np.concatenate([row[:q], row[q + g_:], row[:g_]])
                else:
                    row = np.concatenate([row[:q], rng.integers(0, 4, size=g_, dtype=np.uint8), row[q:]])[:L + pad]
                t[r] = row
            rate = np.where(d, spec.degraded_err, spec.sub_rate)
            m = rng.random((k, L)) < rate[:, None]
            t = t[:, :L].copy()
            t[m] = (t[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            if w == 0:
                a_ = t
            else:
                b_ = t
        sw = strand[s:e].astype(bool)
        r1[s:e] = np.where(sw[:, None], b_, a_)
        r2[s:e] = np.where(sw[:, None], a_, b_)
    return dict(r1=r1, r2=r2, barcode_id=bc_id, barcodes=bcs, contig=c, mol_start=ms, mol_end=ms + ml,
                pos1=p, pos2=p + ins - L, strand=strand, contig_names=[nm for nm, _ in contigs])


def qname(sim, i: int) -> str:
    """Simulated-read qname convention of go/check.py:55-59."""
    return (f"mol:{sim['barcodes'][int(sim['barcode_id'][i])]}:{sim['contig_names'][int(sim['contig'][i])]}:"
            f"{int(sim['mol_start'][i])}:{int(sim['mol_end'][i])}:{int(sim['pos1'][i])}:{int(sim['pos2'][i])}")


def write_9line(path: str, sim) -> None:
    """Write the barcode-sorted 9-line format (go/src/fastqreader/reader.go:91-149)."""
    r1, r2 = _ACGT[sim["r1"]], _ACGT[sim["r2"]]
    L = r1.shape[1]
    qual = b"I" * L
    with gzip.open(path, "wb", compresslevel=1) as f:
        for i in range(r1.shape[0]):
            bc = sim["barcodes"][int(sim["barcode_id"][i])]
            f.write(b"@" + qname(sim, i).encode() + b"\n" + r1[i].tobytes() + b"\n" + qual + b"\n" +
                    r2[i].tobytes() + b"\n" + qual + b"\n" + bc.encode() + b"\n" + b"I" * len(bc) + b"\n" +
                    b"ACGTACGT\n" + b"IIIIIIII\n")


def read_9line(path: str):
    """Minimal reader of the 9-line format -> (qnames, r1 strings, r2 strings, barcodes)."""
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    qn, a, b, bc = [], [], [], []
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 8, 9):
        qn.append(lines[i][1:].decode() if lines[i][:1] == b"@" else lines[i].decode())
        a.append(lines[i + 1])
        b.append(lines[i + 3])
        bc.append(lines[i + 5].decode())
    return qn, a, b, bc


# ---- the five BASELINE.json configurations -------------------------------------------------------

def config_specs(name: str):
    """(GenomeSpec, ReadSpec) for c1..c4 (c3/c4 scaled by env LARIAT_C3_BP when the full 3.1 Gb index is absent)."""
    if name == "c1":
        return GenomeSpec(10_000_000, 4, 12345, name="c1chr"), ReadSpec(20_000, 200, seed=777)
    if name == "c2":
        return (GenomeSpec(100_000_000, 8, 2345, segdup_frac=0.05, name="c2chr"),
                ReadSpec(1_000_000, 10_000, seed=778))
    if name in ("c3", "c4"):
        bp = int(os.environ.get("LARIAT_C3_BP", 3_100_000_000))
        scale = bp / 3_100_000_000
        g = GenomeSpec(bp, 24, 3345, segdup_frac=0.02,
                       families=[(300, max(1, int(1_000_000 * scale)), 0.10, 0.15),
                                 (6000, max(1, int(100_000 * scale)), 0.02, 0.08)],
                       n_tandem=max(1, int(20_000 * scale)), name="c3chr")
        if name == "c3":
            return g, ReadSpec(10_000_000, 100_000, seed=779)
        return g, ReadSpec(10_000_000, 100_000, mol_len=200_000, mols_per_barcode=50, degraded_frac=0.30,
                           repeat_bias=0.8, seed=780)
    if name == "tiny":   # used by the CPU tests: repeat-rich, rescue-heavy, seconds for the oracle
        g = GenomeSpec(400_000, 3, 99, segdup_frac=0.05, segdup_len=(500, 4000),
                       families=[(300, 120, 0.03, 0.10), (1500, 20, 0.01, 0.03)], n_tandem=20, name="tchr")
        return g, ReadSpec(1500, 15, mols_per_barcode=4, mol_len=20_000, degraded_frac=0.25, repeat_bias=0.6,
                           indel_rate=0.003, seed=5)
    raise KeyError(name)
