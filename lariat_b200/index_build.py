"""BWA-compatible index construction.

Produces, byte for byte, the five files `bwa index` writes for an N-free reference
(reference: go/src/gobwa/bwa/bwtindex.c:151-173,262-316, bwt.c:62-84,385-407, bntseq.c:66-96,275-328):

    <prefix>.pac   forward strand, 2 bits/base                .ann / .amb   contig table (text)
    <prefix>.bwt   BWT of forward+reverse-complement text with the occ counts interleaved every 128 symbols
    <prefix>.sa    suffix array sampled every 32 rows

`build_index` is the product path: `lc_index_build` of libcuda_lariat.so (lariat_b200/csrc/lc_index_build.{hpp,cu}) —
chunked suffix sorting with device-wide radix sorts, sized for the 3.1 Gb reference (6.2e9 suffixes) in one GPU's HBM.
It needs a GPU and fails loudly without one.

`build_index_torch` is a small-genome helper for the CPU test-suite only (prefix doubling on torch.sort, ~64 B/base):
tests that need an index for fresh synthetic data on a machine without a GPU use it; nothing in the product does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def build_index(prefix: str, contigs, device=0) -> dict:
    """contigs: list of (name, uint8 code array 0..3).  Writes <prefix>.{pac,ann,amb,bwt,sa} with the CUDA builder."""
    from . import gobwa
    lib = gobwa.load_library()
    if isinstance(device, str):      # "cuda:1" / "cuda"
        device = int(device.split(":")[1]) if ":" in device else 0
    arrs = [np.ascontiguousarray(c, dtype=np.uint8) for _, c in contigs]
    n = len(arrs)
    names = (C.c_char_p * n)(*[nm.encode() for nm, _ in contigs])
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * n)(*[a.size for a in arrs])
    rc = lib.lc_index_build(prefix.encode(), n, names, ptrs, lens, int(device))
    if rc != 0:
        raise RuntimeError(f"lc_index_build failed ({rc}): {gobwa.last_error()}")
    L = int(sum(a.size for a in arrs))
    return dict(l_pac=L, seq_len=2 * L)


def _suffix_array(T):
    """Ranks (1-based, unique) of all suffixes of T (uint8 codes 0..3), '$' < A."""
    import torch
    n = T.numel()
    dev = T.device
    k0 = 24 if n > 1 << 22 else 12
    # initial rank from the first k0 symbols, base-5 digits (0 = past the end)
    v = torch.zeros(n, dtype=torch.int64, device=dev)
    Tp = torch.zeros(n + k0, dtype=torch.int64, device=dev)
    Tp[:n] = T.to(torch.int64) + 1
    for j in range(k0):
        v.mul_(5).add_(Tp[j:j + n])
    del Tp

    def dense_rank(keys):
        s, idx = torch.sort(keys)
        flag = torch.ones(n, dtype=torch.int64, device=dev)
        flag[1:] = (s[1:] != s[:-1]).to(torch.int64)
        r = torch.cumsum(flag, 0)
        rank = torch.empty(n, dtype=torch.int64, device=dev)
        rank[idx] = r
        return rank, int(r[-1].item())

    rank, n_distinct = dense_rank(v)
    del v
    k = k0
    base = n + 2
    while n_distinct < n:
        nxt = torch.zeros(n, dtype=torch.int64, device=dev)
        if k < n:
            nxt[:n - k] = rank[k:]
        if base * base < (1 << 62):
            rank, n_distinct = dense_rank(rank * base + nxt)
        else:   # two stable passes when the pair does not fit one 64-bit key
            _, i1 = torch.sort(nxt, stable=True)
            _, i2 = torch.sort(rank[i1], stable=True)
            idx = i1[i2]
            a, b2 = rank[idx], nxt[idx]
            flag = torch.ones(n, dtype=torch.int64, device=dev)
            flag[1:] = ((a[1:] != a[:-1]) | (b2[1:] != b2[:-1])).to(torch.int64)
            r = torch.cumsum(flag, 0)
            rank = torch.empty(n, dtype=torch.int64, device=dev)
            rank[idx] = r
            n_distinct = int(r[-1].item())
        k <<= 1
    return rank


def build_index_torch(prefix: str, contigs, device: str = "cpu", sa_intv: int = 32) -> dict:
    """TEST HELPER (small genomes, CPU tensors): same files as build_index, by prefix doubling on torch.sort."""
    import torch
    fwd_np = np.concatenate([np.asarray(c, dtype=np.uint8) for _, c in contigs])
    if fwd_np.max(initial=0) > 3:
        raise ValueError("ambiguous bases are not supported by this builder (the reference randomises them with lrand48)")
    L = int(fwd_np.size)
    n = 2 * L
    # ---- .pac / .ann / .amb
    pad = (-L) % 4
    p4 = np.concatenate([fwd_np, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    pac = (p4[:, 0] << 6 | p4[:, 1] << 4 | p4[:, 2] << 2 | p4[:, 3]).astype(np.uint8)
    with open(prefix + ".pac", "wb") as f:
        f.write(pac.tobytes())
        if L % 4 == 0:
            f.write(b"\x00")
        f.write(bytes([L % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write(f"{L} {len(contigs)} 11\n")
        off = 0
        for name, c in contigs:
            f.write(f"0 {name} (null)\n{off} {len(c)} 0\n")
            off += len(c)
    with open(prefix + ".amb", "w") as f:
        f.write(f"{L} {len(contigs)} 0\n")
    # ---- suffix array of forward + reverse complement
    fwd = torch.from_numpy(fwd_np).to(device)
    T = torch.cat([fwd, 3 - fwd.flip(0)])
    del fwd
    rank = _suffix_array(T)                      # rank[i] in 1..n == BWT row of suffix i (row 0 is the empty suffix)
    primary = int(rank[0].item())
    sa_full = torch.empty(n + 1, dtype=torch.int64, device=T.device)
    sa_full[0] = n
    sa_full[rank] = torch.arange(n, dtype=torch.int64, device=T.device)
    del rank
    # ---- BWT: symbol preceding each suffix; the row of suffix 0 ('$') is dropped (B/bwtindex.c:84-121)
    keep = torch.ones(n + 1, dtype=torch.bool, device=T.device)
    keep[primary] = False
    B = T[(sa_full[keep] - 1)]                   # sa_full > 0 on every kept row
    L2 = torch.zeros(5, dtype=torch.int64)
    cnt = torch.bincount(T.to(torch.int64), minlength=4).cpu()
    L2[1:] = torch.cumsum(cnt, 0)
    # ---- SA samples (B/bwt.c:62-84,396-407)
    n_sa = (n + sa_intv) // sa_intv
    sa_s = sa_full[torch.arange(1, n_sa, device=T.device, dtype=torch.int64) * sa_intv].cpu().numpy().astype("<u8")
    del sa_full, keep
    with open(prefix + ".sa", "wb") as f:
        f.write(np.array([primary, *L2[1:].tolist(), sa_intv, n], dtype="<u8").tobytes())
        f.write(sa_s.tobytes())
    # ---- occ-interleaved .bwt (B/bwtindex.c:151-173): per 128 symbols {u64 counts-before[4]; u32 words[8]}, final counts at the end
    nb = (n + 127) // 128
    Bp = torch.full((nb * 128,), 4, dtype=torch.uint8, device=T.device)
    Bp[:n] = B
    del B, T
    blk = Bp.view(nb, 128)
    per = torch.stack([(blk == c).sum(1) for c in range(4)], 1).to(torch.int64)      # [nb, 4]
    before = torch.zeros(nb + 1, 4, dtype=torch.int64, device=Bp.device)
    before[1:] = torch.cumsum(per, 0)
    sym = (Bp & 3).view(-1, 16).to(torch.int64)                                       # padding packs as 0, like calloc
    sym[Bp.view(-1, 16) > 3] = 0
    shifts = torch.arange(15, -1, -1, device=Bp.device, dtype=torch.int64) * 2
    words = (sym << shifts).sum(1)                                                    # [nb*8] u32 values in int64
    n_words = (n + 15) >> 4
    out = torch.zeros(nb, 16, dtype=torch.int64, device=Bp.device)
    out[:, 0:8:2] = before[:nb] & 0xffffffff
    out[:, 1:8:2] = before[:nb] >> 32
    out[:, 8:] = words.view(nb, 8)
    flat = out.view(-1)
    # the last block holds only the words that exist; then the closing counts
    used = (nb - 1) * 16 + 8 + (n_words - (nb - 1) * 8)
    tail = torch.zeros(8, dtype=torch.int64, device=Bp.device)
    tail[0::2] = before[nb] & 0xffffffff
    tail[1::2] = before[nb] >> 32
    body = torch.cat([flat[:used], tail]).cpu().numpy().astype("<u4")
    with open(prefix + ".bwt", "wb") as f:
        f.write(np.array([primary, *L2[1:].tolist()], dtype="<u8").tobytes())
        f.write(body.tobytes())
    return dict(l_pac=L, seq_len=n, primary=primary)


def build_index_from_fasta(fa: str, device=0) -> dict:
    """Index an (N-free, ACGT) FASTA in place: <fa>.{pac,ann,amb,bwt,sa} like `bwa index <fa>` (CUDA builder)."""
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, v in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"Tt", 3)):
        for c in ch:
            lut[c] = v
    contigs, name, chunks = [], None, []
    with open(fa, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if name is not None:
                    contigs.append((name, np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)))
                name, chunks = line[1:].split()[0].decode(), []
            else:
                chunks.append(lut[np.frombuffer(line.rstrip(b"\r\n"), dtype=np.uint8)])
    if name is not None:
        contigs.append((name, np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)))
    return build_index(fa, contigs, device)
