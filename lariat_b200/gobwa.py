"""Host-side mirror of Lariat's cgo bridge (reference: go/src/gobwa/gobwa.go) over libcuda_lariat.so.

The reference's host language is Go, which this build image does not have, so the host layer above the C ABI
is written in Python with the SAME names, argument meaning and error behaviour as gobwa.go:

    GoBwaLoadReference(path)          gobwa.go:128   -> GoBwaReference (None + log line on failure, like the reference)
    GoBwaAllocSettings()              gobwa.go:149
    GoBwaAlign(ref, settings, seq, arena)                                   gobwa.go:174
    GoBwaMemMateSW(ref, settings, read1, read2, arena, score_delta)         gobwa.go:226
    GoBwaSmithWaterman(ref, settings, seq, alignment, arena)                gobwa.go:400
    GoBwaReference.GetSeq / GetReferenceContigsInfo, EnumerateContigs       gobwa.go:27,50,435
    EasyAlignment, SingleReadAlignment, Arena                               gobwa.go:88,417,107

plus the batched form the GPU wants — one call per barcode group (or many groups):

    GoBwaMemMateSWBatch(ref, settings, pairs, score_delta) -> [(algns_read1, algns_read2), ...]
    BatchResult (flat numpy views of the lc_result_t arrays, zero per-read Python work)

There is NO CPU fallback here: if libcuda_lariat.so is missing or no GPU is present these calls raise / return None.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcuda_lariat.so")

ALNREG_DTYPE = np.dtype([
    ("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
    ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"),
    ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"),
    ("n_comp_is_alt", "<u4"), ("frac_rep", "<f4"), ("hash", "<u8")])
ALN_DTYPE = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("is_rev", "<i4"), ("NM", "<i4"), ("n_cigar", "<i4"),
                      ("cigar_off", "<u4"), ("n_mismatch", "<i4")])
ALNX_DTYPE = np.dtype([("pos", "<i8"), ("aend", "<i8"), ("matches", "<i4"), ("mismatches", "<i4"), ("indels", "<i4"), ("indel_length", "<i4"),
                       ("soft_clipped", "<i4"), ("soft_clipped_length", "<i4"), ("n_mm", "<i4"), ("mm_off", "<u4")])
assert ALNREG_DTYPE.itemsize == 88 and ALN_DTYPE.itemsize == 32 and ALNX_DTYPE.itemsize == 48


class lc_opt_t(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "pen_unpaired", "pen_clip5",
                                        "pen_clip3", "w", "zdrop")] + [("max_mem_intv", C.c_uint64)] + \
               [(n, C.c_int) for n in ("T", "flag", "min_seed_len", "min_chain_weight", "max_chain_extend")] + \
               [("split_factor", C.c_float)] + \
               [(n, C.c_int) for n in ("split_width", "max_occ", "max_chain_gap", "n_threads", "chunk_size")] + \
               [(n, C.c_float) for n in ("mask_level", "drop_ratio", "XA_drop_ratio", "mask_level_redun", "mapQ_coef_len")] + \
               [(n, C.c_int) for n in ("mapQ_coef_fac", "max_ins", "max_matesw", "max_XA_hits", "max_XA_hits_alt")] + \
               [("mat", C.c_int8 * 25)]


class lc_contigs_t(C.Structure):
    _fields_ = [("n_seqs", C.c_int32), ("l_pac", C.c_int64), ("names", C.POINTER(C.c_char_p)),
                ("lens", C.POINTER(C.c_int32)), ("offsets", C.POINTER(C.c_int64))]


class lc_result_t(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("n_regs", C.c_int64), ("reg_off", C.POINTER(C.c_int64)),
                ("regs", C.c_void_p), ("aln", C.c_void_p), ("cigars", C.POINTER(C.c_uint32)), ("n_cigar_words", C.c_int64)] + \
               [(n, C.c_float) for n in ("ms_h2d", "ms_seed", "ms_sa", "ms_chain", "ms_extend", "ms_dedup", "ms_pair",
                                         "ms_cigar", "ms_d2h", "ms_total")] + \
               [("gpu_launches", C.c_int64)] + \
               [(n, C.c_uint64) for n in ("cnt_occ_blocks", "cnt_extend_calls", "cnt_sa_calls", "cnt_lf_steps", "cnt_ext_calls",
                                          "cnt_ext_cells", "cnt_sw_calls", "cnt_sw_cells", "cnt_glb_calls", "cnt_glb_cells")] + \
               [("alnx", C.c_void_p), ("mm_ref_pos", C.POINTER(C.c_int32)), ("mm_read_pos", C.POINTER(C.c_int32)), ("n_mm", C.c_int64)]


class lc_pestat_t(C.Structure):
    _fields_ = [("low", C.c_int), ("high", C.c_int), ("failed", C.c_int), ("avg", C.c_double), ("std", C.c_double)]


class lc_alnreg_v(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


# every symbol include/cuda_lariat.h declares
EXPORTS = ["lc_opt_init", "lc_init", "lc_clone", "lc_set_coalescing", "lc_destroy", "lc_last_error", "lc_contigs", "lc_align_pairs", "lc_align_packed",
           "lc_batch_upload", "lc_batch_run", "lc_batch_fetch", "lc_free_result", "lc_fetch_seq", "lc_mem_align1_core",
           "lc_mem_matesw", "lc_mem_reg2aln", "lc_ksw_extend_batch", "lc_ksw_align_batch", "lc_collect_intv_batch",
           "lc_sa_batch", "lc_index_build", "lc_reader_open", "lc_reader_next_set", "lc_reader_close", "lc_set_counting", "lc_set_alnx", "lc_set_chunk_pairs", "lc_build_info", "lc_timer_start", "lc_timer_stop"]

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen libcuda_lariat.so and declare prototypes.  Raises if the library is missing: the product has no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32p, i64p, u8p, u64p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_float)
    lib.lc_opt_init.argtypes = [C.POINTER(lc_opt_t)]; lib.lc_opt_init.restype = None
    lib.lc_init.argtypes = [C.c_char_p, C.c_int, C.POINTER(lc_opt_t)]; lib.lc_init.restype = vp
    lib.lc_destroy.argtypes = [vp]; lib.lc_destroy.restype = None
    lib.lc_clone.argtypes = [vp]; lib.lc_clone.restype = vp
    lib.lc_set_coalescing.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.lc_last_error.argtypes = []; lib.lc_last_error.restype = C.c_char_p
    lib.lc_build_info.argtypes = []; lib.lc_build_info.restype = C.c_char_p
    lib.lc_contigs.argtypes = [vp]; lib.lc_contigs.restype = C.POINTER(lc_contigs_t)
    lib.lc_align_pairs.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), i32p, C.POINTER(C.c_char_p), i32p, C.c_int,
                                   C.POINTER(C.POINTER(lc_result_t))]
    lib.lc_align_packed.argtypes = [vp, C.c_int, u8p, vp, C.c_int, C.POINTER(C.POINTER(lc_result_t))]
    lib.lc_batch_upload.argtypes = [vp, C.c_int, u8p, vp]
    lib.lc_batch_run.argtypes = [vp, C.c_int]
    lib.lc_batch_fetch.argtypes = [vp, C.POINTER(C.POINTER(lc_result_t))]
    lib.lc_free_result.argtypes = [vp, C.POINTER(lc_result_t)]; lib.lc_free_result.restype = None
    lib.lc_fetch_seq.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, vp]; lib.lc_fetch_seq.restype = C.c_int64
    lib.lc_mem_align1_core.argtypes = [vp, C.c_int, vp, C.POINTER(lc_alnreg_v)]
    lib.lc_mem_matesw.argtypes = [vp, C.POINTER(lc_pestat_t), vp, C.c_int, vp, C.POINTER(lc_alnreg_v)]
    lib.lc_mem_reg2aln.argtypes = [vp, C.c_int, vp, vp, vp, C.POINTER(C.POINTER(C.c_uint32))]
    lib.lc_ksw_extend_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, f32p, u64p]
    lib.lc_ksw_align_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, f32p, u64p]
    lib.lc_collect_intv_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int64, f32p]; lib.lc_collect_intv_batch.restype = C.c_int64
    lib.lc_sa_batch.argtypes = [vp, C.c_int64, vp, vp, f32p]
    lib.lc_index_build.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), i64p, C.c_int]
    lib.lc_set_counting.argtypes = [vp, C.c_int]; lib.lc_set_counting.restype = None
    lib.lc_set_alnx.argtypes = [vp, C.c_int]; lib.lc_set_alnx.restype = None
    lib.lc_set_chunk_pairs.argtypes = [vp, C.c_int]; lib.lc_set_chunk_pairs.restype = C.c_int
    lib.lc_timer_start.argtypes = [vp]
    lib.lc_timer_stop.argtypes = [vp, f32p]
    _lib = lib
    return lib


def last_error() -> str:
    return load_library().lc_last_error().decode()


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


# ---------------------------------------------------------------------------------------------------------------
@dataclass
class EasyAlignment:            # gobwa.go:88-98
    Offset: int = 0
    Alignment_end: int = 0
    Contig: str = ""
    Reversed: bool = False
    ChainedHit: object = None   # the 88-byte region record (numpy void of ALNREG_DTYPE) — *C.mem_alnreg_t in the reference
    Score: int = 0
    Secondary: bool = False
    ReadS: int = 0
    ReadE: int = 0


@dataclass
class SingleReadAlignment:      # gobwa.go:417-433
    Pos: int = 0
    Chrom: str = ""
    Flag: int = 0
    Reversed: bool = False
    Alt: int = 0
    Mapq: int = 0
    EditDistance: int = 0
    Cigar: list = field(default_factory=list)     # [op0, len0, op1, len1, ...] like gobwa.go:464-471
    AlternativeMappings: str = ""
    Score: int = 0
    Sub: int = 0
    AltSC: int = 0
    ReadS: int = 0
    ReadE: int = 0


class Arena:                    # gobwa.go:107-126
    """The reference's Arena collects the C arrays BWA malloc'd for one barcode and C.free()s them at the end of
    DoRFAForOneBarcode (lariat.go:494,546).  Here it collects the result handles of the batched calls: the arrays a
    BatchResult(copy=False) views stay valid — whatever other threads or later calls do — until Free() gives them back
    to the library (lc_free_result)."""

    def __init__(self):
        self.Pointers = []

    def Push(self, p):
        self.Pointers.append(p)

    def Free(self):
        for p in self.Pointers:
            if hasattr(p, "free"):
                p.free()
        self.Pointers = []


def NewArena() -> Arena:
    return Arena()


class GoBwaSettings:            # gobwa.go:45-47
    def __init__(self):
        self.Settings = lc_opt_t()
        load_library().lc_opt_init(C.byref(self.Settings))


def GoBwaAllocSettings() -> GoBwaSettings:
    return GoBwaSettings()


class BatchResult:
    """Flat views of one lc_result_t (copied out of the library's buffers so they survive the next call)."""

    def __init__(self, res: lc_result_t, copy: bool = True, owner=None):
        """owner = (lib, ctx, pointer): the library-side block these arrays live in; free() (or Arena.Free) releases it.
        With copy=True the arrays are numpy-owned copies and the block is released right away."""
        self._owner = owner
        n = int(res.n_reads)
        nr = int(res.n_regs)
        self.n_reads = n
        self.reg_off = np.ctypeslib.as_array(res.reg_off, shape=(n + 1,))
        if nr:
            self.regs = np.frombuffer((C.c_char * (88 * nr)).from_address(res.regs), dtype=ALNREG_DTYPE)
            self.aln = np.frombuffer((C.c_char * (32 * nr)).from_address(res.aln), dtype=ALN_DTYPE)
        else:
            self.regs = np.zeros(0, dtype=ALNREG_DTYPE)
            self.aln = np.zeros(0, dtype=ALN_DTYPE)
        nc = int(res.n_cigar_words)
        self.cigars = np.ctypeslib.as_array(res.cigars, shape=(nc,)) if nc else np.zeros(0, dtype=np.uint32)
        # GetAlignments post-processing (lc_set_alnx): per candidate counts + the two mismatch position lists
        self.alnx = self.mm_ref = self.mm_read = None
        if res.alnx:
            self.alnx = np.frombuffer((C.c_char * (48 * nr)).from_address(res.alnx), dtype=ALNX_DTYPE) if nr else np.zeros(0, dtype=ALNX_DTYPE)
            nm = int(res.n_mm)
            self.mm_ref = np.ctypeslib.as_array(res.mm_ref_pos, shape=(nm,)) if nm else np.zeros(0, dtype=np.int32)
            self.mm_read = np.ctypeslib.as_array(res.mm_read_pos, shape=(nm,)) if nm else np.zeros(0, dtype=np.int32)
        if copy:
            self.reg_off, self.regs, self.aln, self.cigars = self.reg_off.copy(), self.regs.copy(), self.aln.copy(), self.cigars.copy()
            if self.alnx is not None:
                self.alnx, self.mm_ref, self.mm_read = self.alnx.copy(), self.mm_ref.copy(), self.mm_read.copy()
            self.free()
        self.timing = {k: float(getattr(res, k)) for k in ("ms_h2d", "ms_seed", "ms_sa", "ms_chain", "ms_extend", "ms_dedup",
                                                           "ms_pair", "ms_cigar", "ms_d2h", "ms_total")}
        self.gpu_launches = int(res.gpu_launches)
        self.counters = {k[4:]: int(getattr(res, k)) for k in ("cnt_occ_blocks", "cnt_extend_calls", "cnt_sa_calls", "cnt_lf_steps",
                                                               "cnt_ext_calls", "cnt_ext_cells", "cnt_sw_calls", "cnt_sw_cells",
                                                               "cnt_glb_calls", "cnt_glb_cells")}

    def free(self):
        """lc_free_result: the caller owns a result until it says so (views made with copy=False are dead afterwards)."""
        if self._owner is not None:
            lib, ctx, ptr = self._owner
            lib.lc_free_result(ctx, ptr)
            self._owner = None

    def cigar_of(self, i: int) -> np.ndarray:
        a = self.aln[i]
        return self.cigars[int(a["cigar_off"]):int(a["cigar_off"]) + int(a["n_cigar"])]

    def dump(self) -> bytes:
        """Serialise in the LRO1 layout (lariat_b200/iofmt.py) so dumps are byte-comparable with the checker side."""
        out = [b"LRO1", np.int32(self.n_reads // 2).tobytes()]
        for r in range(self.n_reads):
            s, e = int(self.reg_off[r]), int(self.reg_off[r + 1])
            out.append(np.int32(e - s).tobytes())
            out.append(self.regs[s:e].tobytes())
            for i in range(s, e):
                a = self.aln[i]
                out.append(np.int64(a["pos"]).tobytes())
                out.append(np.array([a["rid"], a["is_rev"], a["NM"], a["n_cigar"]], dtype="<i4").tobytes())
                out.append(self.cigar_of(i).astype("<u4").tobytes())
        return b"".join(out)


    def dump_alnx(self) -> bytes:
        """The GetAlignments records in the LRX1 layout (lariat_b200/iofmt.py), byte-comparable with `replay_* -x`."""
        out = [b"LRX1", np.int32(self.n_reads // 2).tobytes()]
        for r in range(self.n_reads):
            s, e = int(self.reg_off[r]), int(self.reg_off[r + 1])
            out.append(np.int32(e - s).tobytes())
            for i in range(s, e):
                a = self.alnx[i]
                out.append(np.array([a["pos"], a["aend"]], dtype="<i8").tobytes())
                out.append(np.array([a[k] for k in ("matches", "mismatches", "indels", "indel_length", "soft_clipped", "soft_clipped_length", "n_mm")], dtype="<i4").tobytes())
                o, n = int(a["mm_off"]), int(a["n_mm"])
                out.append(self.mm_ref[o:o + n].astype("<i4").tobytes())
                out.append(self.mm_read[o:o + n].astype("<i4").tobytes())
        return b"".join(out)

    def dump_fast(self, n_pairs: int | None = None) -> bytes:
        """dump() of the first n_pairs pairs without per-read Python work (numpy scatter): bench.py verifies 10^5-pair prefixes with it."""
        R = self.n_reads if n_pairs is None else min(self.n_reads, 2 * n_pairs)
        ro = np.asarray(self.reg_off[:R + 1], dtype=np.int64)
        c0, c1 = int(ro[0]), int(ro[R])
        k = np.diff(ro)                                           # candidates per read
        aln = self.aln[c0:c1]
        nc = aln["n_cigar"].astype(np.int64)
        cand_sz = 24 + 4 * nc
        csum = np.concatenate([[0], np.cumsum(cand_sz)])
        read_sz = 4 + 88 * k + (csum[ro[1:] - c0] - csum[ro[:-1] - c0])
        read_off = 8 + np.concatenate([[0], np.cumsum(read_sz)])
        buf = np.zeros(int(read_off[-1]), dtype=np.uint8)
        buf[:4] = np.frombuffer(b"LRO1", dtype=np.uint8)
        buf[4:8] = np.frombuffer(np.int32(R // 2).tobytes(), dtype=np.uint8)
        hdr = read_off[:-1, None] + np.arange(4)[None, :]
        buf[hdr.reshape(-1)] = k.astype("<i4").view(np.uint8)
        n = c1 - c0
        if n:
            rd = np.repeat(np.arange(R), k)                       # read of each candidate
            j = np.arange(n) - (ro[rd] - c0)                      # its index inside the read
            reg_dst = read_off[rd] + 4 + 88 * j
            buf[(reg_dst[:, None] + np.arange(88)[None, :]).reshape(-1)] = self.regs[c0:c1].view(np.uint8).reshape(-1)
            aln_dst = read_off[rd] + 4 + 88 * k[rd] + (csum[:-1] - csum[ro[rd] - c0])
            fixed = np.zeros((n, 24), dtype=np.uint8)
            fixed[:, :8] = aln["pos"].astype("<i8").view(np.uint8).reshape(n, 8)
            for t, f in enumerate(("rid", "is_rev", "NM", "n_cigar")):
                fixed[:, 8 + 4 * t:12 + 4 * t] = aln[f].astype("<i4").view(np.uint8).reshape(n, 4)
            buf[(aln_dst[:, None] + np.arange(24)[None, :]).reshape(-1)] = fixed.reshape(-1)
            tot = int(nc.sum())
            if tot:
                cand = np.repeat(np.arange(n), nc)
                t = np.arange(tot) - np.repeat(np.concatenate([[0], np.cumsum(nc)])[:-1], nc)
                src = aln["cigar_off"].astype(np.int64)[cand] + t
                dst = aln_dst[cand] + 24 + 4 * t
                buf[(dst[:, None] + np.arange(4)[None, :]).reshape(-1)] = self.cigars[src].astype("<u4").view(np.uint8)
        return buf.tobytes()


class GoBwaReference:           # gobwa.go:21-24
    def __init__(self, ctx, lib):
        self.BWTData = ctx          # opaque lc_ctx* (the reference keeps a *bwaidx_t here)
        self._lib = lib
        ct = lib.lc_contigs(ctx).contents
        self.l_pac = int(ct.l_pac)
        self.names = [ct.names[i].decode() for i in range(ct.n_seqs)]
        self.lens = [int(ct.lens[i]) for i in range(ct.n_seqs)]
        self.offsets = [int(ct.offsets[i]) for i in range(ct.n_seqs)]
        self.contigTids = {n: i for i, n in enumerate(self.names)}

    def close(self):
        if self.BWTData:
            self._lib.lc_destroy(self.BWTData)
            self.BWTData = None

    def GetReferenceContigsInfo(self):      # gobwa.go:27-40
        return list(self.names), list(self.lens)

    def GetSeq(self, chrom: str, start: int, end: int, reversed_: bool) -> bytes:   # gobwa.go:50-80
        tid = self.contigTids[chrom]
        buf = np.empty(max(0, abs(end - start)), dtype=np.uint8)
        n = self._lib.lc_fetch_seq(self.BWTData, tid, start, end, buf.ctypes.data)
        codes = buf[:n]
        out = np.zeros(max(0, end - start), dtype=np.uint8)     # make([]byte, end-start): zero bytes past a clamped fetch
        if reversed_:
            out[:n] = np.frombuffer(b"TGCA", dtype=np.uint8)[codes][::-1]
        else:
            out[:n] = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
        return out.tobytes()

    # ---- batched driver ----
    def align_packed(self, bases: np.ndarray, off: np.ndarray, score_delta: int = 25, copy: bool = True) -> BatchResult:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n_pairs = (len(off) - 1) // 2
        res = C.POINTER(lc_result_t)()
        rc = self._lib.lc_align_packed(self.BWTData, n_pairs, bases.ctypes.data, off.ctypes.data, score_delta, C.byref(res))
        if rc != 0:
            raise RuntimeError(f"lc_align_packed failed ({rc}): {last_error()}")
        return BatchResult(res.contents, copy=copy, owner=(self._lib, self.BWTData, res))

    def clone(self) -> "GoBwaReference":
        """A second context on the same device index (own stream and pools, no second copy of the index): lc_clone."""
        ctx = self._lib.lc_clone(self.BWTData)
        if not ctx:
            raise RuntimeError("lc_clone failed: " + last_error())
        return GoBwaReference(ctx, self._lib)


def GoBwaLoadReference(path: str, device: int = 0, settings: GoBwaSettings | None = None):
    """gobwa.go:128-147.  Returns None (after a log line) when the index cannot be loaded, like the reference."""
    lib = load_library()
    ctx = lib.lc_init(path.encode(), device, C.byref(settings.Settings) if settings else None)
    if not ctx:
        logging.getLogger("gobwa").warning("THAT DIDN'T WORK!!! FILE PROBLEMS? (%s)", last_error())
        return None
    return GoBwaReference(ctx, lib)


def EnumerateContigs(ref: GoBwaReference, callback):    # gobwa.go:435-447
    for n, l in zip(ref.names, ref.lens):
        callback(n, l)


def SequenceConvert(seq) -> bytes:                      # gobwa.go:159-167 (kept for API parity; the library converts on the GPU)
    t = np.full(256, 4, dtype=np.uint8)
    for ch, v in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"Tt", 3)):
        for c in ch:
            t[c] = v
    if isinstance(seq, str):
        seq = seq.encode()
    return t[np.frombuffer(seq, dtype=np.uint8)].tobytes()


def InterpretAlign(ref: GoBwaReference, reg) -> EasyAlignment:      # gobwa.go:339-371
    rid = int(reg["rid"])
    off = ref.offsets[rid]
    rb, re_ = int(reg["rb"]), int(reg["re"])
    res = EasyAlignment()
    if rb < ref.l_pac:
        res.Offset, res.Reversed = rb - off, False
    else:
        res.Offset, res.Reversed = ref.l_pac * 2 - 1 - rb - off, True
    res.Alignment_end = re_ - off if re_ < ref.l_pac else ref.l_pac * 2 - 1 - re_ - off
    res.Contig = ref.names[rid]
    res.Secondary = int(reg["secondary"]) >= 0 or int(reg["secondary_all"]) > 0
    res.ChainedHit = reg
    res.Score, res.ReadS, res.ReadE = int(reg["score"]), int(reg["qb"]), int(reg["qe"])
    return res


def _as_bytes(s) -> bytes:
    return s.encode() if isinstance(s, str) else bytes(s)


def GoBwaMemMateSWBatch(ref: GoBwaReference, settings: GoBwaSettings, pairs, score_delta: int):
    """All pairs of a barcode group (or several) in one call.  pairs: sequence of (read1, read2) ASCII bytes/str.
    Returns (list of (algns_read1, algns_read2) exactly as GoBwaMemMateSW would, BatchResult)."""
    seqs = []
    for a, b in pairs:
        seqs.append(_as_bytes(a)); seqs.append(_as_bytes(b))
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    bases = np.frombuffer(b"".join(seqs), dtype=np.uint8) if off[-1] else np.zeros(1, dtype=np.uint8)
    br = ref.align_packed(bases, off, score_delta)
    out = []
    for p in range(len(seqs) // 2):
        ends = []
        for e in (0, 1):
            r = 2 * p + e
            ends.append([InterpretAlign(ref, br.regs[i]) for i in range(int(br.reg_off[r]), int(br.reg_off[r + 1]))])
        out.append((ends[0], ends[1]))
    return out, br


def GoBwaMemMateSW(ref, settings, read1, read2, arena, score_delta):   # gobwa.go:226-337
    out, br = GoBwaMemMateSWBatch(ref, settings, [(read1, read2)], score_delta)
    arena.Push(br)
    return out[0]


def GoBwaAlign(ref, settings, seq, arena):                              # gobwa.go:174-203 (mem_align1_core on one read)
    lib = load_library()
    codes = np.frombuffer(SequenceConvert(seq), dtype=np.uint8)
    v = lc_alnreg_v()
    rc = lib.lc_mem_align1_core(ref.BWTData, len(codes), codes.ctypes.data, C.byref(v))
    if rc != 0:
        raise RuntimeError(f"lc_mem_align1_core failed ({rc}): {last_error()}")
    regs = np.frombuffer((C.c_char * (88 * v.n)).from_address(v.a), dtype=ALNREG_DTYPE).copy() if v.n else np.zeros(0, dtype=ALNREG_DTYPE)
    _libc.free(v.a)
    arena.Push(regs)
    return [InterpretAlign(ref, r) for r in regs]


def GoBwaSmithWaterman(ref, settings, seq, alignment, arena) -> SingleReadAlignment:   # gobwa.go:400-415,449-488
    lib = load_library()
    codes = np.frombuffer(SequenceConvert(seq), dtype=np.uint8)
    reg = np.array([alignment], dtype=ALNREG_DTYPE)
    aln = np.zeros(1, dtype=ALN_DTYPE)
    cig = C.POINTER(C.c_uint32)()
    rc = lib.lc_mem_reg2aln(ref.BWTData, len(codes), codes.ctypes.data, reg.ctypes.data, aln.ctypes.data, C.byref(cig))
    if rc != 0:
        raise RuntimeError(f"lc_mem_reg2aln failed ({rc}): {last_error()}")
    a = aln[0]
    res = SingleReadAlignment()
    res.Pos, res.Chrom = int(a["pos"]), ref.names[int(a["rid"])] if int(a["rid"]) >= 0 else ""
    res.Reversed, res.EditDistance = bool(a["is_rev"]), int(a["NM"])
    for i in range(int(a["n_cigar"])):
        res.Cigar += [int(cig[i]) & 0xf, int(cig[i]) >> 4]
    _libc.free(cig)
    res.Score = int(alignment["score"])
    res.Sub = max(int(alignment["sub"]), int(alignment["csub"]))
    res.AltSC = int(alignment["alt_sc"])
    res.ReadS, res.ReadE = int(alignment["qb"]), int(alignment["qe"])
    return res
