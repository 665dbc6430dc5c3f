"""The drop-in boundary under the caller's real usage pattern (ownership, threading, small calls), and the exports
round 1 left untested.  Everything goes through the C ABI; the oracle (oracle/liboracle_port.so, replay output) is the checker."""
import ctypes as C
import json
import os
import subprocess
import threading

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from lariat_b200 import gobwa
from test_gpu_parity import _load_pairs

pytestmark = pytest.mark.gpu
TINY = os.path.join(GOLDEN, "tiny", "tiny.fa")
TINY_PAIRS = os.path.join(GOLDEN, "tiny", "tiny.pairs")
WANT = os.path.join(GOLDEN, "tiny", "tiny.ref.out")
DRIVER = os.path.join(ROOT, "lariat_b200", "lc_percall_driver")


@pytest.fixture(scope="module")
def tiny_ref(lib_path):
    ref = gobwa.GoBwaLoadReference(TINY)
    assert ref is not None, gobwa.last_error()
    yield ref
    ref.close()


@pytest.mark.parametrize("args,env", [(["-t", "16", "-c", "7"], {}), (["-t", "5", "-c", "100"], {}), (["-t", "24", "-c", "13", "-k", "3"], {}),
                                      (["-t", "8", "-c", "9"], {"LC_COALESCE_MAX": "0"}), (["-t", "12", "-c", "11"], {"LC_COALESCE_LANES": "1", "LC_COALESCE_BATCH": "40"}),
                                      (["-t", "12", "-c", "50"], {"LC_COALESCE_LINGER_US": "200"})])
def test_c_driver_one_barcode_per_call_from_many_threads(lib_path, tmp_path, args, env):
    """The C driver (lariat_b200/csrc/lc_percall_driver.c) is the Go worker pool minus Go: many threads, one small barcode group per
    call, EVERY result kept alive until the end.  Its dump in barcode order must be the golden bytes, and no held result may change
    while later calls run (the driver re-checksums each before dumping) — with merging on, off, squeezed, and over cloned contexts."""
    assert os.path.exists(DRIVER), "build the library first (make -C lariat_b200/csrc)"
    out = str(tmp_path / "drv.out")
    e = dict(os.environ, **env)
    r = subprocess.run([DRIVER, TINY, TINY_PAIRS, out] + args, capture_output=True, text=True, env=e)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(WANT, "rb").read()
    j = json.loads(r.stderr.strip().splitlines()[-1])
    assert j["pairs"] == 605 and j["pairs_per_s"] > 0


def test_c_driver_timed_mode_holds_previous_results(lib_path):
    """Timed mode: each worker keeps its last 3 results while it makes further calls and checks them before freeing."""
    r = subprocess.run([DRIVER, TINY, TINY_PAIRS, "-", "-t", "16", "-c", "10", "-r", "3", "-h", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stderr.strip().splitlines()[-1])["calls"] == 61 * 3


def test_result_is_callers_until_freed(tiny_ref):
    """A result held with copy=False must survive any number of later calls on the same context (small and large), from this and
    other threads, and on a clone; Arena.Free releases it (G/gobwa.go:107-126, I/lariat.go:546)."""
    bases, off = _load_pairs(TINY_PAIRS)
    want = open(WANT, "rb").read()
    arena = gobwa.NewArena()
    held = tiny_ref.align_packed(bases, off, 25, copy=False)
    arena.Push(held)
    snap = held.dump()
    assert snap == want
    clone = tiny_ref.clone()
    errs = []

    def hammer(ref, lo, hi):
        try:
            for _ in range(5):
                sub = ref.align_packed(bases[off[2 * lo]:off[2 * hi]], off[2 * lo:2 * hi + 1] - off[2 * lo], 25)
                assert sub.n_reads == 2 * (hi - lo)
        except Exception as ex:      # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=hammer, args=(tiny_ref if k % 2 else clone, 40 * k, 40 * k + 90)) for k in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    assert held.dump() == want          # still the caller's
    assert clone.align_packed(bases, off, 25).dump() == want
    clone.close()
    arena.Free()
    assert tiny_ref.align_packed(bases, off, 25).dump() == want      # the library recycles the freed block


def test_coalesced_calls_equal_separate_calls(lib_path, monkeypatch):
    """Merging small calls across threads must not change any call's result: 60 threads x one 10-pair group each, against the
    same groups aligned one by one with merging disabled."""
    bases, off = _load_pairs(TINY_PAIRS, 600)
    ref = gobwa.GoBwaLoadReference(TINY)
    lib = gobwa.load_library()
    got = [None] * 60

    def one(k):
        lo, hi = 10 * k, 10 * k + 10
        got[k] = ref.align_packed(bases[off[2 * lo]:off[2 * hi]], off[2 * lo:2 * hi + 1] - off[2 * lo], 25).dump()

    th = [threading.Thread(target=one, args=(k,)) for k in range(60)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert lib.lc_set_coalescing(ref.BWTData, 0, 65536, 1, 0) == 0
    for k in range(60):
        lo, hi = 10 * k, 10 * k + 10
        assert ref.align_packed(bases[off[2 * lo]:off[2 * hi]], off[2 * lo:2 * hi + 1] - off[2 * lo], 25).dump() == got[k], k
    ref.close()


def test_align_pairs_pointer_form(tiny_ref):
    """lc_align_pairs (per-read pointers, the shape of the Go call G/gobwa.go:226) == lc_align_packed."""
    lib = gobwa.load_library()
    bases, off = _load_pairs(TINY_PAIRS, 150)
    n = 150
    r1 = [bytes(bases[off[2 * p]:off[2 * p + 1]]) for p in range(n)]
    r2 = [bytes(bases[off[2 * p + 1]:off[2 * p + 2]]) for p in range(n)]
    r1[3] = b""                      # a zero-length end (G/gobwa.go:243)
    a1 = (C.c_char_p * n)(*r1); a2 = (C.c_char_p * n)(*r2)
    l1 = (C.c_int32 * n)(*[len(x) for x in r1]); l2 = (C.c_int32 * n)(*[len(x) for x in r2])
    res = C.POINTER(gobwa.lc_result_t)()
    assert lib.lc_align_pairs(tiny_ref.BWTData, n, a1, l1, a2, l2, 25, C.byref(res)) == 0, gobwa.last_error()
    got = gobwa.BatchResult(res.contents, copy=True, owner=(lib, tiny_ref.BWTData, res))
    seqs = [x for p in range(n) for x in (r1[p], r2[p])]
    o2 = np.zeros(2 * n + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=o2[1:])
    want = tiny_ref.align_packed(np.frombuffer(b"".join(seqs), dtype=np.uint8), o2, 25)
    assert got.dump() == want.dump()
    assert got.reg_off[7] == got.reg_off[6]          # read 6 = the empty end of pair 3


def test_bad_arguments_are_rejected_not_executed(tiny_ref):
    lib = gobwa.load_library()
    res = C.POINTER(gobwa.lc_result_t)()
    long_read = np.frombuffer(b"ACGT" * 176, dtype=np.uint8)[:701].copy()          # LC_MAX_READ_LEN + 1 (B/bwamem.c:598-602 guard)
    for n_in_call in (1, 5000):      # the merged-small-call path and the direct path
        bases = np.concatenate([long_read, long_read] * n_in_call)
        off = np.arange(0, 701 * (2 * n_in_call + 1), 701, dtype=np.int64)
        assert lib.lc_align_packed(tiny_ref.BWTData, n_in_call, bases.ctypes.data, off.ctypes.data, 25, C.byref(res)) == -3      # LC_EARG
        assert "LC_MAX_READ_LEN" in gobwa.last_error()
    off = np.array([0, 150, 100], dtype=np.int64)      # decreasing offsets: a negative length
    bases = np.zeros(300, dtype=np.uint8)
    assert lib.lc_align_packed(tiny_ref.BWTData, 1, bases.ctypes.data, off.ctypes.data, 25, C.byref(res)) == -3
    o = gobwa.GoBwaAllocSettings()
    o.Settings.a = 60                                   # 700 * 60 > int16: the 16-bit DP cells would wrap
    for i in range(5):
        for j in range(5):
            o.Settings.mat[5 * i + j] = -1 if (i == 4 or j == 4) else (60 if i == j else -4)
    assert gobwa.GoBwaLoadReference(TINY, settings=o) is None and "range" in gobwa.last_error()


def test_collect_intv_batch_vs_oracle(tiny_ref, oracle_port):
    """lc_collect_intv_batch (mem_collect_intv, B/bwamem.c:114-162) against the oracle's port_collect_intv on every read of the golden set:
    same intervals (x0, size, read span) in the same order.  (x1 is not kept past seeding and is reported as 0.)"""
    lib = gobwa.load_library()
    port = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    port.port_idx_load.restype = C.c_void_p
    port.port_opt_init.restype = C.c_void_p

    class IntvV(C.Structure):
        _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(C.c_uint64))]
    port.port_collect_intv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.POINTER(IntvV)]
    ix = port.port_idx_load(TINY.encode())
    opt = port.port_opt_init()
    bases, off = _load_pairs(TINY_PAIRS, 200)
    codes = np.frombuffer(gobwa.SequenceConvert(bytes(bases)), dtype=np.uint8).copy()
    n = 400
    out_off = np.zeros(n + 1, dtype=np.int64)
    cap = 400 * 64
    out = np.zeros((cap, 4), dtype=np.uint64)
    ms = C.c_float()
    tot = lib.lc_collect_intv_batch(tiny_ref.BWTData, n, codes.ctypes.data, off.ctypes.data, out_off.ctypes.data, out.ctypes.data, cap, C.byref(ms))
    assert tot > 0, gobwa.last_error()
    libc = C.CDLL(None)
    for r in range(n):
        v = IntvV(0, 0, None)
        seq = codes[off[r]:off[r + 1]].tobytes()
        port.port_collect_intv(opt, ix, len(seq), seq, C.byref(v))
        want = [(v.a[4 * i], v.a[4 * i + 2], v.a[4 * i + 3]) for i in range(v.n)]
        got = [(int(x[0]), int(x[2]), int(x[3])) for x in out[out_off[r]:out_off[r + 1]]]
        assert got == want, r
        libc.free(v.a)


def test_mem_matesw_export_vs_oracle(tiny_ref, oracle_port):
    """lc_mem_matesw (mem_matesw, B/bwamem_pair.c:111-180) against the oracle's port_matesw: anchors = the hits of one end, mate =
    the other end; the grown `ma` list and the number of SWs performed must agree (the golden set is rescue-heavy)."""
    lib = gobwa.load_library()
    port = C.CDLL(os.path.join(oracle_port, "liboracle_port.so"))
    port.port_idx_load.restype = C.c_void_p
    port.port_opt_init.restype = C.c_void_p

    class RegV(C.Structure):
        _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]
    port.port_matesw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.POINTER(RegV)]
    port.port_align1_core.restype = RegV
    port.port_align1_core.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p]
    ix = port.port_idx_load(TINY.encode())
    opt = port.port_opt_init()
    pes = (gobwa.lc_pestat_t * 4)()
    for i in range(4):
        pes[i].failed = 1
    pes[1].failed, pes[1].low, pes[1].high = 0, -35, 500          # G/gobwa.go:229-237
    bases, off = _load_pairs(TINY_PAIRS, 120)
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    n_sw = n_grown = 0
    for p in range(120):
        s1 = gobwa.SequenceConvert(bytes(bases[off[2 * p]:off[2 * p + 1]]))
        s2 = gobwa.SequenceConvert(bytes(bases[off[2 * p + 1]:off[2 * p + 2]]))
        a1 = port.port_align1_core(opt, ix, len(s1), C.create_string_buffer(s1, len(s1)).raw)
        a2 = port.port_align1_core(opt, ix, len(s2), C.create_string_buffer(s2, len(s2)).raw)
        if a1.n == 0:
            continue
        anchors = np.frombuffer((C.c_char * (88 * a1.n)).from_address(a1.a), dtype=gobwa.ALNREG_DTYPE).copy()
        ma0 = np.frombuffer((C.c_char * (88 * a2.n)).from_address(a2.a), dtype=gobwa.ALNREG_DTYPE).copy() if a2.n else np.zeros(0, gobwa.ALNREG_DTYPE)
        for k in range(min(3, len(anchors))):
            # the oracle
            wa = libc.malloc(88 * (len(ma0) + 1)); C.memmove(wa, ma0.ctypes.data, 88 * len(ma0))
            wv = RegV(len(ma0), len(ma0) + 1, wa)
            want_n = port.port_matesw(opt, ix, C.addressof(pes), anchors[k:k + 1].ctypes.data, len(s2), s2, C.byref(wv))
            want = bytes((C.c_char * (88 * wv.n)).from_address(wv.a))
            # the product
            ga = libc.malloc(88 * (len(ma0) + 1)); C.memmove(ga, ma0.ctypes.data, 88 * len(ma0))
            gv = gobwa.lc_alnreg_v(len(ma0), len(ma0) + 1, ga)
            codes = np.frombuffer(s2, dtype=np.uint8).copy()
            got_n = lib.lc_mem_matesw(tiny_ref.BWTData, pes, anchors[k:k + 1].ctypes.data, len(s2), codes.ctypes.data, C.byref(gv))
            assert got_n == want_n, (p, k, gobwa.last_error())
            assert bytes((C.c_char * (88 * gv.n)).from_address(gv.a)) == want, (p, k)
            n_sw += want_n; n_grown += int(wv.n > len(ma0))
            libc.free(wv.a); libc.free(gv.a)
    assert n_sw > 20 and n_grown > 5          # the test really exercised Smith-Watermans and insertions


def test_getalignments_post_processing_on_device(lib_path, oracle_port, tmp_path):
    """lc_set_alnx: per candidate matches / mismatches / indels / soft clips / pos, aend and the two mismatch position lists, exactly
    what inference.GetAlignments computes with one GetSeq call per candidate (I/lariat.go:1573-1650) — against the golden records
    (written from the compiled reference), through the direct path, the merged-small-call path, chunks over lanes, and on fresh
    repeat-rich data against the oracle.  n_mismatch of lc_aln_t is checked on the way."""
    lib = gobwa.load_library()
    want = open(os.path.join(GOLDEN, "tiny", "tiny.ref.x"), "rb").read()
    bases, off = _load_pairs(TINY_PAIRS)
    ref = gobwa.GoBwaLoadReference(TINY)
    lib.lc_set_alnx(ref.BWTData, 1)
    br = ref.align_packed(bases, off, 25)                      # 605 pairs: the merged-small-call path
    assert br.dump() == open(WANT, "rb").read() and br.dump_alnx() == want
    # lc_aln_t.n_mismatch == NM - gap bases (B/bwa.c:169-199)
    for i in range(len(br.aln)):
        gaps = sum(int(c) >> 4 for c in br.cigar_of(i) if (int(c) & 0xf) in (1, 2))
        assert int(br.aln[i]["n_mismatch"]) == int(br.aln[i]["NM"]) - gaps
    assert lib.lc_set_coalescing(ref.BWTData, 0, 65536, 1, 0) == 0
    assert ref.align_packed(bases, off, 25).dump_alnx() == want                         # the direct path
    assert lib.lc_set_chunk_pairs(ref.BWTData, 97) == 0
    assert ref.align_packed(bases, off, 25).dump_alnx() == want                         # chunks over two lanes
    lib.lc_set_alnx(ref.BWTData, 0)
    assert ref.align_packed(bases, off, 25).alnx is None
    ref.close()
    # fresh data
    from lariat_b200 import synth
    from test_gpu_parity import _fresh_dataset, _oracle_replay
    g = synth.GenomeSpec(1_500_000, 3, 2025, segdup_frac=0.10, segdup_len=(1000, 8000), segdup_div=0.005,
                         families=[(300, 700, 0.03, 0.10), (2500, 40, 0.01, 0.03)], n_tandem=50, name="xchr")
    r = synth.ReadSpec(2500, 25, degraded_frac=0.3, repeat_bias=0.8, indel_rate=0.004, seed=10)
    fa, pairs = _fresh_dataset(tmp_path, "alnx", g, r)
    w_out, w_x = str(tmp_path / "want.out"), str(tmp_path / "want.x")
    subprocess.check_call([_oracle_replay(oracle_port), fa, pairs, w_out, "-t", "8", "-x", w_x], stderr=subprocess.DEVNULL)
    ref = gobwa.GoBwaLoadReference(fa)
    lib.lc_set_alnx(ref.BWTData, 1)
    b2, o2 = _load_pairs(pairs)
    br = ref.align_packed(b2, o2, 25)
    assert br.dump() == open(w_out, "rb").read()
    assert br.dump_alnx() == open(w_x, "rb").read()
    ref.close()
