"""The north-star target size: a 3.1 Gb reference (BASELINE.json configs[2]).  Its BWT has 6.2e9 symbols, which is the only place the
64-bit interval arithmetic meets real data: occ superblocks >= 1, the 2-bit high fields of the packed interval entries, a u64 dense
suffix array (50 GB).  The genome and reads are generated on the box (seeded), the index is built by lc_index_build, and a prefix of the
reads goes through the C ABI and through the CPU reference (unmodified libbwa driven by oracle/replay.c); the dumps must be the same bytes.
About a minute on a B200; skipped on devices that cannot hold the index (set LARIAT_SKIP_C3FULL=1 to skip it anywhere)."""
import os
import sys

import pytest

from conftest import ROOT, have_ref
from lariat_b200 import gobwa

pytestmark = pytest.mark.gpu

N_PAIRS = 6000


def test_c3_full_size_prefix_vs_cpu_reference(lib_path, tmp_path):
    import torch
    if os.environ.get("LARIAT_SKIP_C3FULL"):
        pytest.skip("LARIAT_SKIP_C3FULL set")
    if not torch.cuda.is_available() or torch.cuda.get_device_properties(0).total_memory < 120e9:
        pytest.skip("needs ~70 GB of device memory for the 3.1 Gb index with its dense suffix array")
    sys.path.insert(0, ROOT)
    import bench
    bench.PAIRS_OVERRIDE = 20000
    try:
        fa, pairs, g, r = bench.prepare_data("c3", 1.0)
    finally:
        bench.PAIRS_OVERRIDE = 0
    assert g.total_bp >= 3_000_000_000
    n, bases, off = bench.load_pairs(pairs)
    assert n >= N_PAIRS
    ref = gobwa.GoBwaLoadReference(fa)
    assert ref is not None, gobwa.last_error()
    try:
        br = ref.align_packed(bases[:off[2 * N_PAIRS]], off[:2 * N_PAIRS + 1], 25)
        got = br.dump()
        n_cand = len(br.regs)
    finally:
        ref.close()
    out = tmp_path / "c3_ref.bin"
    kind, j = bench.cpu_reference(fa, pairs, N_PAIRS, os.cpu_count() or 8, out=str(out))
    want = out.read_bytes()
    assert n_cand > 5 * N_PAIRS          # the repeat-rich regime: many candidates per pair
    assert got == want, f"CUDA path differs from the CPU {kind} on the 3.1 Gb reference"
