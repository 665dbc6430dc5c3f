"""The index builder (lariat_b200/csrc/lc_index_build.{hpp,cu}) against `bwa index`.

CPU: tests/hostsim/ibsim.cpp runs the builder's element-wise bodies and orchestration with host loops (test scaffolding) and
its five output files are compared byte for byte with (a) the committed golden indexes, which the reference's own `bwa index`
wrote (tests/golden/phix is the reference's fixture, go/src/test/inputs/phix), and (b) `oracle/_ref/bwa index` run on fresh
adversarial genomes when the compiled reference is present.  GPU (-m gpu): the same comparisons through the C ABI (lc_index_build).
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import BUILD, GOLDEN, ROOT, have_ref

EXTS = ("pac", "ann", "amb", "bwt", "sa")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@pytest.fixture(scope="session")
def ibsim_bin():
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "ibsim")
    src = os.path.join(ROOT, "tests", "hostsim", "ibsim.cpp")
    deps = [src, os.path.join(ROOT, "lariat_b200", "csrc", "lc_index_build.hpp"), os.path.join(ROOT, "lariat_b200", "csrc", "lc_fmindex.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", out, src])
    return out


def contigs_of_index(prefix):
    """(name, codes) per contig, decoded from <prefix>.pac/.ann (the golden directories hold the index, not the FASTA)."""
    ann = open(prefix + ".ann").read().split("\n")
    L, n = int(ann[0].split()[0]), int(ann[0].split()[1])
    pac = np.fromfile(prefix + ".pac", dtype=np.uint8)
    codes = np.stack([(pac >> 6) & 3, (pac >> 4) & 3, (pac >> 2) & 3, pac & 3], 1).reshape(-1)[:L]
    out = []
    for i in range(n):
        off, ln = int(ann[2 + 2 * i].split()[0]), int(ann[2 + 2 * i].split()[1])
        out.append((ann[1 + 2 * i].split()[1], codes[off:off + ln].copy()))
    return out


def write_fa(path, contigs):
    with open(path, "wb") as f:
        for name, c in contigs:
            f.write(b">" + name.encode() + b"\n" + ACGT[c].tobytes() + b"\n")


def adversarial_genomes():
    """End-of-text and long-tie cases: poly-A ends (zero padding vs the implicit sentinel), exact tandem arrays and exact
    duplications (hundreds of refinement rounds), lengths around the 4 / 32 / 128-symbol packing boundaries."""
    rng = np.random.default_rng(1)
    r = lambda n: rng.integers(0, 4, n, dtype=np.uint8)
    z = lambda n: np.zeros(n, np.uint8)
    unit = r(57)
    big = r(60000)
    big[20000:30000] = big[0:10000]
    big[40000:46000] = 3 - big[2000:8000][::-1]
    return {
        "polyA_end": [np.concatenate([r(777), z(100)]), z(64), np.concatenate([z(40), r(301), z(33)])],
        "polyT_start": [np.concatenate([np.full(90, 3, np.uint8), r(500)])],
        "tandem_exact": [np.concatenate([r(200)] + [unit] * 40 + [r(131)])],
        "len1": [np.array([2], np.uint8)], "len4": [z(4)], "len32": [r(32)], "len64x2": [r(64), r(64)], "len127": [r(127)],
        "allA": [z(1000)], "AT": [np.tile(np.array([0, 3], np.uint8), 700)],
        "dup_exact": [big],
    }


def same_files(a, b):
    return [e for e in EXTS if open(f"{a}.{e}", "rb").read() != open(f"{b}.{e}", "rb").read()]


@pytest.mark.parametrize("golden", ["tiny/tiny.fa", "phix/PhiX.fa"])
@pytest.mark.parametrize("chunk", ["0", "5000", "300"])
def test_builder_logic_reproduces_bwa_index_golden(ibsim_bin, tmp_path, golden, chunk):
    ref = os.path.join(GOLDEN, golden)
    fa = str(tmp_path / "g.fa")
    write_fa(fa, contigs_of_index(ref))
    subprocess.check_call([ibsim_bin, fa, fa, chunk], stderr=subprocess.DEVNULL)
    assert same_files(fa, ref) == []


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref/bwa) not present")
def test_builder_logic_vs_reference_on_adversarial_genomes(ibsim_bin, tmp_path):
    bwa = os.path.join(ROOT, "oracle", "_ref", "bwa")
    for name, seqs in adversarial_genomes().items():
        contigs = [(f"c{i}", s) for i, s in enumerate(seqs)]
        want, got = str(tmp_path / f"{name}.ref.fa"), str(tmp_path / f"{name}.fa")
        write_fa(want, contigs)
        write_fa(got, contigs)
        subprocess.check_call([bwa, "index", want], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
        for chunk in ("0", "97"):
            subprocess.check_call([ibsim_bin, got, got, chunk], stderr=subprocess.DEVNULL)
            assert same_files(got, want) == [], (name, chunk)


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref/bwa) not present")
def test_torch_test_helper_vs_reference(tmp_path):
    """index_build.build_index_torch (the small-genome helper CPU tests use for fresh data) against `bwa index`."""
    from lariat_b200 import index_build, synth
    g = synth.GenomeSpec(300_000, 3, 31, segdup_frac=0.05, segdup_len=(500, 4000), families=[(300, 60, 0.03, 0.10)], n_tandem=10, name="q")
    contigs, _ = synth.make_genome(g)
    want, got = str(tmp_path / "w.fa"), str(tmp_path / "g.fa")
    synth.write_fasta(want, contigs)
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "bwa"), "index", want], stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
    index_build.build_index_torch(got, contigs)
    assert same_files(got, want) == []


# ---- the product path: lc_index_build through the C ABI ---------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("chunk", ["", "5000", "300"])
def test_cuda_builder_reproduces_bwa_index_golden(lib_path, tmp_path, monkeypatch, chunk):
    from lariat_b200 import index_build
    if chunk:
        monkeypatch.setenv("LC_IB_CHUNK", chunk)
    for golden in ("tiny/tiny.fa", "phix/PhiX.fa"):
        ref = os.path.join(GOLDEN, golden)
        out = str(tmp_path / (os.path.basename(golden) + chunk))
        index_build.build_index(out, contigs_of_index(ref), 0)
        assert same_files(out, ref) == [], (golden, chunk)


@pytest.mark.gpu
def test_cuda_builder_adversarial_vs_host_logic(lib_path, ibsim_bin, tmp_path, monkeypatch):
    """The CUDA kernels + cub sorts against the host execution of the same bodies (itself pinned to `bwa index` by the CPU tests):
    adversarial genomes, plus a 6 Mb repeat-rich reference cut into ~30 chunks."""
    from lariat_b200 import index_build, synth
    cases = {k: [(f"c{i}", s) for i, s in enumerate(v)] for k, v in adversarial_genomes().items()}
    g = synth.GenomeSpec(3_000_000, 5, 77, segdup_frac=0.03, segdup_len=(1000, 6000), families=[(300, 1500, 0.10, 0.15), (3000, 60, 0.02, 0.08)],
                         n_tandem=40, name="m")
    cases["repeat6M"] = synth.make_genome(g)[0]
    for name, contigs in cases.items():
        want, got = str(tmp_path / f"{name}.w.fa"), str(tmp_path / f"{name}.g.fa")
        write_fa(want, contigs)
        subprocess.check_call([ibsim_bin, want, want, "200000"], stderr=subprocess.DEVNULL)
        monkeypatch.setenv("LC_IB_CHUNK", "200000" if name == "repeat6M" else "97")
        index_build.build_index(got, contigs, 0)
        assert same_files(got, want) == [], name
