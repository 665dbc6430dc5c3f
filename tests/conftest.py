import os
import subprocess
import sys

import pytest

# The golden sets are small; without this every test batch would take the low-latency route (a warp per read for seeding and
# extension, batches of <= 16384 reads) and the state-machine kernels would go untested.  Tests of that route set the variable themselves.
os.environ.setdefault("LC_SMALL_READS", "0")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "_build")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_port():
    """Builds the plain-C oracle (oracle/liboracle_port.so + replay_port) with gcc; test infrastructure only."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return os.path.join(ROOT, "oracle")


@pytest.fixture(scope="session")
def hostsim_bin():
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "hostsim")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "lariat_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "lariat_b200", "csrc"))
                    if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-o", out, src])
    return out


@pytest.fixture(scope="session")
def hostsim_sb_bin():
    """The same harness with 2^12-block superblocks and an 18-bit low word in the packed 64-bit interval entries: on the tiny golden
    index (8e5 symbols) the superblock index runs 0..3 and the 2-bit high fields are non-zero — the arithmetic a >= 2^32-symbol
    index exercises in the product build (LC_SB_SHIFT 26, LC_PK_LOW_BITS 32)."""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "hostsim_sb")
    src = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
    deps = [src] + [os.path.join(ROOT, "lariat_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "lariat_b200", "csrc"))
                    if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-DLC_SB_SHIFT=12", "-DLC_PK_LOW_BITS=18", "-o", out, src])
    return out


@pytest.fixture(scope="session")
def lib_path():
    p = os.path.join(ROOT, "lariat_b200", "libcuda_lariat.so")
    if not os.path.exists(p):
        import __graft_entry__ as g
        g.build()
    return p


def have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bwa"))
