"""The C-ABI library: loads, exports every symbol include/cuda_lariat.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import GOLDEN, ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "cuda_lariat.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(lc_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(lib_path):
    lib = C.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cuda_lariat.h but not exported"
    from lariat_b200 import gobwa
    assert sorted(gobwa.EXPORTS) == names


def test_struct_layouts_match_reference_abi(lib_path):
    from lariat_b200 import gobwa
    assert gobwa.ALNREG_DTYPE.itemsize == 88            # sizeof(mem_alnreg_t), go/src/gobwa/bwa/bwamem.h:59-77
    assert C.sizeof(gobwa.lc_opt_t) == 168              # sizeof(mem_opt_t) with gcc on x86-64
    o = gobwa.GoBwaAllocSettings().Settings
    assert (o.a, o.b, o.o_del, o.e_del, o.w, o.zdrop, o.min_seed_len, o.max_occ, o.split_width) == (1, 4, 6, 1, 100, 100, 19, 500, 10)
    assert o.max_mem_intv == 20 and abs(o.split_factor - 1.5) < 1e-9 and list(o.mat)[:6] == [1, -4, -4, -4, -1, -4]


def test_no_cpu_fallback(lib_path):
    """Without a CUDA device lc_init must fail loudly (LC_ENODEV), never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lariat_b200 import gobwa
    ref = gobwa.GoBwaLoadReference(os.path.join(GOLDEN, "phix", "PhiX.fa"))
    assert ref is None
    assert "no CUDA device" in gobwa.last_error()


def test_product_does_not_reference_oracle():
    """Nothing under lariat_b200/ may include, link or call anything under oracle/."""
    for dp, _, fs in os.walk(os.path.join(ROOT, "lariat_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                # no include / link / dlopen / exec of the checker from product code (prose mentions in comments are fine)
                assert not re.search(r'#include\s*"[^"]*(oracle|port\.h)', txt), f
                assert "liboracle" not in txt and "replay_port" not in txt and "replay_ref" not in txt and "_ref/" not in txt, f
                assert not re.search(r"(import|from)\s+oracle", txt), f
