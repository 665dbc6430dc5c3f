"""lariat_b200 — B200-native BWA-MEM candidate generation for Lariat's per-barcode loop.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of libcuda_lariat.so),
``gobwa.py`` (host-side mirror of go/src/gobwa/gobwa.go over that ABI), ``synth.py`` (synthetic data),
``index_build.py`` (BWA-compatible index construction).
"""
__version__ = "0.1.0"
