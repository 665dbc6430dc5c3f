"""Host-side mirror of go/src/fastqreader (reader.go) over the library's block reader (lc_reader_*, lariat_b200/csrc/lc_reader.cpp).

    OpenFastQ(path) -> FastQReader            reader.go:63-77
    FastQReader.ReadBarcodeSet(trim)          reader.go:190-260 -> (BarcodeSet | None, err, full_barcode)

A BarcodeSet keeps the reads the way the GPU path takes them — `bases` / `off` go straight into GoBwaReference.align_packed — and
decodes the per-record text fields on demand (record(i) -> dict with the reference's FastQRecord field names).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import gobwa

FIELDS = ("ReadInfo", "ReadGroupId", "Barcode10X", "RawBarcode10X", "Barcode10XQual", "Barcode", "BarcodeQual", "TrimBases", "TrimQuals")


class lc_readset_t(C.Structure):
    _fields_ = [("n_records", C.c_int32), ("complete", C.c_int32), ("bases", C.POINTER(C.c_uint8)), ("quals", C.POINTER(C.c_uint8)),
                ("off", C.POINTER(C.c_int64)), ("text", C.POINTER(C.c_char)), ("text_off", C.POINTER(C.c_int64))]


class BarcodeSet:
    def __init__(self, s: lc_readset_t):
        n = int(s.n_records)
        self.n = n
        self.off = np.ctypeslib.as_array(s.off, shape=(2 * n + 1,)).copy()
        nb = int(self.off[-1])
        self.bases = np.ctypeslib.as_array(s.bases, shape=(max(nb, 1),))[:nb].copy()
        self.quals = np.ctypeslib.as_array(s.quals, shape=(max(nb, 1),))[:nb].copy()
        self._toff = np.ctypeslib.as_array(s.text_off, shape=(n * len(FIELDS) + 1,)).copy()
        self._text = C.string_at(s.text, int(self._toff[-1]))

    def __len__(self):
        return self.n

    def record(self, i: int) -> dict:
        k0 = i * len(FIELDS)
        d = {f: self._text[self._toff[k0 + k]:self._toff[k0 + k + 1]] for k, f in enumerate(FIELDS)}
        o = self.off
        d.update(Read1=self.bases[o[2 * i]:o[2 * i + 1]].tobytes(), ReadQual1=self.quals[o[2 * i]:o[2 * i + 1]].tobytes(),
                 Read2=self.bases[o[2 * i + 1]:o[2 * i + 2]].tobytes(), ReadQual2=self.quals[o[2 * i + 1]:o[2 * i + 2]].tobytes())
        return d


class FastQReader:
    def __init__(self, path: str):
        self._lib = gobwa.load_library()
        self._lib.lc_reader_open.restype = C.c_void_p
        self._lib.lc_reader_open.argtypes = [C.c_char_p]
        self._lib.lc_reader_next_set.argtypes = [C.c_void_p, C.c_int, C.POINTER(lc_readset_t)]
        self._lib.lc_reader_close.argtypes = [C.c_void_p]
        self._rd = self._lib.lc_reader_open(path.encode())
        if not self._rd:
            raise OSError(gobwa.last_error())

    def ReadBarcodeSet(self, trim: int = 0):
        s = lc_readset_t()
        rc = self._lib.lc_reader_next_set(self._rd, trim, C.byref(s))
        if rc < 0:
            raise RuntimeError(gobwa.last_error())
        if rc == 0:
            return None, EOFError(), False
        return BarcodeSet(s), None, bool(s.complete)

    def close(self):
        if self._rd:
            self._lib.lc_reader_close(self._rd)
            self._rd = None


def OpenFastQ(path: str) -> FastQReader:
    return FastQReader(path)
