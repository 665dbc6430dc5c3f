"""oracle/reader_port.py — TEST INFRASTRUCTURE ONLY: a plain-Python restatement of Lariat's 9-line FASTQ reader.

Follows go/src/fastqreader/reader.go statement by statement (ReadOneLine :91-149, ReadBarcodeSet :190-260; the file is inflated
with Python's gzip instead of the `gunzip -c` pipe of zipread.go:63-85).  Pinned by the reference's own assertions
(go/src/test/fastq_reader_test.go) on the reference's own fixture (tests/golden/reader/1.fq), see tests/test_reader.py.
The product reader (lariat_b200/csrc/lc_reader.cpp) is checked against this; nothing in the product imports it.
"""
import gzip
import io


class FastQRecord:      # reader.go:15-33
    __slots__ = ("Read1", "ReadQual1", "Read2", "ReadQual2", "TrimBases", "TrimQuals", "Barcode10X", "Barcode10XQual", "RawBarcode10X",
                 "Barcode", "BarcodeQual", "ReadInfo", "ReadGroupId")


class FastQReader:
    def __init__(self, path):      # OpenFastQ (:63-77)
        raw = open(path, "rb").read()
        self.buf = io.BytesIO(gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw)
        self.Line = 0
        self.DefferedError = None
        self.Pending = None
        self.LastBarcode = None

    def _readline(self):            # bufio ReadString('\n'): error unless the line ends with a newline
        line = self.buf.readline()
        if not line.endswith(b"\n"):
            raise EOFError
        return line

    def ReadOneLine(self, trim):    # :91-149
        result = FastQRecord()
        while True:
            self.Line += 1
            line = self._readline()
            if line[0:1] == b"@":
                fields = line[1:-1].split()          # strings.Fields
                result.ReadInfo = fields[0]
                result.ReadGroupId = b"" if len(fields) < 2 else fields[-1]
                break
        stuff = []
        for _ in range(8):
            stuff.append(self._readline()[:-1])
        to_trim = min(len(stuff[0]), trim)
        result.Read1, result.ReadQual1 = stuff[0][to_trim:], stuff[1][to_trim:]
        result.TrimBases, result.TrimQuals = stuff[0][:to_trim], stuff[1][:to_trim]
        result.Read2, result.ReadQual2 = stuff[2], stuff[3]
        barcodes = stuff[4].split(b",")
        result.Barcode10X, result.RawBarcode10X = barcodes[0], barcodes[-1]
        result.Barcode10XQual, result.Barcode, result.BarcodeQual = stuff[5], stuff[6], stuff[7]
        return result

    def ReadBarcodeSet(self, trim):  # :190-260 -> (records, err, bool)
        new_barcode = False
        if self.DefferedError is not None:
            return None, self.DefferedError, False
        record_array = []
        index = 0
        if self.Pending is not None:
            record_array.append(self.Pending)
            self.Pending = None
            index += 1
        while index < 30000:
            record_array.append(None)
            try:
                record_array[index] = self.ReadOneLine(trim)
            except EOFError as e:
                if index == 0:
                    return None, e, False
                self.DefferedError = e
                break
            not_whitelist = b"-" not in record_array[0].Barcode10X
            if record_array[0].Barcode10X != record_array[index].Barcode10X or (not_whitelist and index >= 200):
                self.Pending = record_array[index]
                new_barcode = True
                break
            elif self.LastBarcode is not None and record_array[0].Barcode10X == self.LastBarcode and index >= 200:
                new_barcode = False
                break
            index += 1
        if record_array:
            self.LastBarcode = bytes(record_array[0].Barcode10X)
        end = len(record_array)
        if new_barcode or self.DefferedError is not None:      # (the only error this restatement produces is io.EOF)
            end -= 1
        else:
            return record_array[:end], None, False
        return record_array[:end], None, True
