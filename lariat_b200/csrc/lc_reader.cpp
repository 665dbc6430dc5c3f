// lc_reader.cpp — the 9-line barcode-sorted FASTQ reader of Lariat, producing read sets already packed for lc_align_packed.
//
// Replaces go/src/fastqreader/reader.go (ReadOneLine :91-149, ReadBarcodeSet :190-260) and zipread.go (:63-85, an external `gunzip -c`
// behind a pipe).  The reference reads line by line through bufio with one allocation per line and per record and hands the worker
// goroutines slices of FastQRecord; every read is then copied and converted once more on its way into BWA (G/gobwa.go:159-167,238-239).
// At GPU rates that loop, not the aligner, is the bottleneck.  Here the file is inflated in 4 MiB blocks (zlib's
// gzread: gzip or plain text alike), records are cut out of the block without per-line allocations, and a barcode set comes out as
//   bases / quals / off   : the reads packed back to back, read 2i = Read1 of record i AFTER trimming, read 2i + 1 = Read2 —
//                           exactly the arrays lc_align_packed takes, so a set goes to the GPU without another copy or conversion;
//   text / text_off       : the other fields of every record (name, read group, barcodes, qualities, trimmed prefix).
// Grouping follows ReadBarcodeSet statement by statement: a set ends when Barcode10X changes, after 30000 records, after 200 records
// of a barcode that is not on the whitelist (no '-' in it), or — "abnormal break" — after 200 records of a set that continues the
// previous set's barcode; the record that ended a set is kept for the next one.
// Host-side C++ (no CUDA): part of libcuda_lariat.so because it is the feed of the device path; links zlib.
#include <zlib.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/cuda_lariat.h"

void lc_api_set_error(const std::string &msg);

namespace {

struct Record {                      // one FastQRecord (reader.go:15-33), fields as byte ranges of `buf`
	std::string buf;
	uint32_t f[10][2];               // [field][begin, end): 0 r1 (untrimmed), 1 q1, 2 r2, 3 q2, 4 barcode line, 5 barcode qual, 6 index, 7 index qual, 8 info, 9 rgid
	uint32_t bc10x[2], rawbc[2];
};

struct SetBuf {
	std::vector<uint8_t> bases, quals;
	std::vector<int64_t> off;
	std::string text;
	std::vector<int64_t> text_off;
	void clear() { bases.clear(); quals.clear(); off.assign(1, 0); text.clear(); text_off.assign(1, 0); }
};

} // namespace

struct lc_reader {
	gzFile fp = nullptr;
	std::vector<char> blk;           // inflated bytes not consumed yet: [pos, end)
	size_t pos = 0, end = 0;
	bool eof = false;
	long line = 0;
	bool deferred_eof = false;       // DefferedError (reader.go:44,215-223)
	bool have_pending = false;
	Record pending;                  // Pending (reader.go:45)
	bool have_last = false;
	std::string last_barcode;        // LastBarcode (reader.go:46)
	SetBuf out;
	std::vector<Record> recs;

	bool fill()
	{
		if (eof) return false;
		if (pos > 0) { memmove(blk.data(), blk.data() + pos, end - pos); end -= pos; pos = 0; }
		if (blk.size() - end < (1u << 20)) blk.resize(blk.size() + (4u << 20));
		const int n = gzread(fp, blk.data() + end, (unsigned)(blk.size() - end));
		if (n <= 0) { eof = true; return false; }
		end += (size_t)n;
		return true;
	}
	// bufio ReadString('\n'): the next line INCLUDING its newline; false at end of data (an unterminated last line is an error, as in Go)
	bool read_line(const char *&p, size_t &n)
	{
		for (;;) {
			const char *nl = (const char *)memchr(blk.data() + pos, '\n', end - pos);
			if (nl) { p = blk.data() + pos; n = (size_t)(nl - p) + 1; pos += n; ++line; return true; }
			if (!fill()) return false;
		}
	}
	// ReadOneLine (reader.go:91-149) without the trimming, which is applied when the set is packed
	bool read_record(Record &r)
	{
		const char *p; size_t n;
		r.buf.clear();
		for (;;) {                                          // search for the next start-of-record (:94-113)
			if (!read_line(p, n)) return false;
			if (p[0] == '@') break;
		}
		{   // fields := strings.Fields(line[1:len-1]); ReadInfo = fields[0]; ReadGroupId = the last field if there are two or more
			const char *s = p + 1, *e = p + n - 1;
			auto is_sp = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; };
			const char *a = s;
			while (a < e && is_sp(*a)) ++a;
			const char *b = a;
			while (b < e && !is_sp(*b)) ++b;
			const char *le = e;
			while (le > b && is_sp(le[-1])) --le;
			const char *lb = le;
			while (lb > b && !is_sp(lb[-1])) --lb;
			r.f[8][0] = 0; r.buf.append(a, b); r.f[8][1] = (uint32_t)r.buf.size();
			r.f[9][0] = (uint32_t)r.buf.size();
			if (lb > b) r.buf.append(lb, le);                // a second (or later) field exists
			r.f[9][1] = (uint32_t)r.buf.size();
		}
		for (int i = 0; i < 8; ++i) {                       // the 8 sequences of the record, newline stripped (:116-125)
			if (!read_line(p, n)) return false;
			r.f[i][0] = (uint32_t)r.buf.size(); r.buf.append(p, n - 1); r.f[i][1] = (uint32_t)r.buf.size();
		}
		// barcodes := strings.Split(line, ","): Barcode10X = the first piece, RawBarcode10X = the last (:141-143)
		const char *bb = r.buf.data() + r.f[4][0], *be = r.buf.data() + r.f[4][1];
		const char *c1 = (const char *)memchr(bb, ',', (size_t)(be - bb));
		r.bc10x[0] = r.f[4][0]; r.bc10x[1] = c1 ? (uint32_t)(c1 - r.buf.data()) : r.f[4][1];
		const char *cl = be;
		while (cl > bb && cl[-1] != ',') --cl;
		r.rawbc[0] = (uint32_t)(cl - r.buf.data()); r.rawbc[1] = r.f[4][1];
		return true;
	}
};

static bool same_barcode(const Record &a, const Record &b)
{
	const uint32_t la = a.bc10x[1] - a.bc10x[0], lb = b.bc10x[1] - b.bc10x[0];
	return la == lb && memcmp(a.buf.data() + a.bc10x[0], b.buf.data() + b.bc10x[0], la) == 0;
}
static bool not_whitelist(const Record &a) { return memchr(a.buf.data() + a.bc10x[0], '-', a.bc10x[1] - a.bc10x[0]) == nullptr; }   // reader.go:165-172

extern "C" {

lc_reader *lc_reader_open(const char *path)
{
	if (!path) { lc_api_set_error("null path"); return nullptr; }
	gzFile fp = gzopen(path, "rb");
	if (!fp) { lc_api_set_error(std::string("cannot open ") + path); return nullptr; }
	gzbuffer(fp, 1u << 20);
	lc_reader *r = new lc_reader;
	r->fp = fp;
	r->blk.resize(8u << 20);
	return r;
}

void lc_reader_close(lc_reader *r)
{
	if (!r) return;
	if (r->fp) gzclose(r->fp);
	delete r;
}

// ReadBarcodeSet (reader.go:190-260).  1 = a set was returned, 0 = end of file, < 0 = error.
int lc_reader_next_set(lc_reader *r, int trim, lc_readset_t *out)
{
	if (!r || !out || trim < 0) { lc_api_set_error("bad arguments"); return LC_EARG; }
	memset(out, 0, sizeof(*out));
	if (r->deferred_eof) return 0;                            // :192-194
	std::vector<Record> &recs = r->recs;
	size_t n = 0;
	auto slot = [&](size_t i) -> Record & { if (recs.size() <= i) recs.resize(i + 1); return recs[i]; };
	if (r->have_pending) { std::swap(slot(0), r->pending); r->have_pending = false; n = 1; }      // :208-213
	bool new_barcode = false, hit_eof = false;
	size_t index = n;
	for (; index < 30000; ++index) {                          // :216
		Record &rec = slot(index);
		n = index + 1;
		if (!r->read_record(rec)) {                           // :218-233
			if (index == 0) { r->deferred_eof = true; return 0; }
			r->deferred_eof = true; hit_eof = true;
			break;
		}
		if (!same_barcode(recs[0], rec) || (not_whitelist(recs[0]) && index >= 200)) {               // :235-243
			std::swap(r->pending, rec); r->have_pending = true;
			new_barcode = true;
			break;
		} else if (r->have_last && index >= 200 &&
				   recs[0].bc10x[1] - recs[0].bc10x[0] == r->last_barcode.size() &&
				   memcmp(recs[0].buf.data() + recs[0].bc10x[0], r->last_barcode.data(), r->last_barcode.size()) == 0) {   // :244-248 "abnormal break"
			break;
		}
	}
	if (n > 0) { r->last_barcode.assign(recs[0].buf.data() + recs[0].bc10x[0], recs[0].bc10x[1] - recs[0].bc10x[0]); r->have_last = true; }   // :251-255
	size_t end = n;                                           // :262-268
	int complete;
	if (new_barcode || hit_eof) { end -= 1; complete = 1; } else complete = 0;
	// ---- pack
	SetBuf &o = r->out;
	o.clear();
	for (size_t i = 0; i < end; ++i) {
		const Record &rec = recs[i];
		const char *b = rec.buf.data();
		const uint32_t l1 = rec.f[0][1] - rec.f[0][0];
		const uint32_t tt = l1 < (uint32_t)trim ? l1 : (uint32_t)trim;      // Min(len(R1), trim) (:128)
		const uint32_t lq1 = rec.f[1][1] - rec.f[1][0];
		const uint32_t tq = lq1 < tt ? lq1 : tt;              // (the Go slice expression would panic on a quality line shorter than the trim)
		o.bases.insert(o.bases.end(), b + rec.f[0][0] + tt, b + rec.f[0][1]);
		o.quals.insert(o.quals.end(), b + rec.f[1][0] + tq, b + rec.f[1][1]);
		o.quals.resize(o.bases.size(), (uint8_t)'!');         // keep quals aligned with bases if a record's two lines differ in length
		o.off.push_back((int64_t)o.bases.size());
		o.bases.insert(o.bases.end(), b + rec.f[2][0], b + rec.f[2][1]);
		o.quals.insert(o.quals.end(), b + rec.f[3][0], b + rec.f[3][1]);
		o.quals.resize(o.bases.size(), (uint8_t)'!');
		o.off.push_back((int64_t)o.bases.size());
		auto put = [&](const char *s, const char *e) { o.text.append(s, e); o.text_off.push_back((int64_t)o.text.size()); };
		put(b + rec.f[8][0], b + rec.f[8][1]);                // LC_RD_INFO
		put(b + rec.f[9][0], b + rec.f[9][1]);                // LC_RD_RGID
		put(b + rec.bc10x[0], b + rec.bc10x[1]);              // LC_RD_BARCODE10X
		put(b + rec.rawbc[0], b + rec.rawbc[1]);              // LC_RD_RAWBARCODE10X
		put(b + rec.f[5][0], b + rec.f[5][1]);                // LC_RD_BARCODE10X_QUAL
		put(b + rec.f[6][0], b + rec.f[6][1]);                // LC_RD_SAMPLE_INDEX
		put(b + rec.f[7][0], b + rec.f[7][1]);                // LC_RD_SAMPLE_INDEX_QUAL
		put(b + rec.f[0][0], b + rec.f[0][0] + tt);           // LC_RD_TRIM_BASES
		put(b + rec.f[1][0], b + rec.f[1][0] + tq);           // LC_RD_TRIM_QUALS
	}
	o.bases.push_back(0); o.quals.push_back(0);               // never hand out a null pointer for an empty set
	out->n_records = (int32_t)end;
	out->complete = complete;
	out->bases = o.bases.data(); out->quals = o.quals.data(); out->off = o.off.data();
	out->text = o.text.data(); out->text_off = o.text_off.data();
	return 1;
}

} // extern "C"
