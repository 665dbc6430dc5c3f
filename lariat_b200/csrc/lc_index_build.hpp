// lc_index_build.hpp — BWA-compatible index construction at human-genome scale.
//
// Replaces `bwa index` = bwa_idx_build (B/bwtindex.c:251-316): pack the FASTA (bns_fasta2bntseq, B/bntseq.c:275-328), BWT of
// forward + reverse-complement text (bwt_pac2bwt, B/bwtindex.c:84-121), occ interleave (bwt_bwtupdate_core, :151-173) and the
// suffix array sampled every 32 rows (bwt_cal_sa, B/bwt.c:62-84; dump formats B/bwt.c:385-407).  The five files it writes are byte
// for byte what the reference writes for an N-free reference.
//
// The reference builds the BWT single-threaded (rope/RLE: hours for 3.1 Gb) and walks the whole text through invPsi for the SA
// samples.  Here the suffixes of the doubled text (n = 6.2e9 symbols at 3.1 Gb) are sorted in lexicographic CHUNKS that fit HBM:
//   1. a 10-mer histogram splits the suffixes into ranges of at most `chunk` suffixes;
//   2. per chunk: collect its suffix positions (descending, see below), radix-sort them by their first 32 symbols (one 64-bit key
//      fetched from the 2-bit packed text), then refine only the groups that are still tied — 16 more symbols per round, one
//      64-bit radix sort of (group slot << 32 | next 16 symbols) over the shrinking tied set;
//   3. the chunk's rows are final: emit their BWT symbols (2-bit packed) and every 32nd suffix-array value;
//   4. occ interleave + file writes.
// Only the text (1.55 GB), the BWT (1.55 GB), the SA samples (1.55 GB) and one chunk's sort buffers are resident.
// End of text: keys are zero-padded past the end ('A's), every sort is stable and the initial order is by DESCENDING position, so
// among suffixes whose padded keys stay equal the shorter one comes first — which is exactly '$' < 'A' (B/bwtindex.c:95-103 appends
// no symbol; the sentinel is implicit).  A suffix whose real symbols are all compared ("exhausted") is final and leaves its group.
//
// Everything is written as element-wise functors over a Backend that provides for-each launches, a stable 64-bit radix sort and
// scans: the product instantiates it with CUDA (lc_index_build.cu: cub::DeviceRadixSort / DeviceScan); tests/hostsim/ibsim.cpp
// instantiates it with host loops to check the logic and the file formats without a GPU (test code, never part of the product).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "lc_fmindex.cuh"

#define LC_IB_K 10                      // histogram k-mer (4^10 = 1 Mi ranges: ~6 k suffixes each at 3.1 Gb)
#define LC_IB_NBUCKET (1u << (2 * LC_IB_K))

struct LcIbText { const uint64_t *w; uint64_t n; };      // doubled text: 32 symbols per word, first symbol in the top two bits, zero past n (+2 spare words)

LC_HD uint64_t lc_ib_key(const LcIbText &t, uint64_t q)   // the 32 symbols from position q on, zero-padded past the end
{
	if (q >= t.n) return 0;
	const uint64_t i = q >> 5;
	const int s = (int)(q & 31) << 1;
	const uint64_t a = lc_ldg(t.w + i);
	return s ? (a << s) | (lc_ldg(t.w + i + 1) >> (64 - s)) : a;
}
LC_HD int lc_ib_sym(const LcIbText &t, uint64_t p) { return (int)(lc_ldg(t.w + (p >> 5)) >> (62 - ((p & 31) << 1))) & 3; }
template <typename T> LC_HD void lc_atomic_or(T *p, T v)
{
#ifdef __CUDA_ARCH__
	atomicOr(p, v);
#else
	*p |= v;
#endif
}

struct LcIbC4 { unsigned long long c[4]; };
struct LcIbC4Add { LC_HD LcIbC4 operator()(const LcIbC4 &a, const LcIbC4 &b) const { LcIbC4 r; for (int i = 0; i < 4; ++i) r.c[i] = a.c[i] + b.c[i]; return r; } };

// ---- element-wise bodies ------------------------------------------------------------------------------------------------------
struct LcIbPack {            // word i of the doubled text from the forward codes (one byte per base)
	const uint8_t *codes; uint64_t L; uint64_t *out;
	LC_HD void operator()(size_t i) const
	{
		uint64_t v = 0;
		const uint64_t p0 = (uint64_t)i << 5, n = 2 * L;
		for (int k = 0; k < 32; ++k) {
			const uint64_t p = p0 + k;
			const unsigned s = p < L ? (unsigned)lc_ldg(codes + p) : p < n ? 3u - (unsigned)lc_ldg(codes + (n - 1 - p)) : 0u;
			v = v << 2 | (s & 3);
		}
		out[i] = v;
	}
};
struct LcIbHist {            // 10-mer histogram over all suffixes; word i = positions 32 i .. 32 i + 31
	LcIbText t; unsigned long long *hist;
	LC_HD void operator()(size_t i) const
	{
		const uint64_t a = lc_ldg(t.w + i), b = lc_ldg(t.w + i + 1), p0 = (uint64_t)i << 5;
		for (int s = 0; s < 32 && p0 + s < t.n; ++s) {
			const uint64_t v = s ? (a << (2 * s)) | (b >> (64 - 2 * s)) : a;
			lc_atomic_add(hist + (v >> (64 - 2 * LC_IB_K)), 1ull);
		}
	}
};
LC_HD uint32_t lc_ib_match_mask(const LcIbText &t, size_t i, uint32_t lo, uint32_t hi)
{
	const uint64_t a = lc_ldg(t.w + i), b = lc_ldg(t.w + i + 1), p0 = (uint64_t)i << 5;
	uint32_t m = 0;
	for (int s = 0; s < 32 && p0 + s < t.n; ++s) {
		const uint64_t v = s ? (a << (2 * s)) | (b >> (64 - 2 * s)) : a;
		const uint32_t k = (uint32_t)(v >> (64 - 2 * LC_IB_K));
		if (k >= lo && k < hi) m |= 1u << s;
	}
	return m;
}
struct LcIbCount { LcIbText t; uint32_t lo, hi; uint32_t *cnt; LC_HD void operator()(size_t i) const { cnt[i] = (uint32_t)lc_popc(lc_ib_match_mask(t, i, lo, hi)); } };
struct LcIbCollect {         // positions of the chunk in DESCENDING order: P[m - 1 - ascending rank]
	LcIbText t; uint32_t lo, hi; const uint32_t *off; uint64_t *P; uint64_t m;
	LC_HD void operator()(size_t i) const
	{
		uint32_t mask = lc_ib_match_mask(t, i, lo, hi);
		uint64_t a = off[i];
		while (mask) {
			const int s = lc_ctz(mask);
			mask &= mask - 1;
			P[m - 1 - a++] = ((uint64_t)i << 5) + s;
		}
	}
};
struct LcIbKeys0 { LcIbText t; const uint64_t *P; uint64_t *key; LC_HD void operator()(size_t j) const { key[j] = lc_ib_key(t, P[j]); } };
struct LcIbKeysN {           // later rounds: group slot in the high word, the next 16 symbols in the low word
	LcIbText t; const uint64_t *P; const uint32_t *gid; uint64_t depth; uint64_t *key;
	LC_HD void operator()(size_t j) const { key[j] = (uint64_t)gid[j] << 32 | (lc_ib_key(t, P[j] + depth) >> 32); }
};
struct LcIbHeads {           // after the sort: where a group of equal keys starts; exhausted suffixes are final (singletons)
	const uint64_t *key, *P; uint64_t n, cmp_depth; int round0; uint32_t *head, *gs;
	LC_HD void operator()(size_t j) const
	{
		const bool first = j == 0;
		bool h = first || key[j] != key[j - 1] || P[j] + cmp_depth >= n || P[j - 1] + cmp_depth >= n;
		head[j] = h ? 1u : 0u;
		if (!round0) gs[j] = (first || (key[j] >> 32) != (key[j - 1] >> 32)) ? (uint32_t)j : 0u;
	}
};
struct LcIbSlot {            // the element's row inside the chunk; write it; head ? slot : 0 for the group-start scan
	const uint64_t *key, *P; const uint32_t *head, *gstart; int round0; uint64_t *sa; uint32_t *hs;
	LC_HD void operator()(size_t j) const
	{
		const uint32_t slot = round0 ? (uint32_t)j : (uint32_t)(key[j] >> 32) + ((uint32_t)j - gstart[j]);
		sa[slot] = P[j];
		hs[j] = head[j] ? slot : 0u;
	}
};
struct LcIbTied { const uint32_t *head; uint64_t t; uint32_t *tied; LC_HD void operator()(size_t j) const { tied[j] = (head[j] && (j + 1 == t || head[j + 1])) ? 0u : 1u; } };
struct LcIbCompact {
	const uint32_t *tied, *off, *ngid; const uint64_t *P; uint32_t *gid_out; uint64_t *P_out;
	LC_HD void operator()(size_t j) const { if (tied[j]) { gid_out[off[j]] = ngid[j]; P_out[off[j]] = P[j]; } }
};
// rows r0 .. r0 + m - 1 of the sorted suffix list (row 0 is the empty suffix): BWT symbol of each row into the packed raw BWT (row
// numbering, the '$' row still in place), every 32nd suffix-array value, and the row of suffix 0 (`primary`)
struct LcIbEmit {
	LcIbText t; const uint64_t *sa; uint64_t r0, m; unsigned long long *braw; uint64_t *sa_s; unsigned long long *primary;
	LC_HD void operator()(size_t i) const
	{
		const uint64_t k = (r0 >> 5) + i;
		unsigned long long v = 0;
		for (int s = 0; s < 32; ++s) {
			const uint64_t r = (k << 5) + s;
			if (r < r0 || r >= r0 + m) continue;
			const uint64_t p = sa[r - r0];
			if (p == 0) *primary = r;
			else v |= (unsigned long long)lc_ib_sym(t, p - 1) << (62 - 2 * s);
			if (s == 0) sa_s[k - 1] = p;        // r is a multiple of 32 and >= 32 (r0 >= 1): sample number r / 32
		}
		if (v) lc_atomic_or(braw + k, v);
	}
};
// word w (32 symbols) of the BWT string: the raw rows with the '$' row (`primary`) taken out
struct LcIbBwtWords {
	const unsigned long long *braw; uint64_t primary, n; uint64_t *bw;
	LC_HD void operator()(size_t w) const
	{
		const uint64_t j0 = (uint64_t)w << 5;
		uint64_t v;
		if (j0 + 31 < primary) v = braw[w];
		else if (j0 >= primary) v = (braw[w] << 2) | (braw[w + 1] >> 62);
		else {
			v = 0;
			for (int s = 0; s < 32; ++s) {
				const uint64_t j = j0 + s, r = j + (j >= primary);
				v |= ((braw[r >> 5] >> (62 - ((r & 31) << 1))) & 3ull) << (62 - 2 * s);
			}
		}
		if (j0 + 32 > n) v = j0 >= n ? 0 : v & ~(~0ull >> (2 * (n - j0)));      // nothing past the last symbol
		bw[w] = v;
	}
};
struct LcIbBlockCount {      // symbols of each kind in block b (128 symbols = 4 words)
	const uint64_t *bw; uint64_t n; LcIbC4 *cnt;
	LC_HD void operator()(size_t b) const
	{
		LcIbC4 c; c.c[0] = c.c[1] = c.c[2] = c.c[3] = 0;
		for (int q = 0; q < 4; ++q) {
			const uint64_t j0 = ((uint64_t)b * 4 + q) << 5;
			if (j0 >= n) break;
			const uint64_t v = bw[b * 4 + q], lo = v & 0x5555555555555555ull, hi = (v >> 1) & 0x5555555555555555ull;
			const int valid = (int)(n - j0 < 32 ? n - j0 : 32);
			const int n3 = lc_popcll(hi & lo), n2 = lc_popcll(hi) - n3, n1 = lc_popcll(lo) - n3;
			c.c[3] += n3; c.c[2] += n2; c.c[1] += n1; c.c[0] += valid - n3 - n2 - n1;
		}
		cnt[b] = c;
	}
};
struct LcIbInterleave {      // {u64 counts-before[4]; u32 words[8]} per 128 symbols (B/bwtindex.c:151-173)
	const uint64_t *bw; const LcIbC4 *before; uint32_t *out;
	LC_HD void operator()(size_t b) const
	{
		uint32_t *o = out + 16 * b;
		for (int c = 0; c < 4; ++c) { o[2 * c] = (uint32_t)before[b].c[c]; o[2 * c + 1] = (uint32_t)(before[b].c[c] >> 32); }
		for (int q = 0; q < 4; ++q) { const uint64_t v = bw[b * 4 + q]; o[8 + 2 * q] = (uint32_t)(v >> 32); o[8 + 2 * q + 1] = (uint32_t)v; }
	}
};

struct LcIbStats { uint64_t n = 0; int n_chunks = 0; uint64_t max_chunk = 0; int max_rounds = 0; uint64_t sorted_elems = 0; double sec_sort = 0, sec_total = 0; };

template <class BE> struct LcIbBuf {
	BE &be; void *p = nullptr;
	explicit LcIbBuf(BE &b) : be(b) {}
	~LcIbBuf() { if (p) be.dfree(p); }
	bool alloc(size_t bytes) { if (p) be.dfree(p); p = be.dmalloc(bytes ? bytes : 8); return p != nullptr; }
	void free_now() { if (p) be.dfree(p); p = nullptr; }
	template <typename T> T *as() { return (T *)p; }
};

static inline bool lc_ib_write(const std::string &fn, const void *a, size_t na, const void *b = nullptr, size_t nb = 0, const void *c = nullptr, size_t nc = 0)
{
	FILE *f = fopen(fn.c_str(), "wb");
	if (!f) return false;
	bool ok = (!na || fwrite(a, 1, na, f) == na) && (!nb || fwrite(b, 1, nb, f) == nb) && (!nc || fwrite(c, 1, nc, f) == nc);
	return fclose(f) == 0 && ok;
}

// codes[i]: the bases of contig i as 2-bit codes 0..3, one per byte (ambiguous bases are not supported: the reference draws them from
// lrand48, B/bntseq.c:261).  chunk: suffixes sorted at a time (0 = default 4e8).  Writes <prefix>.pac/.ann/.amb/.bwt/.sa.
template <class BE> int lc_ib_build(BE &be, const char *prefix, int n_contigs, const char *const *names, const uint8_t *const *codes, const int64_t *lens,
									uint64_t chunk, std::string &err, LcIbStats *st)
{
	uint64_t L = 0;
	for (int i = 0; i < n_contigs; ++i) { if (lens[i] < 0 || !codes[i] || !names[i]) { err = "bad contig"; return -3; } L += (uint64_t)lens[i]; }
	if (n_contigs <= 0 || L == 0) { err = "empty reference"; return -3; }
	const uint64_t n = 2 * L;
	if (n >= (1ull << 34)) { err = "reference too large: more than 2^34 BWT symbols"; return -3; }
	for (int i = 0; i < n_contigs; ++i) if (lens[i] > 0x7fffffffll) { err = "contig longer than 2^31 - 1 bases"; return -3; }
	if (!chunk) chunk = 400000000ull;
	const std::string pre(prefix);
	const uint64_t n_words = (n + 31) >> 5;
	LcIbBuf<BE> d_text(be), d_hist(be), d_cnt(be), d_off(be), d_braw(be), d_sas(be), d_prim(be);
	auto nomem = [&]() { err = "device allocation failed in the index builder"; return -5; };
	// ---- doubled text, 2 bits per symbol
	if (!d_text.alloc(8 * (n_words + 2))) return nomem();
	be.dzero(d_text.p, 8 * (n_words + 2));
	{
		LcIbBuf<BE> d_codes(be);
		if (!d_codes.alloc(L + 64)) return nomem();
		uint64_t o = 0;
		for (int i = 0; i < n_contigs; ++i) { if (lens[i]) be.h2d((uint8_t *)d_codes.p + o, codes[i], (size_t)lens[i]); o += (uint64_t)lens[i]; }
		LcIbPack f = { d_codes.template as<uint8_t>(), L, d_text.template as<uint64_t>() };
		be.launch(n_words, f);
		be.sync();
	}
	const LcIbText T = { d_text.template as<uint64_t>(), n };
	// ---- .pac (forward strand, four bases per byte, first base in the top bits), .ann, .amb
	{
		const uint64_t fw = (L + 31) >> 5;
		std::vector<uint64_t> h(fw + 1);
		be.d2h(h.data(), d_text.p, 8 * fw);
		be.sync();
		std::vector<uint8_t> pac((size_t)(L / 4 + 2), 0);
		const size_t nbytes = (size_t)((L + 3) / 4);
		for (size_t k = 0; k < nbytes; ++k) pac[k] = (uint8_t)(h[k >> 3] >> (56 - 8 * (k & 7)));
		if (L & 3) pac[nbytes - 1] &= (uint8_t)(0xff << (8 - 2 * (L & 3)));      // the word goes on with the reverse strand: not part of .pac
		size_t sz = nbytes;
		if ((L & 3) == 0) pac[sz++] = 0;
		pac[sz++] = (uint8_t)(L & 3);
		if (!lc_ib_write(pre + ".pac", pac.data(), sz)) { err = "cannot write " + pre + ".pac"; return -2; }
		std::string ann = std::to_string(L) + " " + std::to_string(n_contigs) + " 11\n", amb = std::to_string(L) + " " + std::to_string(n_contigs) + " 0\n";
		uint64_t o = 0;
		for (int i = 0; i < n_contigs; ++i) {
			ann += std::string("0 ") + names[i] + " (null)\n" + std::to_string(o) + " " + std::to_string(lens[i]) + " 0\n";
			o += (uint64_t)lens[i];
		}
		if (!lc_ib_write(pre + ".ann", ann.data(), ann.size()) || !lc_ib_write(pre + ".amb", amb.data(), amb.size())) { err = "cannot write " + pre + ".ann/.amb"; return -2; }
	}
	// ---- chunk plan from the 10-mer histogram
	if (!d_hist.alloc(8 * (size_t)LC_IB_NBUCKET)) return nomem();
	be.dzero(d_hist.p, 8 * (size_t)LC_IB_NBUCKET);
	{ LcIbHist f = { T, d_hist.template as<unsigned long long>() }; be.launch(n_words, f); }
	std::vector<unsigned long long> hist(LC_IB_NBUCKET);
	be.d2h(hist.data(), d_hist.p, 8 * (size_t)LC_IB_NBUCKET);
	be.sync();
	d_hist.free_now();
	struct Chunk { uint32_t lo, hi; uint64_t m; };
	std::vector<Chunk> chunks;
	{
		Chunk c = { 0, 0, 0 };
		for (uint32_t k = 0; k < LC_IB_NBUCKET; ++k) {
			if (c.m && c.m + hist[k] > chunk) { c.hi = k; chunks.push_back(c); c.lo = k; c.m = 0; }
			c.m += hist[k];
		}
		c.hi = LC_IB_NBUCKET;
		if (c.m) chunks.push_back(c);
	}
	uint64_t mmax = 0, tot = 0;
	for (auto &c : chunks) { mmax = std::max(mmax, c.m); tot += c.m; }
	if (tot != n) { err = "index builder: histogram does not add up"; return -4; }
	if (mmax >= 0xffffffffull) { err = "index builder: a 10-mer range holds more than 2^32 suffixes"; return -3; }
	if (st) { st->n = n; st->n_chunks = (int)chunks.size(); st->max_chunk = mmax; }
	// ---- outputs that stay resident: raw BWT by row (n + 1 rows), SA samples, primary
	const uint64_t n_rw = ((n + 1 + 31) >> 5) + 2, n_sa = (n + 32) / 32;
	if (!d_braw.alloc(8 * n_rw) || !d_sas.alloc(8 * n_sa) || !d_prim.alloc(8)) return nomem();
	be.dzero(d_braw.p, 8 * n_rw); be.dzero(d_sas.p, 8 * n_sa); be.dzero(d_prim.p, 8);
	// ---- per-chunk buffers
	LcIbBuf<BE> d_k0(be), d_k1(be), d_p0(be), d_p1(be), d_sa(be), d_head(be), d_gs(be), d_sc(be), d_tied(be), d_toff(be), d_g0(be);
	if (!d_cnt.alloc(4 * (n_words + 1)) || !d_off.alloc(4 * (n_words + 1)) || !d_k0.alloc(8 * mmax) || !d_k1.alloc(8 * mmax) || !d_p0.alloc(8 * mmax) ||
		!d_p1.alloc(8 * mmax) || !d_sa.alloc(8 * mmax) || !d_head.alloc(4 * (mmax + 1)) || !d_gs.alloc(4 * mmax) || !d_sc.alloc(4 * mmax) ||
		!d_tied.alloc(4 * mmax) || !d_toff.alloc(4 * mmax) || !d_g0.alloc(4 * mmax)) return nomem();
	uint64_t r0 = 1;      // row 0 is the empty suffix
	for (size_t ci = 0; ci < chunks.size(); ++ci) {
		const Chunk &c = chunks[ci];
		const uint64_t m = c.m;
		{ LcIbCount f = { T, c.lo, c.hi, d_cnt.template as<uint32_t>() }; be.launch(n_words, f); }
		be.scan_sum_u32(d_cnt.template as<uint32_t>(), d_off.template as<uint32_t>(), n_words);
		{ LcIbCollect f = { T, c.lo, c.hi, d_off.template as<uint32_t>(), d_p0.template as<uint64_t>(), m }; be.launch(n_words, f); }
		// round 0: the whole chunk by its first 32 symbols; later rounds: the tied elements by (slot, next 16 symbols)
		uint64_t t = m, depth = 0;
		uint64_t *P = d_p0.template as<uint64_t>(), *P2 = d_p1.template as<uint64_t>();
		uint32_t *gid = d_g0.template as<uint32_t>();
		int round = 0;
		for (; t > 0; ++round) {
			const int round0 = round == 0;
			if (round0) { LcIbKeys0 f = { T, P, d_k0.template as<uint64_t>() }; be.launch(t, f); }
			else { LcIbKeysN f = { T, P, gid, depth, d_k0.template as<uint64_t>() }; be.launch(t, f); }
			be.sort_pairs_u64(d_k0.template as<uint64_t>(), d_k1.template as<uint64_t>(), P, P2, t, 64);
			if (st) st->sorted_elems += t;
			const uint64_t cmp_depth = depth + (round0 ? 32 : 16);
			{ LcIbHeads f = { d_k1.template as<uint64_t>(), P2, n, cmp_depth, round0, d_head.template as<uint32_t>(), d_gs.template as<uint32_t>() }; be.launch(t, f); }
			if (!round0) be.scan_max_u32(d_gs.template as<uint32_t>(), d_sc.template as<uint32_t>(), t);
			{ LcIbSlot f = { d_k1.template as<uint64_t>(), P2, d_head.template as<uint32_t>(), d_sc.template as<uint32_t>(), round0, d_sa.template as<uint64_t>(), d_gs.template as<uint32_t>() }; be.launch(t, f); }
			be.scan_max_u32(d_gs.template as<uint32_t>(), d_sc.template as<uint32_t>(), t);          // d_sc = slot of each element's group start = its next gid
			{ LcIbTied f = { d_head.template as<uint32_t>(), t, d_tied.template as<uint32_t>() }; be.launch(t, f); }
			be.scan_sum_u32(d_tied.template as<uint32_t>(), d_toff.template as<uint32_t>(), t);
			uint32_t last[2] = { 0, 0 };
			be.d2h(&last[0], d_toff.template as<uint32_t>() + (t - 1), 4); be.d2h(&last[1], d_tied.template as<uint32_t>() + (t - 1), 4);
			be.sync();
			const uint64_t t2 = (uint64_t)last[0] + last[1];
			if (t2) { LcIbCompact f = { d_tied.template as<uint32_t>(), d_toff.template as<uint32_t>(), d_sc.template as<uint32_t>(), P2, gid, P }; be.launch(t, f); }
			t = t2;
			depth = cmp_depth;
			if (depth > n + 64) { err = "index builder: refinement does not terminate"; return -4; }
		}
		if (st) st->max_rounds = std::max(st->max_rounds, round);
		{
			LcIbEmit f = { T, d_sa.template as<uint64_t>(), r0, m, d_braw.template as<unsigned long long>(), d_sas.template as<uint64_t>(), d_prim.template as<unsigned long long>() };
			be.launch((size_t)(((r0 + m - 1) >> 5) - (r0 >> 5) + 1), f);
		}
		r0 += m;
		std::string e;
		if (!be.ok(e)) { err = "index builder: " + e; return -4; }
	}
	d_k0.free_now(); d_k1.free_now(); d_p0.free_now(); d_p1.free_now(); d_sa.free_now(); d_head.free_now(); d_gs.free_now(); d_sc.free_now();
	d_tied.free_now(); d_toff.free_now(); d_g0.free_now(); d_cnt.free_now(); d_off.free_now();
	// row 0 (the empty suffix) is preceded by the last symbol of the text
	unsigned long long primary = 0, w0 = 0;
	be.d2h(&primary, d_prim.p, 8); be.d2h(&w0, d_braw.p, 8);
	{
		uint64_t lastw = 0;
		be.d2h(&lastw, (const uint64_t *)d_text.p + ((n - 1) >> 5), 8);
		be.sync();
		w0 |= ((lastw >> (62 - (((n - 1) & 31) << 1))) & 3ull) << 62;
		be.h2d(d_braw.p, &w0, 8);
		be.sync();
	}
	if (primary == 0 || primary > n) { err = "index builder: primary row not found"; return -4; }
	// ---- .sa: primary, L2[1..4], sampling interval, n, then the samples of rows 32, 64, ...
	d_text.free_now();
	const uint64_t nb = (n + 127) >> 7;
	LcIbBuf<BE> d_bw(be), d_bc(be), d_bb(be), d_out(be);
	if (!d_bw.alloc(8 * (4 * nb + 4)) || !d_bc.alloc(sizeof(LcIbC4) * (nb + 1)) || !d_bb.alloc(sizeof(LcIbC4) * (nb + 1))) return nomem();
	be.dzero(d_bw.p, 8 * (4 * nb + 4));
	{ LcIbBwtWords f = { d_braw.template as<unsigned long long>(), primary, n, d_bw.template as<uint64_t>() }; be.launch(n_words, f); }
	be.dzero(d_bc.p, sizeof(LcIbC4) * (nb + 1));
	{ LcIbBlockCount f = { d_bw.template as<uint64_t>(), n, d_bc.template as<LcIbC4>() }; be.launch(nb, f); }
	be.scan_c4(d_bc.template as<LcIbC4>(), d_bb.template as<LcIbC4>(), nb + 1);          // entry nb = the totals
	d_braw.free_now();
	LcIbC4 totals;
	be.d2h(&totals, d_bb.template as<LcIbC4>() + nb, sizeof(totals));
	be.sync();
	uint64_t L2[5] = { 0, 0, 0, 0, 0 };
	for (int c = 0; c < 4; ++c) L2[c + 1] = L2[c] + totals.c[c];
	if (L2[4] != n) { err = "index builder: BWT symbol counts do not add up"; return -4; }
	{
		std::vector<uint64_t> sas(n_sa);
		be.d2h(sas.data(), d_sas.p, 8 * n_sa);
		be.sync();
		d_sas.free_now();
		const uint64_t hdr[7] = { primary, L2[1], L2[2], L2[3], L2[4], 32, n };
		if (!lc_ib_write(pre + ".sa", hdr, sizeof(hdr), sas.data(), 8 * (n_sa - 1))) { err = "cannot write " + pre + ".sa"; return -2; }
	}
	// ---- .bwt: primary, L2[1..4], the interleaved blocks (the last one only as far as its words exist), the closing counts
	if (!d_out.alloc(64 * nb)) return nomem();
	{ LcIbInterleave f = { d_bw.template as<uint64_t>(), d_bb.template as<LcIbC4>(), d_out.template as<uint32_t>() }; be.launch(nb, f); }
	{
		std::vector<uint32_t> body(16 * (size_t)nb);
		be.d2h(body.data(), d_out.p, 64 * nb);
		be.sync();
		const uint64_t n_w32 = (n + 15) >> 4, used = (nb - 1) * 16 + 8 + (n_w32 - (nb - 1) * 8);
		const uint64_t hdr[5] = { primary, L2[1], L2[2], L2[3], L2[4] };
		uint32_t tail[8];
		for (int c = 0; c < 4; ++c) { tail[2 * c] = (uint32_t)totals.c[c]; tail[2 * c + 1] = (uint32_t)(totals.c[c] >> 32); }
		if (!lc_ib_write(pre + ".bwt", hdr, sizeof(hdr), body.data(), 4 * used, tail, sizeof(tail))) { err = "cannot write " + pre + ".bwt"; return -2; }
	}
	std::string e;
	if (!be.ok(e)) { err = "index builder: " + e; return -4; }
	return 0;
}
