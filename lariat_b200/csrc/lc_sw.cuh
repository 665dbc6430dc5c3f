// lc_sw.cuh — the three Smith-Waterman flavours of the path, one job per call (scalar form).
//
// These are the per-thread forms used inside the per-read control kernels (extension, mate rescue, CIGAR);
// the warp-cooperative forms used by the throughput kernels live in lc_sw_warp.cuh and are tested against
// these and against the oracle.  Sequences are read through accessor functors so that reversed queries and
// reference windows are never materialised: the reference window is decoded from the 2-bit .pac on the fly.
#pragma once
#include "lc_fmindex.cuh"

struct LcScore {           // subset of mem_opt_t that scores cells; mat is bwa_fill_scmat(a, b) (B/bwa.c:109-118)
	int a, b, o_del, e_del, o_ins, e_ins;
	LC_HD int sub(int t, int q) const { return (t > 3 || q > 3) ? -1 : (t == q ? a : -b); }
};

struct LcSeqFwd {          // s[i] = base[i0 + i]
	const uint8_t *p;
	LC_HD int operator()(int i) const { return p[i]; }
};
struct LcSeqRev {          // s[i] = base[i0 - i]
	const uint8_t *p;      // points at element i0
	LC_HD int operator()(int i) const { return *(p - i); }
};
struct LcSeqRevComp {      // reverse complement of a read of length l: s[i] = comp(base[l-1-i])   (B/bwamem_pair.c:133-135)
	const uint8_t *p; int l;
	LC_HD int operator()(int i) const { int c = p[l - 1 - i]; return c < 4 ? 3 - c : 4; }
};
struct LcRefFwd {          // s[i] = ref[p0 + i] in doubled coordinates
	const LcIndex *ix; int64_t p0;
	LC_HD int operator()(int i) const { return lc_ref_base(*ix, p0 + i); }
};
struct LcRefRev {          // s[i] = ref[p0 - i]
	const LcIndex *ix; int64_t p0;
	LC_HD int operator()(int i) const { return lc_ref_base(*ix, p0 - i); }
};

struct LcExtResult { int score, qle, tle, gtle, gscore, max_off; };

// Banded extension from a seed with z-drop and zero-pruned band (ksw_extend2, B/ksw.c:380-479).
// H/E: scratch of qlen + 2 ints each.  end_bonus only bounds the band (B/ksw.c:402-407).
template <class QF, class TF>
LC_HDN LcExtResult lc_extend2(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int end_bonus, int zdrop, int h0,
							  int *H, int *E, LcCounters *cnt)
{
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins;
	int best = h0, best_i = -1, best_j = -1, g_i = -1, gscore = -1, max_off = 0, beg = 0, end = qlen, i, j;
	unsigned long long cells = 0;
	for (j = 0; j <= qlen + 1; ++j) H[j] = E[j] = 0;
	H[0] = h0;
	H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j - 1] > sc.e_ins; ++j) H[j] = H[j - 1] - sc.e_ins;
	{
		int smax = sc.a > 0 ? sc.a : 0;   // max entry of the 5x5 matrix: a, -b or -1
		int mi = (int)((double)(qlen * smax + end_bonus - sc.o_ins) / sc.e_ins + 1.);
		int md = (int)((double)(qlen * smax + end_bonus - sc.o_del) / sc.e_del + 1.);
		w = lc_min(w, lc_max(mi, 1));
		w = lc_min(w, lc_max(md, 1));
	}
	for (i = 0; i < tlen; ++i) {
		const int tb = target(i);
		int f = 0, hleft = 0, rowmax = 0, rowarg = -1;
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		if (beg == 0) { hleft = h0 - (sc.o_del + sc.e_del * (i + 1)); if (hleft < 0) hleft = 0; }
		for (j = beg; j < end; ++j) {
			int diag = H[j], e = E[j], M, h, t;
			H[j] = hleft;
			M = diag ? diag + sc.sub(tb, query(j)) : 0;
			h = lc_max(M, e); h = lc_max(h, f);
			hleft = h;
			if (!(rowmax > h)) rowarg = j;
			rowmax = lc_max(rowmax, h);
			t = lc_max(M - oe_del, 0);
			E[j] = lc_max(e - sc.e_del, t);
			t = lc_max(M - oe_ins, 0);
			f = lc_max(f - sc.e_ins, t);
		}
		cells += (unsigned long long)(end > beg ? end - beg : 0);
		H[end] = hleft; E[end] = 0;
		if (j == qlen) {
			if (!(gscore > hleft)) g_i = i;
			gscore = lc_max(gscore, hleft);
		}
		if (rowmax == 0) break;
		if (rowmax > best) {
			best = rowmax; best_i = i; best_j = rowarg;
			int off = rowarg - i; if (off < 0) off = -off;
			max_off = lc_max(max_off, off);
		} else if (zdrop > 0) {
			int di = i - best_i, dj = rowarg - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * sc.e_del > zdrop) break; }
			else { if (best - rowmax - (dj - di) * sc.e_ins > zdrop) break; }
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) ;
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) ;
		end = lc_min(j + 2, qlen);
	}
	if (cnt) { lc_atomic_add(&cnt->ext_calls, 1ull); lc_atomic_add(&cnt->ext_cells, cells); }
	LcExtResult r = { best, best_j + 1, best_i + 1, g_i + 1, gscore, max_off };
	return r;
}

// ---- local SW with the reference's striped-SSE2 semantics (ksw_u8/ksw_i16/ksw_align2, B/ksw.c:63-365) -----------------
// The SSE2 code is a Farrar-striped affine-gap local SW whose cell values equal the plain recurrences
//   H(i,j) = max(0, H(i-1,j-1)+s, E(i,j), F(i,j));  E(i+1,j) = max(0, E-e_del, H-o_del-e_del);  F(i,j+1) likewise
// PROVIDED the query is padded to P = ceil(qlen/lanes)*lanes zero-score columns (B/ksw.c:69-70,96,105): pad cells carry
// diagonal values and take part in the per-row maximum that drives the sub-optimal-score list.  (The oracle port
// emulates the lanes literally; tests compare the two on random and adversarial inputs.)
struct LcSwResult { int score, te, qe, score2, te2, tb, qb; };
#define LC_XBYTE  0x10000
#define LC_XSTOP  0x20000
#define LC_XSUBO  0x40000
#define LC_XSTART 0x80000

template <class QF, class TF>
LC_HDN LcSwResult lc_sw_pass(int qlen, QF query, int tlen, TF target, const LcScore &sc, int size, int xtra,
							 int *H, int *E, int *Hmax, unsigned long long *b, LcCounters *cnt)
{
	const int lanes = size == 1 ? 16 : 8, P = (qlen + lanes - 1) / lanes * lanes;
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins;
	const int shift = size == 1 ? lc_max(sc.b, 1) : 0;   // bias = -min(mat) (B/ksw.c:80-85); min entry is -b or -1
	const int minsc = (xtra & LC_XSUBO) ? (xtra & 0xffff) : 0x10000, endsc = (xtra & LC_XSTOP) ? (xtra & 0xffff) : 0x10000;
	LcSwResult r = { 0, -1, -1, -1, -1, -1, -1 };
	int gmax = 0, te = -1, n_b = 0, i, j, rows = 0;
	for (j = 0; j < P; ++j) H[j] = E[j] = Hmax[j] = 0;
	for (i = 0; i < tlen; ++i) {
		const int tb = target(i);
		int f = 0, hd = 0, imax = 0;
		++rows;
		for (j = 0; j < P; ++j) {
			int s = j < qlen ? sc.sub(tb, query(j)) : 0;
			int h = lc_max(hd + s, 0), e = E[j];
			h = lc_max(h, e); h = lc_max(h, f);
			hd = H[j]; H[j] = h;
			imax = lc_max(imax, h);
			E[j] = lc_max(lc_max(e - sc.e_del, h - oe_del), 0);
			f = lc_max(lc_max(f - sc.e_ins, h - oe_ins), 0);
		}
		if (imax >= minsc) {                         // run-length list of rows scoring >= minsc (B/ksw.c:192-200)
			if (n_b == 0 || (int)(unsigned)b[n_b - 1] + 1 != i) b[n_b++] = (unsigned long long)imax << 32 | (unsigned)i;
			else if ((int)(b[n_b - 1] >> 32) < imax) b[n_b - 1] = (unsigned long long)imax << 32 | (unsigned)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			for (j = 0; j < P; ++j) Hmax[j] = H[j];
			if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
		}
	}
	if (cnt) { lc_atomic_add(&cnt->sw_calls, 1ull); lc_atomic_add(&cnt->sw_cells, (unsigned long long)rows * qlen); }
	r.score = (size == 1 && gmax + shift >= 255) ? 255 : gmax;
	r.te = te;
	if (!(size == 1 && r.score == 255)) {
		int mx = -1;
		for (j = 0; j < P; ++j) if (Hmax[j] > mx) { mx = Hmax[j]; r.qe = j; }   // smallest column holding the maximum
		if (n_b) {
			int rad = (r.score + sc.a - 1) / sc.a, low = te - rad, high = te + rad;   // q->max == a
			for (i = 0; i < n_b; ++i) {
				int e = (int)(unsigned)b[i];
				if ((e < low || e > high) && (int)(b[i] >> 32) > r.score2) { r.score2 = (int)(b[i] >> 32); r.te2 = e; }
			}
		}
	}
	return r;
}

// second-pass views: reversed prefix [0..n] followed by the untouched remainder (revseq in place, B/ksw.c:357)
template <class F> struct LcPrefixRev {
	F f; int n;            // n = last index of the reversed prefix
	LC_HD int operator()(int i) const { return i <= n ? f(n - i) : f(i); }
};

// ksw_align2 as mem_matesw calls it (B/ksw.c:343-365, B/bwamem_pair.c:150-151). scratch: H,E,Hmax of P ints, b of tlen/2+2
template <class QF, class TF>
LC_HDN LcSwResult lc_sw_align2(int qlen, QF query, int tlen, TF target, const LcScore &sc, int xtra,
							   int *H, int *E, int *Hmax, unsigned long long *b, LcCounters *cnt)
{
	const int size = (xtra & LC_XBYTE) ? 1 : 2;
	LcSwResult r = lc_sw_pass(qlen, query, tlen, target, sc, size, xtra, H, E, Hmax, b, cnt);
	if ((xtra & LC_XSTART) == 0 || ((xtra & LC_XSUBO) && r.score < (xtra & 0xffff))) return r;
	LcPrefixRev<QF> rq = { query, r.qe };
	LcPrefixRev<TF> rt = { target, r.te };
	LcSwResult rr = lc_sw_pass(r.qe + 1, rq, tlen, rt, sc, size, LC_XSTOP | r.score, H, E, Hmax, b, cnt);
	if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
	return r;
}

// ---- banded global alignment with traceback (ksw_global2, B/ksw.c:504-606) ------------------------------------------------
#define LC_NEG_INF (-0x40000000)
LC_HD int lc_cigar_push(uint32_t *c, int n, int op, int len)
{
	if (n == 0 || op != (int)(c[n - 1] & 0xf)) c[n++] = (uint32_t)len << 4 | op;
	else c[n - 1] += (uint32_t)len << 4;
	return n;
}

// z: traceback scratch of n_col * tlen bytes (n_col = min(qlen, 2w+1)), or null for score only.
// cigar: output buffer (capacity >= qlen + tlen + 2 ops worst case is never needed: ops merge; caller gives cap and
// gets -1 in *n_cigar on overflow).  Returns the score.
template <class QF, class TF>
LC_HDN int lc_global2(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int *H, int *E,
					  uint8_t *z, uint32_t *cigar, int cigar_cap, int *n_cigar, LcCounters *cnt)
{
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins;
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	int i, j;
	unsigned long long cells = 0;
	if (n_cigar) *n_cigar = 0;
	H[0] = 0; E[0] = LC_NEG_INF;
	for (j = 1; j <= qlen && j <= w; ++j) { H[j] = -(sc.o_ins + sc.e_ins * j); E[j] = LC_NEG_INF; }
	for (; j <= qlen; ++j) H[j] = E[j] = LC_NEG_INF;
	for (i = 0; i < tlen; ++i) {
		const int tb = target(i);
		int f = LC_NEG_INF, beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int hleft = beg == 0 ? -(sc.o_del + sc.e_del * (i + 1)) : LC_NEG_INF;
		uint8_t *zi = z ? z + (size_t)i * n_col : 0;
		for (j = beg; j < end; ++j) {
			int M = H[j] + sc.sub(tb, query(j)), e = E[j], h, t;
			unsigned d;
			H[j] = hleft;
			d = M >= e ? 0 : 1; h = M >= e ? M : e;
			d = h >= f ? d : 2; h = h >= f ? h : f;
			hleft = h;
			t = M - oe_del; e -= sc.e_del;
			if (e > t) d |= 1 << 2; else e = t;
			E[j] = e;
			t = M - oe_ins; f -= sc.e_ins;
			if (f > t) d |= 2 << 4; else f = t;
			if (zi) zi[j - beg] = (uint8_t)d;
		}
		cells += (unsigned long long)(end > beg ? end - beg : 0);
		H[end] = hleft; E[end] = LC_NEG_INF;
	}
	if (cnt) { lc_atomic_add(&cnt->glb_calls, 1ull); lc_atomic_add(&cnt->glb_cells, cells); }
	if (z && cigar) {
		int n = 0, state = 0, k;
		i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		// the reference builds the CIGAR back to front and reverses it; ops merge as they are pushed
		while (i >= 0 && k >= 0) {
			if (n + 2 > cigar_cap) { *n_cigar = -1; return H[qlen]; }
			state = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (state << 1) & 3;
			if (state == 0) { n = lc_cigar_push(cigar, n, 0, 1); --i; --k; }
			else if (state == 1) { n = lc_cigar_push(cigar, n, 2, 1); --i; }
			else { n = lc_cigar_push(cigar, n, 1, 1); --k; }
		}
		if (n + 2 > cigar_cap) { *n_cigar = -1; return H[qlen]; }
		if (i >= 0) n = lc_cigar_push(cigar, n, 2, i + 1);
		if (k >= 0) n = lc_cigar_push(cigar, n, 1, k + 1);
		for (i = 0; i < n >> 1; ++i) { uint32_t t = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = t; }
		*n_cigar = n;
	}
	return H[qlen];
}
