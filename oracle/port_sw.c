/* port_sw.c — oracle port: the three Smith-Waterman flavours on Lariat's path.  TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "port.h"

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

/* ---- banded extension from a seed (ksw_extend2, B/ksw.c:380-479) --------------------------------
 * Row i = target base, column j = query base.  Per column we keep H(i-1,j-1) and E(i,j); F runs along the
 * row.  Gaps open from M (the diagonal move) only.  The active column range [beg,end) shrinks where H and E
 * are both zero, is clamped by the band, and the loop stops on an all-zero row or on z-drop. */
int port_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
				 int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
				 int *qle_, int *tle_, int *gtle_, int *gscore_, int *max_off_)
{
	int *H = calloc(qlen + 2, sizeof(int)), *E = calloc(qlen + 2, sizeof(int));
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int best = h0, best_i = -1, best_j = -1, g_i = -1, gscore = -1, max_off = 0, beg = 0, end = qlen, i, j, smax = 0;
	uint64_t cells = 0;
	assert(h0 > 0);
	/* row -1: h0 decays along an insertion (B/ksw.c:395-397) */
	H[0] = h0;
	H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j-1] > e_ins; ++j) H[j] = H[j-1] - e_ins;
	/* band cannot usefully exceed the longest gap the score can pay for (B/ksw.c:399-407) */
	for (i = 0; i < m * m; ++i) smax = MAX2(smax, mat[i]);
	{
		int mi = (int)((double)(qlen * smax + end_bonus - o_ins) / e_ins + 1.);
		int md = (int)((double)(qlen * smax + end_bonus - o_del) / e_del + 1.);
		w = MIN2(w, MAX2(mi, 1));
		w = MIN2(w, MAX2(md, 1));
	}
	for (i = 0; i < tlen; ++i) {
		const int8_t *sc = mat + target[i] * m;
		int f = 0, hleft, rowmax = 0, rowarg = -1;
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		hleft = 0;
		if (beg == 0) { hleft = h0 - (o_del + e_del * (i + 1)); if (hleft < 0) hleft = 0; }
		for (j = beg; j < end; ++j) {
			int diag = H[j], e = E[j], M, h, t;
			H[j] = hleft;                                   /* becomes H(i,j-1) = next row's diagonal for column j */
			M = diag ? diag + sc[query[j]] : 0;
			h = MAX2(M, e); h = MAX2(h, f);
			hleft = h;
			if (!(rowmax > h)) rowarg = j;                  /* last column attaining the row maximum */
			rowmax = MAX2(rowmax, h);
			t = M - oe_del; t = MAX2(t, 0);
			e -= e_del; E[j] = MAX2(e, t);
			t = M - oe_ins; t = MAX2(t, 0);
			f -= e_ins; f = MAX2(f, t);
		}
		cells += (uint64_t)(end > beg ? end - beg : 0);
		H[end] = hleft; E[end] = 0;
		if (j == qlen) {                                    /* reached the query end: candidate to-end score */
			if (!(gscore > hleft)) g_i = i;
			gscore = MAX2(gscore, hleft);
		}
		if (rowmax == 0) break;
		if (rowmax > best) {
			best = rowmax; best_i = i; best_j = rowarg;
			max_off = MAX2(max_off, abs(rowarg - i));
		} else if (zdrop > 0) {
			int di = i - best_i, dj = rowarg - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
		}
		for (j = beg; j < end && H[j] == 0 && E[j] == 0; ++j) ;
		beg = j;
		for (j = end; j >= beg && H[j] == 0 && E[j] == 0; --j) ;
		end = MIN2(j + 2, qlen);
	}
	free(H); free(E);
	port_counters.ext_calls++; port_counters.ext_cells += cells;
	if (qle_) *qle_ = best_j + 1;
	if (tle_) *tle_ = best_i + 1;
	if (gtle_) *gtle_ = g_i + 1;
	if (gscore_) *gscore_ = gscore;
	if (max_off_) *max_off_ = max_off;
	return best;
}

/* ---- local SW with the reference's striped-SIMD semantics (ksw_u8 / ksw_i16 / ksw_align2, B/ksw.c:63-365) ----
 * Emulated lane by lane: `lanes` = 16 (8-bit, biased by `shift`, saturating) or 8 (16-bit).  The query is padded
 * to slen*lanes columns with zero scores (B/ksw.c:96,105); pad columns take part in the per-row maximum. */
typedef struct {
	int qlen, slen, lanes, shift, mdiff, max, size;
	int *qp;            /* [m][slen][lanes] */
	int *H0, *H1, *E, *Hmax;
} sq_t;

static sq_t *sq_init(int size, int qlen, const uint8_t *query, int m, const int8_t *mat)
{
	sq_t *q = calloc(1, sizeof(*q));
	int lanes = size == 1 ? 16 : 8, slen = (qlen + lanes - 1) / lanes, mn = 127, mx = 0;
	q->size = size; q->qlen = qlen; q->slen = slen; q->lanes = lanes;
	for (int a = 0; a < m * m; ++a) { if (mat[a] < mn) mn = mat[a]; if (mat[a] > mx) mx = mat[a]; }
	q->max = mx;
	q->shift = (256 - mn) & 0xff;       /* uint8 wrap of 256 - (negative min) (B/ksw.c:85) */
	q->mdiff = (mx + q->shift) & 0xff;
	q->qp = malloc(sizeof(int) * m * slen * lanes);
	q->H0 = calloc(4 * (size_t)slen * lanes, sizeof(int));
	q->H1 = q->H0 + slen * lanes; q->E = q->H1 + slen * lanes; q->Hmax = q->E + slen * lanes;
	for (int a = 0; a < m; ++a)
		for (int j = 0; j < slen; ++j)
			for (int l = 0; l < lanes; ++l) {
				int k = j + l * slen;
				int s = k >= qlen ? 0 : mat[a * m + query[k]];
				q->qp[(a * slen + j) * lanes + l] = size == 1 ? s + q->shift : s;
			}
	return q;
}
static void sq_free(sq_t *q) { free(q->qp); free(q->H0); free(q); }

static inline int sat_add(int a, int b, int size) { int s = a + b; if (size == 1) return s > 255 ? 255 : s; return s > 32767 ? 32767 : (s < -32768 ? -32768 : s); }
static inline int sat_subu(int a, int b) { return a > b ? a - b : 0; }

static port_swr_t sq_run(sq_t *q, int tlen, const uint8_t *target, int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	const int slen = q->slen, L = q->lanes, size = q->size;
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	port_swr_t r = { 0, -1, -1, -1, -1, -1, -1 };
	int minsc = (xtra & PORT_XSUBO) ? xtra & 0xffff : 0x10000;
	int endsc = (xtra & PORT_XSTOP) ? xtra & 0xffff : 0x10000;
	int *H0 = q->H0, *H1 = q->H1, *E = q->E, *Hmax = q->Hmax, *tmp;
	uint64_t *b = 0; int n_b = 0, m_b = 0, gmax = 0, te = -1, i, rows = 0;
	int h[16], e[16], f[16], mx[16];
	memset(H0, 0, sizeof(int) * 4 * slen * L);
	for (i = 0; i < tlen; ++i) {
		const int *S = q->qp + (size_t)target[i] * slen * L;
		int j, l, k, imax = 0, done = 0;
		++rows;
		for (l = 0; l < L; ++l) { f[l] = 0; mx[l] = 0; h[l] = l ? H0[(slen - 1) * L + l - 1] : 0; }
		for (j = 0; j < slen; ++j) {
			for (l = 0; l < L; ++l) {
				int hv = sat_add(h[l], S[j * L + l], size), t;
				if (size == 1) hv = sat_subu(hv, q->shift);
				e[l] = E[j * L + l];
				hv = MAX2(hv, e[l]); hv = MAX2(hv, f[l]);
				mx[l] = MAX2(mx[l], hv);
				H1[j * L + l] = hv;
				e[l] = sat_subu(e[l], e_del); t = sat_subu(hv, oe_del); E[j * L + l] = MAX2(e[l], t);
				f[l] = sat_subu(f[l], e_ins); t = sat_subu(hv, oe_ins); f[l] = MAX2(f[l], t);
				h[l] = H0[j * L + l];
			}
		}
		for (k = 0; k < 16 && !done; ++k) {                 /* lazy-F: carry F across segment boundaries */
			for (l = L - 1; l > 0; --l) f[l] = f[l-1];
			f[0] = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (l = 0; l < L; ++l) {
					int hv = MAX2(H1[j * L + l], f[l]);
					H1[j * L + l] = hv;
					hv = sat_subu(hv, oe_ins);
					f[l] = sat_subu(f[l], e_ins);
					if (f[l] > hv) any = 1;
				}
				if (!any) { done = 1; break; }
			}
		}
		for (l = 0; l < L; ++l) imax = MAX2(imax, mx[l]);
		if (imax >= minsc) {                                /* run-length list of high-scoring rows (B/ksw.c:192-200) */
			if (n_b == 0 || (int32_t)b[n_b-1] + 1 != i) {
				if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = realloc(b, 8 * m_b); }
				b[n_b++] = (uint64_t)imax << 32 | (uint32_t)i;
			} else if ((int)(b[n_b-1] >> 32) < imax) b[n_b-1] = (uint64_t)imax << 32 | (uint32_t)i;
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			memcpy(Hmax, H1, sizeof(int) * slen * L);
			if ((size == 1 && gmax + q->shift >= 255) || gmax >= endsc) break;
		}
		tmp = H1; H1 = H0; H0 = tmp;
	}
	port_counters.sw_calls++; port_counters.sw_cells += (uint64_t)rows * q->qlen;
	r.score = (size == 1 && gmax + q->shift >= 255) ? 255 : gmax;
	r.te = te;
	if (!(size == 1 && r.score == 255)) {
		int mxv = -1, low, high, n = slen * L;
		if (size == 2) r.qe = -1;
		for (i = 0; i < n; ++i) {                           /* memory order: vector i/L, lane i%L -> query idx */
			int v = Hmax[i], pos = i / L + i % L * slen;
			if (v > mxv) mxv = v, r.qe = pos;
			else if (v == mxv && pos < r.qe) r.qe = pos;
		}
		if (b) {
			i = (r.score + q->max - 1) / q->max;
			low = te - i; high = te + i;
			for (i = 0; i < n_b; ++i) {
				int ee = (int32_t)b[i];
				if ((ee < low || ee > high) && (int)(b[i] >> 32) > r.score2) r.score2 = b[i] >> 32, r.te2 = ee;
			}
		}
	}
	free(b);
	return r;
}

static void rev_(uint8_t *s, int l) { for (int i = 0; i < l >> 1; ++i) { uint8_t t = s[i]; s[i] = s[l-1-i]; s[l-1-i] = t; } }

port_swr_t port_sw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
						  int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
	int size = (xtra & PORT_XBYTE) ? 1 : 2;
	sq_t *q = sq_init(size, qlen, query, m, mat);
	port_swr_t r = sq_run(q, tlen, target, o_del, e_del, o_ins, e_ins, xtra), rr;
	sq_free(q);
	if ((xtra & PORT_XSTART) == 0 || ((xtra & PORT_XSUBO) && r.score < (xtra & 0xffff))) return r;
	/* second pass on the reversed prefixes finds the start (B/ksw.c:357-363) */
	rev_(query, r.qe + 1); rev_(target, r.te + 1);
	q = sq_init(size, r.qe + 1, query, m, mat);
	rr = sq_run(q, tlen, target, o_del, e_del, o_ins, e_ins, PORT_XSTOP | r.score);
	rev_(query, r.qe + 1); rev_(target, r.te + 1);
	sq_free(q);
	if (r.score == rr.score) r.tb = r.te - rr.te, r.qb = r.qe - rr.qe;
	return r;
}

/* ---- banded global alignment with traceback (ksw_global2, B/ksw.c:504-606) ---- */
#define NEG_INF (-0x40000000)
static uint32_t *cig_push(int *n, int *cap, uint32_t *c, int op, int len)
{
	if (*n == 0 || op != (int)(c[*n - 1] & 0xf)) {
		if (*n == *cap) { *cap = *cap ? *cap << 1 : 4; c = realloc(c, (size_t)*cap << 2); }
		c[(*n)++] = (uint32_t)len << 4 | op;
	} else c[*n - 1] += (uint32_t)len << 4;
	return c;
}

int port_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
				 int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t **cigar_)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int want_tb = n_cigar_ && cigar_, n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1, i, j, score;
	int *H = malloc(sizeof(int) * (qlen + 2)), *E = malloc(sizeof(int) * (qlen + 2));
	uint8_t *z = want_tb ? malloc((size_t)n_col * tlen + 1) : 0;
	uint64_t cells = 0;
	if (n_cigar_) *n_cigar_ = 0;
	H[0] = 0; E[0] = NEG_INF;
	for (j = 1; j <= qlen && j <= w; ++j) H[j] = -(o_ins + e_ins * j), E[j] = NEG_INF;
	for (; j <= qlen; ++j) H[j] = E[j] = NEG_INF;
	for (i = 0; i < tlen; ++i) {
		const int8_t *sc = mat + target[i] * m;
		int f = NEG_INF, beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int hleft = beg == 0 ? -(o_del + e_del * (i + 1)) : NEG_INF;
		uint8_t *zi = want_tb ? z + (size_t)i * n_col : 0;
		for (j = beg; j < end; ++j) {
			int M = H[j] + sc[query[j]], e = E[j], h, t;
			uint8_t d;
			H[j] = hleft;
			d = M >= e ? 0 : 1;  h = M >= e ? M : e;
			d = h >= f ? d : 2;  h = h >= f ? h : f;
			hleft = h;
			t = M - oe_del; e -= e_del;
			if (e > t) d |= 1 << 2; else e = t;
			E[j] = e;
			t = M - oe_ins; f -= e_ins;
			if (f > t) d |= 2 << 4; else f = t;
			if (zi) zi[j - beg] = d;
		}
		cells += (uint64_t)(end > beg ? end - beg : 0);
		H[end] = hleft; E[end] = NEG_INF;
	}
	score = H[qlen];
	port_counters.glb_calls++; port_counters.glb_cells += cells;
	if (want_tb) {
		int n = 0, cap = 0, state = 0, k;
		uint32_t *c = 0;
		i = tlen - 1; k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
		while (i >= 0 && k >= 0) {
			state = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (state << 1) & 3;
			if (state == 0) c = cig_push(&n, &cap, c, 0, 1), --i, --k;
			else if (state == 1) c = cig_push(&n, &cap, c, 2, 1), --i;
			else c = cig_push(&n, &cap, c, 1, 1), --k;
		}
		if (i >= 0) c = cig_push(&n, &cap, c, 2, i + 1);
		if (k >= 0) c = cig_push(&n, &cap, c, 1, k + 1);
		for (i = 0; i < n >> 1; ++i) { uint32_t t = c[i]; c[i] = c[n-1-i]; c[n-1-i] = t; }
		*n_cigar_ = n; *cigar_ = c;
	}
	free(H); free(E); free(z);
	return score;
}
