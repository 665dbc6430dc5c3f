/* port_fmindex.c — oracle port: FM-index walks and packed-reference access.  TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "port.h"

static inline void vpush(port_intvv_t *v, const port_intv_t *x)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 16; v->a = realloc(v->a, v->m * sizeof(port_intv_t)); }
	v->a[v->n++] = *x;
}

/* Count, per base, the symbols among the `n` most-significant 2-bit cells of a 32-bit BWT word.
 * The reference does this with a 256-entry byte table (B/bwt.c:42-51,165-167); the result is the same count. */
static inline void count_word(uint32_t w, int n, uint64_t c[4])
{
	uint32_t valid = n >= 16 ? 0x55555555u : (n <= 0 ? 0u : (0x55555555u & ~((1u << ((16 - n) << 1)) - 1)));
	uint32_t hi = (w >> 1) & 0x55555555u, lo = w & 0x55555555u;
	c[0] += __builtin_popcount(~(hi | lo) & valid);
	c[1] += __builtin_popcount(~hi & lo & valid);
	c[2] += __builtin_popcount(hi & ~lo & valid);
	c[3] += __builtin_popcount(hi & lo & valid);
}

/* occ(k, ·) for all four bases: number of each base in BWT[0..k].  Follows bwt_occ4 (B/bwt.c:169-186). */
void port_occ4(const port_idx_t *ix, uint64_t k, uint64_t cnt[4])
{
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= ix->primary);                       /* '$' is not stored */
	const uint32_t *blk = ix->bwt + ((k >> 7) << 4);
	memcpy(cnt, blk, 32);
	const uint32_t *sym = blk + 8;
	int r = (int)(k & 127) + 1;                    /* symbols of this block included */
	for (int wi = 0; r > 0; ++wi, r -= 16) count_word(sym[wi], r, cnt);
}

/* Bidirectional extension by each base (bwt_extend, B/bwt.c:262-275).  Also tallies the occ blocks the
 * reference's bwt_2occ4 would touch (one if both positions fall in one 128-symbol block, B/bwt.c:194). */
void port_extend(const port_idx_t *ix, const port_intv_t *in, port_intv_t out[4], int is_back)
{
	uint64_t tk[4], tl[4];
	int f = !is_back;
	uint64_t k = in->x[f] - 1, l = in->x[f] - 1 + in->x[2];
	port_occ4(ix, k, tk);
	port_occ4(ix, l, tl);
	{
		uint64_t k_ = k - (k >= ix->primary), l_ = l - (l >= ix->primary);
		port_counters.extend_calls++;
		port_counters.occ_blocks += (k == (uint64_t)-1 || l == (uint64_t)-1 || (k_ >> 7) != (l_ >> 7)) ? 2 : 1;
	}
	for (int c = 0; c < 4; ++c) {
		out[c].x[f] = ix->L2[c] + 1 + tk[c];
		out[c].x[2] = tl[c] - tk[c];
	}
	out[3].x[is_back] = in->x[is_back] + (in->x[f] <= ix->primary && in->x[f] + in->x[2] - 1 >= ix->primary);
	out[2].x[is_back] = out[3].x[is_back] + out[3].x[2];
	out[1].x[is_back] = out[2].x[is_back] + out[2].x[2];
	out[0].x[is_back] = out[1].x[is_back] + out[1].x[2];
}

static inline void set_intv(const port_idx_t *ix, int c, port_intv_t *ik) /* bwt_set_intv, B/bwt.h:80 */
{
	ik->x[0] = ix->L2[c] + 1;
	ik->x[2] = ix->L2[c + 1] - ix->L2[c];
	ik->x[1] = ix->L2[3 - c] + 1;
	ik->info = 0;
}

static void reverse_v(port_intvv_t *v)
{
	for (size_t i = 0, j = v->n; i + 1 < j; ++i) { --j; port_intv_t t = v->a[i]; v->a[i] = v->a[j]; v->a[j] = t; }
}

/* All SMEMs through position x (bwt_smem1a, B/bwt.c:289-351).  Returns the end of the longest forward match. */
int port_smem1a(const port_idx_t *ix, int len, const uint8_t *q, int x, int min_intv, uint64_t max_intv,
				port_intvv_t *mem, port_intvv_t *s0, port_intvv_t *s1)
{
	port_intv_t ik, ok[4];
	port_intvv_t *prev = s0, *curr = s1, *sw;
	int i, ret;
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	set_intv(ix, q[x], &ik);
	ik.info = x + 1;
	curr->n = 0;
	/* forward: record the interval every time its size changes */
	for (i = x + 1; i < len; ++i) {
		if (ik.x[2] < max_intv) { vpush(curr, &ik); break; }
		if (q[i] > 3) { vpush(curr, &ik); break; }
		int c = 3 - q[i];
		port_extend(ix, &ik, ok, 0);
		if (ok[c].x[2] != ik.x[2]) {
			vpush(curr, &ik);
			if (ok[c].x[2] < (uint64_t)min_intv) break;
		}
		ik = ok[c]; ik.info = i + 1;
	}
	if (i == len) vpush(curr, &ik);
	reverse_v(curr);                                /* longest match first */
	ret = (int)curr->a[0].info;
	sw = curr; curr = prev; prev = sw;
	/* backward: shrink the candidate list one base at a time, emitting matches that cannot be extended */
	for (i = x - 1; i >= -1; --i) {
		int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		curr->n = 0;
		for (size_t j = 0; j < prev->n; ++j) {
			const port_intv_t *p = &prev->a[j];
			if (c >= 0 && ik.x[2] >= max_intv) port_extend(ix, p, ok, 1);
			if (c < 0 || ik.x[2] < max_intv || ok[c].x[2] < (uint64_t)min_intv) {
				if (curr->n == 0 && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) {
					ik = *p; ik.info |= (uint64_t)(i + 1) << 32;
					vpush(mem, &ik);
				}
			} else if (curr->n == 0 || ok[c].x[2] != curr->a[curr->n - 1].x[2]) {
				ok[c].info = p->info;
				vpush(curr, &ok[c]);
			}
		}
		if (curr->n == 0) break;
		sw = curr; curr = prev; prev = sw;
	}
	reverse_v(mem);                                 /* by start coordinate */
	return ret;
}

/* LAST-like forward seed (bwt_seed_strategy1, B/bwt.c:358-379) */
int port_seed_strategy1(const port_idx_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, port_intv_t *mem)
{
	port_intv_t ik, ok[4];
	memset(mem, 0, sizeof(*mem));
	if (q[x] > 3) return x + 1;
	set_intv(ix, q[x], &ik);
	for (int i = x + 1; i < len; ++i) {
		if (q[i] > 3) return i + 1;
		int c = 3 - q[i];
		port_extend(ix, &ik, ok, 0);
		if (ok[c].x[2] < (uint64_t)max_intv && i - x >= min_len) {
			*mem = ok[c];
			mem->info = (uint64_t)x << 32 | (uint32_t)(i + 1);
			return i + 1;
		}
		ik = ok[c];
	}
	return len;
}

/* Suffix-array value by LF-walking to a sampled row (bwt_sa / bwt_invPsi / bwt_occ, B/bwt.c:53-59,86-129) */
uint64_t port_sa(const port_idx_t *ix, uint64_t k)
{
	uint64_t steps = 0, mask = (uint64_t)ix->sa_intv - 1;
	port_counters.sa_calls++;
	while (k & mask) {
		++steps;
		if (k == ix->primary) { k = 0; continue; }
		uint64_t x = k - (k > ix->primary);
		uint32_t w = ix->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)];
		int c = (w >> ((~x & 15) << 1)) & 3;
		uint64_t cnt[4];
		port_occ4(ix, k, cnt);
		k = ix->L2[c] + cnt[c];
	}
	port_counters.lf_steps += steps;
	return steps + ix->sa[k / ix->sa_intv];
}

/* Three-pass seeding (mem_collect_intv, B/bwamem.c:114-162) */
static int intv_lt(const void *a, const void *b) { return ((const port_intv_t*)a)->info < ((const port_intv_t*)b)->info; }

void port_collect_intv(const port_opt_t *opt, const port_idx_t *ix, int len, const uint8_t *seq, port_intvv_t *out)
{
	port_intvv_t m1 = {0,0,0}, t0 = {0,0,0}, t1 = {0,0,0};
	int x, split_len = (int)(opt->min_seed_len * opt->split_factor + .499);
	out->n = 0;
	for (x = 0; x < len; ) {                        /* pass 1: all SMEMs */
		if (seq[x] > 3) { ++x; continue; }
		x = port_smem1a(ix, len, seq, x, 1, 0, &m1, &t0, &t1);
		for (size_t i = 0; i < m1.n; ++i) {
			int slen = (int)((uint32_t)m1.a[i].info - (uint32_t)(m1.a[i].info >> 32));
			if (slen >= opt->min_seed_len) vpush(out, &m1.a[i]);
		}
	}
	size_t old_n = out->n;                          /* pass 2: re-seed inside long, near-unique SMEMs */
	for (size_t k = 0; k < old_n; ++k) {
		port_intv_t p = out->a[k];
		int start = (int)(p.info >> 32), end = (int32_t)p.info;
		if (end - start < split_len || p.x[2] > (uint64_t)opt->split_width) continue;
		port_smem1a(ix, len, seq, (start + end) >> 1, (int)p.x[2] + 1, 0, &m1, &t0, &t1);
		for (size_t i = 0; i < m1.n; ++i)
			if ((int)((uint32_t)m1.a[i].info - (uint32_t)(m1.a[i].info >> 32)) >= opt->min_seed_len) vpush(out, &m1.a[i]);
	}
	if (opt->max_mem_intv > 0)                      /* pass 3: LAST-like seeds */
		for (x = 0; x < len; ) {
			if (seq[x] > 3) { ++x; continue; }
			port_intv_t m;
			x = port_seed_strategy1(ix, len, seq, x, opt->min_seed_len, (int)opt->max_mem_intv, &m);
			if (m.x[2] > 0) vpush(out, &m);
		}
	port_introsort(out->a, out->n, sizeof(port_intv_t), intv_lt);
	free(m1.a); free(t0.a); free(t1.a);
}

/* ---- packed reference (B/bntseq.c:349-446) ---- */
static inline int pac_at(const uint8_t *pac, int64_t l) { return pac[l >> 2] >> ((~l & 3) << 1) & 3; }

int port_pos2rid(const port_idx_t *ix, int64_t pos_f)
{
	int lo = 0, mid = 0, hi = ix->n_seqs;
	if (pos_f >= ix->l_pac) return -1;
	while (lo < hi) {
		mid = (lo + hi) >> 1;
		if (pos_f >= ix->anns[mid].offset) {
			if (mid == ix->n_seqs - 1 || pos_f < ix->anns[mid + 1].offset) break;
			lo = mid + 1;
		} else hi = mid;
	}
	return mid;
}

static inline int64_t depos(const port_idx_t *ix, int64_t pos, int *is_rev)
{
	*is_rev = pos >= ix->l_pac;
	return *is_rev ? (ix->l_pac << 1) - 1 - pos : pos;
}

int port_intv2rid(const port_idx_t *ix, int64_t rb, int64_t re)
{
	int r, a, b;
	if (rb < ix->l_pac && re > ix->l_pac) return -2;
	a = port_pos2rid(ix, depos(ix, rb, &r));
	b = rb < re ? port_pos2rid(ix, depos(ix, re - 1, &r)) : a;
	return a == b ? a : -1;
}

uint8_t *port_get_seq(const port_idx_t *ix, int64_t beg, int64_t end, int64_t *len)
{
	int64_t L = ix->l_pac;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > L << 1) end = L << 1;
	if (beg < 0) beg = 0;
	if (!(beg >= L || end <= L)) { *len = 0; return 0; }   /* bridging the strand boundary */
	uint8_t *s = malloc(end - beg + 1);
	*len = end - beg;
	for (int64_t p = beg; p < end; ++p)
		s[p - beg] = p >= L ? 3 - pac_at(ix->pac, (L << 1) - 1 - p) : pac_at(ix->pac, p);
	return s;
}

uint8_t *port_fetch_seq(const port_idx_t *ix, int64_t *beg, int64_t mid, int64_t *end, int *rid)
{
	int is_rev;
	int64_t len;
	if (*end < *beg) { int64_t t = *beg; *beg = *end; *end = t; }
	assert(*beg <= mid && mid < *end);
	*rid = port_pos2rid(ix, depos(ix, mid, &is_rev));
	int64_t fb = ix->anns[*rid].offset, fe = fb + ix->anns[*rid].len;
	if (is_rev) { int64_t t = fb; fb = (ix->l_pac << 1) - fe; fe = (ix->l_pac << 1) - t; }
	if (*beg < fb) *beg = fb;
	if (*end > fe) *end = fe;
	uint8_t *s = port_get_seq(ix, *beg, *end, &len);
	assert(s && *end - *beg == len);
	return s;
}
