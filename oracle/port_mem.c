/* port_mem.c — oracle port: seeding -> chaining -> filtering -> extension -> dedup, and reg -> CIGAR.
 * TEST INFRASTRUCTURE ONLY.  Restates B/bwamem.c (lines cited per function) in plain C. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <assert.h>
#include "port.h"

#define MAX2(a, b) ((a) > (b) ? (a) : (b))
#define MIN2(a, b) ((a) < (b) ? (a) : (b))

typedef struct { int64_t rbeg; int32_t qbeg, len; int score; } seed_t;        /* B/bwamem.c:168-172 */
typedef struct {                                                              /* B/bwamem.c:174-180 */
	int n, m, first, rid;
	uint32_t w, kept, is_alt;
	float frac_rep;
	int64_t pos;
	seed_t *seeds;
} chain_t;

/* ---- ordered chain set ---------------------------------------------------------------------------
 * The reference keeps chains in a B-tree keyed by `pos` (kbtree.h, t = 5 for 40-byte keys in 512-byte nodes:
 * B/kbtree.h:59, B/bwamem.c:187,261).  Keys may repeat; which twin a lookup returns and where a new twin lands
 * in the final in-order walk depend on the tree's shape, so the tree is emulated node for node.
 * Keys are indices into a chain array (the reference stores chain structs by value; moving them between nodes
 * never changes their content, so an index is equivalent). */
#define BT_T 5
#define BT_MAXK (2 * BT_T - 1)
typedef struct { int n, internal; int key[BT_MAXK]; int child[BT_MAXK + 1]; } bt_node_t;
typedef struct { bt_node_t *nd; int n_nodes, cap, root, n_keys; const chain_t *ch; } bt_t;

static int bt_new(bt_t *b, int internal)
{
	if (b->n_nodes == b->cap) { b->cap = b->cap ? b->cap << 1 : 16; b->nd = realloc(b->nd, b->cap * sizeof(bt_node_t)); }
	memset(&b->nd[b->n_nodes], 0, sizeof(bt_node_t));
	b->nd[b->n_nodes].internal = internal;
	return b->n_nodes++;
}
static inline int bt_cmp(const bt_t *b, int64_t pos, int key) /* chain_cmp(a=query, b=key), B/bwamem.c:186 */
{
	int64_t kp = b->ch[key].pos;
	return (kp < pos) - (pos < kp);
}
/* __kb_getp_aux (B/kbtree.h:117-131): index of the first key >= pos if it equals pos (*r = 0), else of the last key < pos */
static int bt_find(const bt_t *b, const bt_node_t *x, int64_t pos, int *r)
{
	int lo = 0, hi = x->n, rr;
	if (x->n == 0) return -1;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if (bt_cmp(b, pos, x->key[mid]) > 0) lo = mid + 1;   /* key[mid] < pos */
		else hi = mid;
	}
	if (lo == x->n) { if (r) *r = 1; return x->n - 1; }
	rr = bt_cmp(b, pos, x->key[lo]);
	if (r) *r = rr;
	return rr < 0 ? lo - 1 : lo;
}
/* kb_intervalp, lower bound only (B/kbtree.h:151-168) */
static int bt_lower(const bt_t *b, int64_t pos)
{
	int x = b->root, lower = -1;
	for (;;) {
		const bt_node_t *nd = &b->nd[x];
		int r = 0, i = bt_find(b, nd, pos, &r);
		if (i >= 0 && r == 0) return nd->key[i];
		if (i >= 0) lower = nd->key[i];
		if (!nd->internal) return lower;
		x = nd->child[i + 1];
	}
}
/* __kb_split (B/kbtree.h:176-191): child y = x->child[i] is full; move its upper half to a new node z */
static void bt_split(bt_t *b, int xi, int i, int yi)
{
	int zi = bt_new(b, b->nd[yi].internal);
	bt_node_t *x = &b->nd[xi], *y = &b->nd[yi], *z = &b->nd[zi];
	z->n = BT_T - 1;
	memcpy(z->key, y->key + BT_T, sizeof(int) * (BT_T - 1));
	if (y->internal) memcpy(z->child, y->child + BT_T, sizeof(int) * BT_T);
	y->n = BT_T - 1;
	memmove(x->child + i + 2, x->child + i + 1, sizeof(int) * (x->n - i));
	x->child[i + 1] = zi;
	memmove(x->key + i + 1, x->key + i, sizeof(int) * (x->n - i));
	x->key[i] = y->key[BT_T - 1];
	++x->n;
}
/* __kb_putp_aux / kb_putp (B/kbtree.h:192-224) */
static void bt_put(bt_t *b, int key)
{
	int64_t pos = b->ch[key].pos;
	int xi;
	++b->n_keys;
	if (b->nd[b->root].n == BT_MAXK) {
		int s = bt_new(b, 1);
		b->nd[s].child[0] = b->root;
		bt_split(b, s, 0, b->root);
		b->root = s;
	}
	xi = b->root;
	for (;;) {
		bt_node_t *x = &b->nd[xi];
		int i = bt_find(b, x, pos, 0);
		if (!x->internal) {
			if (i != x->n - 1) memmove(x->key + i + 2, x->key + i + 1, sizeof(int) * (x->n - i - 1));
			x->key[i + 1] = key;
			++x->n;
			return;
		}
		++i;
		if (b->nd[x->child[i]].n == BT_MAXK) {
			bt_split(b, xi, i, x->child[i]);
			x = &b->nd[xi];
			if (bt_cmp(b, pos, x->key[i]) > 0) ++i;
		}
		xi = x->child[i];
	}
}
static void bt_walk(const bt_t *b, int xi, int *out, int *n) /* in-order (__kb_traverse, B/kbtree.h:336-358) */
{
	const bt_node_t *x = &b->nd[xi];
	for (int i = 0; i < x->n; ++i) {
		if (x->internal) bt_walk(b, x->child[i], out, n);
		out[(*n)++] = x->key[i];
	}
	if (x->internal) bt_walk(b, x->child[x->n], out, n);
}

/* test_and_merge (B/bwamem.c:190-211) */
static int try_merge(const port_opt_t *opt, int64_t l_pac, chain_t *c, const seed_t *p, int seed_rid)
{
	const seed_t *last = &c->seeds[c->n - 1], *first = &c->seeds[0];
	int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len, x, y;
	if (seed_rid != c->rid) return 0;
	if (p->qbeg >= first->qbeg && p->qbeg + p->len <= qend && p->rbeg >= first->rbeg && p->rbeg + p->len <= rend) return 1;
	if ((last->rbeg < l_pac || first->rbeg < l_pac) && p->rbeg >= l_pac) return 0;
	x = p->qbeg - last->qbeg; y = p->rbeg - last->rbeg;
	if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last->len < opt->max_chain_gap && y - last->len < opt->max_chain_gap) {
		if (c->n == c->m) { c->m <<= 1; c->seeds = realloc(c->seeds, c->m * sizeof(seed_t)); }
		c->seeds[c->n++] = *p;
		return 1;
	}
	return 0;
}

/* mem_chain (B/bwamem.c:251-315) */
static int build_chains(const port_opt_t *opt, const port_idx_t *ix, int len, const uint8_t *seq, chain_t **out)
{
	port_intvv_t iv = {0,0,0};
	chain_t *ch = 0;
	int n_ch = 0, cap = 0, b = 0, e = 0, l_rep = 0;
	bt_t bt;
	*out = 0;
	if (len < opt->min_seed_len) return 0;
	memset(&bt, 0, sizeof bt);
	bt.root = bt_new(&bt, 0);
	port_collect_intv(opt, ix, len, seq, &iv);
	for (size_t i = 0; i < iv.n; ++i) {             /* span of the read covered by over-frequent seeds */
		const port_intv_t *p = &iv.a[i];
		int sb = (int)(p->info >> 32), se = (int)(uint32_t)p->info;
		if (p->x[2] <= (uint64_t)opt->max_occ) continue;
		if (sb > e) l_rep += e - b, b = sb, e = se;
		else e = MAX2(e, se);
	}
	l_rep += e - b;
	for (size_t i = 0; i < iv.n; ++i) {
		const port_intv_t *p = &iv.a[i];
		int slen = (int)((uint32_t)p->info - (uint32_t)(p->info >> 32)), count = 0;
		int64_t step = p->x[2] > (uint64_t)opt->max_occ ? p->x[2] / opt->max_occ : 1;
		for (int64_t k = 0; k < (int64_t)p->x[2] && count < opt->max_occ; k += step, ++count) {
			seed_t s;
			int rid, add = 1;
			s.rbeg = (int64_t)port_sa(ix, p->x[0] + k);
			s.qbeg = (int)(p->info >> 32);
			s.score = s.len = slen;
			rid = port_intv2rid(ix, s.rbeg, s.rbeg + s.len);
			if (rid < 0) continue;
			if (bt.n_keys) {
				int lower;
				bt.ch = ch;
				lower = bt_lower(&bt, s.rbeg);
				if (lower >= 0 && try_merge(opt, ix->l_pac, &ch[lower], &s, rid)) add = 0;
			}
			if (add) {
				if (n_ch == cap) { cap = cap ? cap << 1 : 16; ch = realloc(ch, cap * sizeof(chain_t)); }
				chain_t *c = &ch[n_ch];
				memset(c, 0, sizeof *c);
				c->n = 1; c->m = 4; c->seeds = calloc(c->m, sizeof(seed_t)); c->seeds[0] = s;
				c->rid = rid; c->pos = s.rbeg; c->is_alt = !!ix->anns[rid].is_alt;
				bt.ch = ch;
				bt_put(&bt, n_ch++);
			}
		}
	}
	free(iv.a);
	{   /* emit in B-tree order */
		int *order = malloc(sizeof(int) * (n_ch + 1)), n = 0;
		chain_t *res = malloc(sizeof(chain_t) * (n_ch + 1));
		bt.ch = ch;
		bt_walk(&bt, bt.root, order, &n);
		assert(n == n_ch);
		for (int i = 0; i < n; ++i) { res[i] = ch[order[i]]; res[i].frac_rep = (float)l_rep / len; }
		free(order); free(ch); free(bt.nd);
		*out = res;
	}
	return n_ch;
}

/* mem_chain_weight (B/bwamem.c:213-232) */
static int chain_weight(const chain_t *c)
{
	int64_t end;
	int j, wq = 0, wr = 0;
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->qbeg >= end) wq += s->len;
		else if (s->qbeg + s->len > end) wq += s->qbeg + s->len - end;
		end = MAX2(end, (int64_t)s->qbeg + s->len);
	}
	for (j = 0, end = 0; j < c->n; ++j) {
		const seed_t *s = &c->seeds[j];
		if (s->rbeg >= end) wr += s->len;
		else if (s->rbeg + s->len > end) wr += s->rbeg + s->len - end;
		end = MAX2(end, s->rbeg + s->len);
	}
	wq = MIN2(wq, wr);
	return wq < 1 << 30 ? wq : (1 << 30) - 1;
}

static int chain_w_gt(const void *a, const void *b) { return ((const chain_t*)a)->w > ((const chain_t*)b)->w; }
#define C_BEG(c) ((c).seeds[0].qbeg)
#define C_END(c) ((c).seeds[(c).n-1].qbeg + (c).seeds[(c).n-1].len)

/* mem_chain_flt (B/bwamem.c:327-385) */
static int filter_chains(const port_opt_t *opt, int n, chain_t *a)
{
	int i, k, *kept_idx, n_kept = 0;
	if (n == 0) return 0;
	for (i = k = 0; i < n; ++i) {
		chain_t *c = &a[i];
		c->first = -1; c->kept = 0;
		c->w = chain_weight(c);
		if ((int)c->w < opt->min_chain_weight) free(c->seeds);
		else a[k++] = *c;
	}
	n = k;
	port_introsort(a, n, sizeof(chain_t), chain_w_gt);
	kept_idx = malloc(sizeof(int) * n);
	a[0].kept = 3; kept_idx[n_kept++] = 0;
	for (i = 1; i < n; ++i) {
		int large_ovlp = 0;
		for (k = 0; k < n_kept; ++k) {
			int j = kept_idx[k];
			int b_max = MAX2(C_BEG(a[j]), C_BEG(a[i])), e_min = MIN2(C_END(a[j]), C_END(a[i]));
			if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
				int li = C_END(a[i]) - C_BEG(a[i]), lj = C_END(a[j]) - C_BEG(a[j]), min_l = MIN2(li, lj);
				if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
					large_ovlp = 1;
					if (a[j].first < 0) a[j].first = i;
					if (a[i].w < a[j].w * opt->drop_ratio && (int)(a[j].w - a[i].w) >= opt->min_seed_len << 1) break;
				}
			}
		}
		if (k == n_kept) { kept_idx[n_kept++] = i; a[i].kept = large_ovlp ? 2 : 3; }
	}
	for (i = 0; i < n_kept; ++i) { chain_t *c = &a[kept_idx[i]]; if (c->first >= 0) a[c->first].kept = 1; }
	free(kept_idx);
	for (i = k = 0; i < n; ++i) {
		if (a[i].kept == 0 || a[i].kept == 3) continue;
		if (++k >= opt->max_chain_extend) break;
	}
	for (; i < n; ++i) if (a[i].kept < 3) a[i].kept = 0;
	for (i = k = 0; i < n; ++i) {
		if (a[i].kept == 0) free(a[i].seeds);
		else a[k++] = a[i];
	}
	return k;
}

/* cal_max_gap (B/bwamem.c:621-628) */
static int max_gap(const port_opt_t *opt, int qlen)
{
	int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
	int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
	int l = MAX2(l_del, l_ins);
	l = MAX2(l, 1);
	return MIN2(l, opt->w << 1);
}

static int u64_lt(const void *a, const void *b) { return *(const uint64_t*)a < *(const uint64_t*)b; }

static port_reg_t *rv_pushp(port_regv_t *v)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 2; v->a = realloc(v->a, v->m * sizeof(port_reg_t)); }
	return &v->a[v->n++];
}

/* mem_chain2aln (B/bwamem.c:632-786) */
static void chain_to_regs(const port_opt_t *opt, const port_idx_t *ix, int l_query, const uint8_t *query, const chain_t *c, port_regv_t *av)
{
	int64_t l_pac = ix->l_pac, r0 = l_pac << 1, r1 = 0;
	int i, k, rid;
	uint8_t *rseq;
	uint64_t *srt;
	if (c->n == 0) return;
	for (i = 0; i < c->n; ++i) {                    /* widest reference window any seed could extend into */
		const seed_t *t = &c->seeds[i];
		int64_t b = t->rbeg - (t->qbeg + max_gap(opt, t->qbeg));
		int tail = l_query - t->qbeg - t->len;
		int64_t e = t->rbeg + t->len + (tail + max_gap(opt, tail));
		r0 = MIN2(r0, b); r1 = MAX2(r1, e);
	}
	r0 = MAX2(r0, 0); r1 = MIN2(r1, l_pac << 1);
	if (r0 < l_pac && l_pac < r1) { if (c->seeds[0].rbeg < l_pac) r1 = l_pac; else r0 = l_pac; }
	rseq = port_fetch_seq(ix, &r0, c->seeds[0].rbeg, &r1, &rid);
	assert(c->rid == rid);
	srt = malloc(sizeof(uint64_t) * c->n);
	for (i = 0; i < c->n; ++i) srt[i] = (uint64_t)c->seeds[i].score << 32 | (uint32_t)i;
	port_introsort(srt, c->n, 8, u64_lt);
	for (k = c->n - 1; k >= 0; --k) {
		const seed_t *s = &c->seeds[(uint32_t)srt[k]];
		port_reg_t *a;
		int aw0, aw1, moff;
		for (i = 0; i < (int)av->n; ++i) {          /* is the seed already explained by an earlier region? */
			const port_reg_t *p = &av->a[i];
			int64_t rd; int qd, w, g;
			if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
			if (s->len - p->seedlen0 > .1 * l_query) continue;
			qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
			g = max_gap(opt, qd < rd ? qd : (int)rd); w = MIN2(g, p->w);
			if (qd - rd < w && rd - qd < w) break;
			qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
			g = max_gap(opt, qd < rd ? qd : (int)rd); w = MIN2(g, p->w);
			if (qd - rd < w && rd - qd < w) break;
		}
		if (i < (int)av->n) {                       /* contained: extend anyway only if an overlapping off-diagonal seed exists */
			for (i = k + 1; i < c->n; ++i) {
				const seed_t *t;
				if (srt[i] == 0) continue;
				t = &c->seeds[(uint32_t)srt[i]];
				if (t->len < s->len * .95) continue;
				if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
				if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
			}
			if (i == c->n) { srt[k] = 0; continue; }
		}
		a = rv_pushp(av);
		memset(a, 0, sizeof *a);
		a->w = aw0 = aw1 = opt->w;
		a->score = a->truesc = -1;
		a->rid = c->rid;
		if (s->qbeg) {                              /* extend to the left on reversed sequences */
			int qle, tle, gtle, gscore;
			int64_t tl = s->rbeg - r0;
			uint8_t *qs = malloc(s->qbeg), *rs = malloc(tl + 1);
			for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
			for (i = 0; i < tl; ++i) rs[i] = rseq[tl - 1 - i];
			for (i = 0; i < 2; ++i) {               /* MAX_BAND_TRY */
				int prev = a->score;
				aw0 = opt->w << i;
				a->score = port_extend2(s->qbeg, qs, (int)tl, rs, 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins,
										aw0, opt->pen_clip5, opt->zdrop, s->len * opt->a, &qle, &tle, &gtle, &gscore, &moff);
				if (a->score == prev || moff < (aw0 >> 1) + (aw0 >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip5) { a->qb = s->qbeg - qle; a->rb = s->rbeg - tle; a->truesc = a->score; }
			else { a->qb = 0; a->rb = s->rbeg - gtle; a->truesc = gscore; }
			free(qs); free(rs);
		} else { a->score = a->truesc = s->len * opt->a; a->qb = 0; a->rb = s->rbeg; }
		if (s->qbeg + s->len != l_query) {          /* extend to the right */
			int qle, tle, gtle, gscore, sc0 = a->score, qe = s->qbeg + s->len;
			int64_t re = s->rbeg + s->len - r0;
			for (i = 0; i < 2; ++i) {
				int prev = a->score;
				aw1 = opt->w << i;
				a->score = port_extend2(l_query - qe, query + qe, (int)(r1 - r0 - re), rseq + re, 5, opt->mat, opt->o_del, opt->e_del,
										opt->o_ins, opt->e_ins, aw1, opt->pen_clip3, opt->zdrop, sc0, &qle, &tle, &gtle, &gscore, &moff);
				if (a->score == prev || moff < (aw1 >> 1) + (aw1 >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - opt->pen_clip3) { a->qe = qe + qle; a->re = r0 + re + tle; a->truesc += a->score - sc0; }
			else { a->qe = l_query; a->re = r0 + re + gtle; a->truesc += gscore - sc0; }
		} else { a->qe = l_query; a->re = s->rbeg + s->len; }
		for (i = 0, a->seedcov = 0; i < c->n; ++i) {
			const seed_t *t = &c->seeds[i];
			if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) a->seedcov += t->len;
		}
		a->w = MAX2(aw0, aw1);
		a->seedlen0 = s->len;
		a->frac_rep = c->frac_rep;
	}
	free(srt); free(rseq);
}

/* bwa_gen_cigar2 (B/bwa.c:121-207) without the MD string (Lariat never reads it: gobwa.go:449-487) */
uint32_t *port_gen_cigar2(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, const port_idx_t *ix,
						  int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM)
{
	int64_t l_pac = ix->l_pac, rlen;
	uint32_t *cigar = 0;
	uint8_t *rseq;
	int i;
	if (n_cigar) *n_cigar = 0;
	if (NM) *NM = -1;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	rseq = port_get_seq(ix, rb, re, &rlen);
	if (re - rb != rlen) { free(rseq); return 0; }
	if (rb >= l_pac) {          /* reverse strand: flip both so that gaps end up left-aligned on the forward strand */
		for (i = 0; i < l_query >> 1; ++i) { uint8_t t = query[i]; query[i] = query[l_query-1-i]; query[l_query-1-i] = t; }
		for (i = 0; i < rlen >> 1; ++i) { uint8_t t = rseq[i]; rseq[i] = rseq[rlen-1-i]; rseq[rlen-1-i] = t; }
	}
	if (l_query == re - rb && w_ == 0) {            /* no gap possible */
		if (n_cigar) { cigar = malloc(8); cigar[0] = (uint32_t)l_query << 4; *n_cigar = 1; }
		for (i = 0, *score = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
	} else {
		int w, mg, mi, md, min_w, dl = abs((int)rlen - l_query);
		mi = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
		md = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
		mg = MAX2(mi, md); mg = MAX2(mg, 1);
		w = (mg + dl + 1) >> 1;
		w = MIN2(w, w_);
		min_w = dl + 3;
		w = MAX2(w, min_w);
		*score = port_global2(l_query, query, (int)rlen, rseq, 5, mat, o_del, e_del, o_ins, e_ins, w, n_cigar, n_cigar ? &cigar : 0);
	}
	if (NM && n_cigar) {
		int k, x = 0, y = 0, n_mm = 0, n_gap = 0;
		for (k = 0; k < *n_cigar; ++k) {
			int op = cigar[k] & 0xf, len = cigar[k] >> 4;
			if (op == 0) { for (i = 0; i < len; ++i) n_mm += query[x + i] != rseq[y + i]; x += len; y += len; }
			else if (op == 2) { if (k > 0 && k < *n_cigar - 1) n_gap += len; y += len; }
			else if (op == 1) { x += len; n_gap += len; }
		}
		*NM = n_mm + n_gap;
	}
	if (rb >= l_pac)
		for (i = 0; i < l_query >> 1; ++i) { uint8_t t = query[i]; query[i] = query[l_query-1-i]; query[l_query-1-i] = t; }
	free(rseq);
	return cigar;
}

/* mem_patch_reg (B/bwamem.c:406-435) */
static int patch_reg(const port_opt_t *opt, const port_idx_t *ix, uint8_t *query, const port_reg_t *a, const port_reg_t *b, int *w_)
{
	int w, score, q_s, r_s;
	double r;
	if (ix == 0 || query == 0) return 0;
	assert(a->rid == b->rid && a->rb <= b->rb);
	if (a->rb < ix->l_pac && b->rb >= ix->l_pac) return 0;
	if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
	w = (int)((a->re - b->rb) - (a->qe - b->qb));
	w = w > 0 ? w : -w;
	r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
	r = r > 0. ? r : -r;
	if (a->re < b->rb || a->qe < b->qb) { if (w > opt->w << 1 || r >= 0.05f) return 0; }
	else if (w > opt->w << 2 || r >= 0.05f * 2) return 0;
	w += a->w + b->w;
	w = MIN2(w, opt->w << 2);
	port_gen_cigar2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, ix, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, 0, 0);
	q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
	r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
	if ((double)score / MAX2(q_s, r_s) < 0.90f) return 0;
	*w_ = w;
	return score;
}

static int reg_re_lt(const void *a, const void *b) { return ((const port_reg_t*)a)->re < ((const port_reg_t*)b)->re; }
static int reg_best_lt(const void *a_, const void *b_)
{
	const port_reg_t *a = a_, *b = b_;
	return a->score > b->score || (a->score == b->score && (a->rb < b->rb || (a->rb == b->rb && a->qb < b->qb)));
}

/* mem_sort_dedup_patch (B/bwamem.c:437-489); ix == NULL disables patching (the call from mem_matesw) */
int port_sort_dedup_patch(const port_opt_t *opt, const port_idx_t *ix, uint8_t *query, int n, port_reg_t *a)
{
	int m, i, j;
	if (n <= 1) return n;
	port_introsort(a, n, sizeof(port_reg_t), reg_re_lt);
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		port_reg_t *p = &a[i];
		if (p->rid != a[i-1].rid || p->rb >= a[i-1].re + opt->max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt->max_chain_gap; --j) {
			port_reg_t *q = &a[j];
			int64_t orr, oq, mr, mq;
			int score, w;
			if (q->qe == q->qb) continue;
			orr = q->re - p->rb;
			oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
			mr = MIN2(q->re - q->rb, p->re - p->rb);
			mq = MIN2(q->qe - q->qb, p->qe - p->qb);
			if (orr > opt->mask_level_redun * mr && oq > opt->mask_level_redun * mq) {
				if (p->score < q->score) { p->qe = p->qb; break; }
				else q->qe = q->qb;
			} else if (q->rb < p->rb && (score = patch_reg(opt, ix, query, q, p, &w)) > 0) {
				p->n_comp += q->n_comp + 1;
				p->seedcov = MAX2(p->seedcov, q->seedcov);
				p->sub = MAX2(p->sub, q->sub);
				p->csub = MAX2(p->csub, q->csub);
				p->qb = q->qb; p->rb = q->rb;
				p->truesc = p->score = score;
				p->w = w;
				q->qb = q->qe;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	n = m;
	port_introsort(a, n, sizeof(port_reg_t), reg_best_lt);
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i-1].score && a[i].rb == a[i-1].rb && a[i].qb == a[i-1].qb) a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	return m;
}

/* mem_align1_core (B/bwamem.c:1048-1084). mem_flt_chained_seeds (B/bwamem.c:598-615) returns at once for reads
 * shorter than ~725 bp (its guard 5.5*ln(l) > 0.05*l, :600-602); longer reads are outside this port's scope. */
port_regv_t port_align1_core(const port_opt_t *opt, const port_idx_t *ix, int l_seq, uint8_t *seq)
{
	port_regv_t regs = {0,0,0};
	chain_t *ch;
	int n, i;
	for (i = 0; i < l_seq; ++i) if (seq[i] > 4) seq[i] = 4;
	assert(!(5.5 * log(l_seq > 0 ? l_seq : 1) <= 0.05 * l_seq) && "reads >= ~725bp need mem_flt_chained_seeds");
	n = build_chains(opt, ix, l_seq, seq, &ch);
	n = filter_chains(opt, n, ch);
	for (i = 0; i < n; ++i) { chain_to_regs(opt, ix, l_seq, seq, &ch[i], &regs); free(ch[i].seeds); }
	free(ch);
	regs.n = port_sort_dedup_patch(opt, ix, seq, (int)regs.n, regs.a);
	for (size_t k = 0; k < regs.n; ++k)
		if (regs.a[k].rid >= 0 && ix->anns[regs.a[k].rid].is_alt) regs.a[k].is_alt = 1;
	return regs;
}

/* infer_bw (B/bwamem.c:792-799) */
static int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
	int w;
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
	if (w < abs(l1 - l2)) w = abs(l1 - l2);
	return w;
}

/* mem_reg2aln (B/bwamem.c:1086-1156): CIGAR, NM, strand, position.  mapq/MD/XA are not produced
 * (Lariat reads only is_rev, NM, cigar, rid, pos: gobwa.go:449-487). */
port_aln_t port_reg2aln(const port_opt_t *opt, const port_idx_t *ix, int l_query, const uint8_t *query_, const port_reg_t *ar)
{
	port_aln_t a;
	int i, w2, tmp, qb, qe, NM, score, is_rev, last_sc = -(1 << 30);
	int64_t pos, rb, re;
	uint8_t *query;
	memset(&a, 0, sizeof a);
	if (ar == 0 || ar->rb < 0 || ar->re < 0) { a.rid = -1; a.pos = -1; a.flag |= 0x4; return a; }
	qb = ar->qb; qe = ar->qe; rb = ar->rb; re = ar->re;
	query = malloc(l_query + 1);
	for (i = 0; i < l_query; ++i) query[i] = query_[i] < 5 ? query_[i] : 4;
	if (ar->secondary >= 0) a.flag |= 0x100;
	tmp = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt->a, opt->o_del, opt->e_del);
	w2 = infer_bw(qe - qb, (int)(re - rb), ar->truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = MAX2(w2, tmp);
	if (w2 > opt->w) w2 = MIN2(w2, ar->w);
	i = 0; a.cigar = 0;
	do {
		free(a.cigar);
		w2 = MIN2(w2, opt->w << 2);
		a.cigar = port_gen_cigar2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w2, ix, qe - qb, &query[qb], rb, re, &score, &a.n_cigar, &NM);
		if (score == last_sc || w2 == opt->w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar->truesc - opt->a);
	a.NM = NM;
	is_rev = !(rb < ix->l_pac);
	pos = rb < ix->l_pac ? rb : (ix->l_pac << 1) - 1 - (re - 1);
	a.is_rev = is_rev;
	if (a.n_cigar > 0) {                            /* squeeze out a leading or trailing deletion */
		if ((a.cigar[0] & 0xf) == 2) { pos += a.cigar[0] >> 4; --a.n_cigar; memmove(a.cigar, a.cigar + 1, a.n_cigar * 4); }
		else if ((a.cigar[a.n_cigar-1] & 0xf) == 2) --a.n_cigar;
	}
	if (qb != 0 || qe != l_query) {                 /* soft clips */
		int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
		a.cigar = realloc(a.cigar, 4 * (a.n_cigar + 2));
		if (clip5) { memmove(a.cigar + 1, a.cigar, a.n_cigar * 4); a.cigar[0] = (uint32_t)clip5 << 4 | 3; ++a.n_cigar; }
		if (clip3) a.cigar[a.n_cigar++] = (uint32_t)clip3 << 4 | 3;
	}
	a.rid = port_pos2rid(ix, pos);
	assert(a.rid == ar->rid);
	a.pos = pos - ix->anns[a.rid].offset;
	a.score = ar->score; a.sub = MAX2(ar->sub, ar->csub);
	a.is_alt = ar->is_alt; a.alt_sc = ar->alt_sc;
	free(query);
	return a;
}
