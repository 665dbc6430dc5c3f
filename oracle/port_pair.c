/* port_pair.c — oracle port: mate rescue (mem_matesw, B/bwamem_pair.c:111-180).  TEST INFRASTRUCTURE ONLY. */
#include <stdlib.h>
#include <string.h>
#include "port.h"

/* mem_infer_dir (B/bwamem_pair.c:23-30): orientation class (FF, FR, RF, RR) and distance of two hits */
static int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
	int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

int port_matesw(const port_opt_t *opt, const port_idx_t *ix, const port_pes_t pes[4], const port_reg_t *a,
				int l_ms, const uint8_t *ms, port_regv_t *ma)
{
	int64_t l_pac = ix->l_pac;
	int skip[4], n = 0, r, i, rid = -1;
	for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (i = 0; i < (int)ma->n; ++i) {              /* orientations already satisfied by an existing mate hit */
		int64_t dist;
		r = infer_dir(l_pac, a->rb, ma->a[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
	if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
	for (r = 0; r < 4; ++r) {
		int is_rev, is_larger;
		uint8_t *seq, *ref = 0;
		int64_t rb, re;
		if (skip[r]) continue;
		is_rev = (r >> 1) != (r & 1);
		is_larger = !(r >> 1);
		seq = malloc(l_ms + 1);
		if (is_rev) for (i = 0; i < l_ms; ++i) seq[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
		else memcpy(seq, ms, l_ms);
		if (!is_rev) {
			rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
			re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
		} else {
			rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
			re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		if (rb < re) ref = port_fetch_seq(ix, &rb, (rb + re) >> 1, &re, &rid);
		if (a->rid == rid && re - rb >= opt->min_seed_len) {
			int xtra = PORT_XSUBO | PORT_XSTART | (l_ms * opt->a < 250 ? PORT_XBYTE : 0) | (opt->min_seed_len * opt->a);
			port_swr_t aln = port_sw_align2(l_ms, seq, (int)(re - rb), ref, 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, xtra);
			port_reg_t b;
			memset(&b, 0, sizeof b);
			if (aln.score >= opt->min_seed_len && aln.qb >= 0) {
				int at;
				b.rid = a->rid;
				b.is_alt = a->is_alt;
				b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
				b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
				b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
				b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
				b.score = aln.score;
				b.csub = aln.score2;
				b.secondary = -1;
				b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
				if (ma->n == ma->m) { ma->m = ma->m ? ma->m << 1 : 2; ma->a = realloc(ma->a, ma->m * sizeof(port_reg_t)); }
				++ma->n;
				for (at = 0; at < (int)ma->n - 1; ++at) if (ma->a[at].score < b.score) break;   /* keep sorted by score */
				for (i = (int)ma->n - 1; i > at; --i) ma->a[i] = ma->a[i-1];
				ma->a[at] = b;
			}
			++n;
		}
		if (n) ma->n = port_sort_dedup_patch(opt, 0, 0, (int)ma->n, ma->a);
		free(seq); free(ref);
	}
	return n;
}
