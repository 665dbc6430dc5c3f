/* port_util.c — oracle port: tables, options, index loader, exact-order introsort.  TEST INFRASTRUCTURE ONLY. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "port.h"

port_counters_t port_counters;

/* ASCII -> 2-bit code; restates the table at B/bntseq.c:47-64. '-' maps to 5 there; every consumer on
 * Lariat's path folds codes >=4 to 4 (B/bwamem.c:1054-1055, :1101-1102), so 4 is used directly. */
const unsigned char port_nt4_table[256] = {
#define R16 4,4,4,4,4,4,4,4,4,4,4,4,4,4,4,4
	R16, R16, R16, R16,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,4,4,4, 4,4,4,4, 4,4,4,4,
	4,0,4,1, 4,4,4,2, 4,4,4,4, 4,4,4,4,  4,4,4,4, 3,4,4,4, 4,4,4,4, 4,4,4,4,
	R16, R16, R16, R16, R16, R16, R16, R16
#undef R16
};

/* B/bwa.c:109-118 */
void port_fill_scmat(int a, int b, int8_t mat[25])
{
	for (int r = 0; r < 5; ++r)
		for (int c = 0; c < 5; ++c)
			mat[r * 5 + c] = (r == 4 || c == 4) ? -1 : (r == c ? a : -b);
}

/* B/bwamem.c:48-84 */
port_opt_t *port_opt_init(void)
{
	port_opt_t *o = calloc(1, sizeof(*o));
	o->a = 1; o->b = 4;
	o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
	o->w = 100; o->T = 30; o->zdrop = 100;
	o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5;
	o->max_mem_intv = 20; o->min_seed_len = 19; o->split_width = 10;
	o->max_occ = 500; o->max_chain_gap = 10000; o->max_ins = 10000;
	o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f; o->split_factor = 1.5f;
	o->chunk_size = 10000000; o->n_threads = 1;
	o->max_XA_hits = 5; o->max_XA_hits_alt = 200; o->max_matesw = 50;
	o->mask_level_redun = 0.95f; o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = log(o->mapQ_coef_len);
	port_fill_scmat(o->a, o->b, o->mat);
	return o;
}

/* ---- index files: .bwt (B/bwt.c:443-462), .sa (B/bwt.c:421-441), .ann/.pac (B/bntseq.c:98-206, B/bwa.c:262-289) ---- */
static void *slurp(const char *fn, size_t *sz)
{
	FILE *f = fopen(fn, "rb");
	if (!f) return 0;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	void *p = malloc(n + 8);
	if (fread(p, 1, n, f) != (size_t)n) { free(p); fclose(f); return 0; }
	fclose(f);
	*sz = n;
	return p;
}

port_idx_t *port_idx_load(const char *prefix)
{
	char fn[4096];
	size_t sz;
	port_idx_t *ix = calloc(1, sizeof(*ix));
	snprintf(fn, sizeof fn, "%s.bwt", prefix);
	uint8_t *raw = slurp(fn, &sz);
	if (!raw) { free(ix); return 0; }
	memcpy(&ix->primary, raw, 8);
	memcpy(&ix->L2[1], raw + 8, 32);
	ix->L2[0] = 0;
	ix->seq_len = ix->L2[4];
	ix->bwt_size = (sz - 40) >> 2;
	ix->bwt = malloc(ix->bwt_size * 4 + 64);
	memcpy(ix->bwt, raw + 40, ix->bwt_size * 4);
	free(raw);

	snprintf(fn, sizeof fn, "%s.sa", prefix);
	raw = slurp(fn, &sz);
	if (!raw) return 0;
	uint64_t hdr[7];
	memcpy(hdr, raw, 56); /* primary, L2[1..4], sa_intv, seq_len */
	if (hdr[0] != ix->primary || hdr[6] != ix->seq_len) { fprintf(stderr, "[port] SA/BWT mismatch\n"); return 0; }
	ix->sa_intv = (int)hdr[5];
	ix->n_sa = (ix->seq_len + ix->sa_intv) / ix->sa_intv;
	ix->sa = malloc(ix->n_sa * 8);
	ix->sa[0] = (uint64_t)-1;
	memcpy(ix->sa + 1, raw + 56, (ix->n_sa - 1) * 8);
	free(raw);

	snprintf(fn, sizeof fn, "%s.ann", prefix);
	FILE *f = fopen(fn, "r");
	if (!f) return 0;
	long long lp; unsigned seed;
	if (fscanf(f, "%lld%d%u", &lp, &ix->n_seqs, &seed) != 3) return 0;
	ix->l_pac = lp;
	ix->anns = calloc(ix->n_seqs, sizeof(port_ann_t));
	for (int i = 0; i < ix->n_seqs; ++i) {
		port_ann_t *p = &ix->anns[i];
		char name[8192], rest[8192];
		long long off;
		if (fscanf(f, "%u%8191s", &p->gi, name) != 2) return 0;
		p->name = strdup(name);
		if (!fgets(rest, sizeof rest, f)) return 0; /* remainder of the line = " comment" */
		size_t rl = strlen(rest);
		if (rl && rest[rl-1] == '\n') rest[--rl] = 0;
		p->anno = strdup(rl > 1 && strcmp(rest, " (null)") != 0 ? rest + 1 : "");
		if (fscanf(f, "%lld%d%d", &off, &p->len, &p->n_ambs) != 3) return 0;
		p->offset = off;
	}
	fclose(f);
	/* optional .alt list marks ALT contigs (B/bntseq.c:179-204) */
	snprintf(fn, sizeof fn, "%s.alt", prefix);
	if ((f = fopen(fn, "r")) != 0) {
		char line[8192];
		while (fgets(line, sizeof line, f)) {
			if (line[0] == '@') continue;
			line[strcspn(line, "\t\r\n")] = 0;
			for (int i = 0; i < ix->n_seqs; ++i)
				if (!strcmp(ix->anns[i].name, line)) ix->anns[i].is_alt = 1;
		}
		fclose(f);
	}
	snprintf(fn, sizeof fn, "%s.pac", prefix);
	raw = slurp(fn, &sz);
	if (!raw) return 0;
	ix->pac = calloc(ix->l_pac / 4 + 2, 1);
	memcpy(ix->pac, raw, (size_t)(ix->l_pac / 4 + 1) < sz ? (size_t)(ix->l_pac / 4 + 1) : sz);
	free(raw);
	return ix;
}

void port_idx_destroy(port_idx_t *ix)
{
	if (!ix) return;
	for (int i = 0; i < ix->n_seqs; ++i) { free(ix->anns[i].name); free(ix->anns[i].anno); }
	free(ix->anns); free(ix->pac); free(ix->sa); free(ix->bwt); free(ix);
}

/* ---- exact-order introsort --------------------------------------------------------------------
 * Restates ks_introsort (B/ksort.h:176-226) including its comb-sort fallback (B/ksort.h:154-175) and final
 * insertion sort (B/ksort.h:146-153).  The sort is UNSTABLE and the permutation it leaves among equal keys is
 * observable in BWA-MEM's results (chain filter, region dedup), so the comparison/swap sequence must match. */
#define EL(i) (b + (size_t)(i) * sz)
static void sw_(unsigned char *x, unsigned char *y, size_t sz, unsigned char *t) { memcpy(t, x, sz); memcpy(x, y, sz); memcpy(y, t, sz); }

static void ins_sort(unsigned char *b, long n, size_t sz, port_lt_fn lt, unsigned char *t)
{
	for (long i = 1; i < n; ++i)
		for (long j = i; j > 0 && lt(EL(j), EL(j-1)); --j) sw_(EL(j), EL(j-1), sz, t);
}

static void comb_sort(unsigned char *b, size_t n, size_t sz, port_lt_fn lt, unsigned char *t)
{
	const double shrink = 1.2473309501039786540366528676643;
	size_t gap = n;
	int swapped;
	do {
		if (gap > 2) {
			gap = (size_t)(gap / shrink);
			if (gap == 9 || gap == 10) gap = 11;
		}
		swapped = 0;
		for (size_t i = 0; i + gap < n; ++i)
			if (lt(EL(i + gap), EL(i))) { sw_(EL(i), EL(i + gap), sz, t); swapped = 1; }
	} while (swapped || gap > 2);
	if (gap != 1) ins_sort(b, (long)n, sz, lt, t);
}

void port_introsort(void *base, size_t n, size_t sz, port_lt_fn lt)
{
	unsigned char *b = base, *t = malloc(sz * 2), *pv = t + sz;
	struct { long lo, hi; int depth; } stack[128];
	int top = 0, d;
	if (n < 1) { free(t); return; }
	if (n == 2) { if (lt(EL(1), EL(0))) sw_(EL(0), EL(1), sz, t); free(t); return; }
	for (d = 2; (1ul << d) < n; ++d) ;
	d <<= 1;
	long s = 0, e = (long)n - 1;
	for (;;) {
		if (s < e) {
			if (--d == 0) { comb_sort(EL(s), (size_t)(e - s + 1), sz, lt, t); e = s; continue; }
			long i = s, j = e, k = i + ((j - i) >> 1) + 1;
			/* median of three (positions s, mid+1, e) */
			if (lt(EL(k), EL(i))) { if (lt(EL(k), EL(j))) k = j; }
			else k = lt(EL(j), EL(i)) ? i : j;
			memcpy(pv, EL(k), sz);
			if (k != e) sw_(EL(k), EL(e), sz, t);
			for (;;) {
				do ++i; while (lt(EL(i), pv));
				do --j; while (i <= j && lt(pv, EL(j)));
				if (j <= i) break;
				sw_(EL(i), EL(j), sz, t);
			}
			sw_(EL(i), EL(e), sz, t);
			if (i - s > e - i) {
				if (i - s > 16) { stack[top].lo = s; stack[top].hi = i - 1; stack[top].depth = d; ++top; }
				s = e - i > 16 ? i + 1 : e;
			} else {
				if (e - i > 16) { stack[top].lo = i + 1; stack[top].hi = e; stack[top].depth = d; ++top; }
				e = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == 0) { ins_sort(b, (long)n, sz, lt, t); free(t); return; }
			--top; s = stack[top].lo; e = stack[top].hi; d = stack[top].depth;
		}
	}
}
