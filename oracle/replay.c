/* replay.c — CPU replay of Lariat's per-barcode candidate loop.   TEST INFRASTRUCTURE ONLY.
 *
 * Restates, in C, the Go-side control flow that sits above the cgo boundary (the Go toolchain is not
 * available in the build container):
 *   - gobwa.GoBwaMemMateSW           /root/reference/go/src/gobwa/gobwa.go:226-337
 *       pes table (FR only, low=-35 high=500)                         :229-237
 *       mem_align1_core on each non-empty end                          :243-260
 *       best score per end                                             :264-283
 *       rescue phase 1 (read1 from read2's hits, <=50 anchors, delta)  :286-300
 *       rescue phase 2 (read2 from read1's post-rescue hits)           :310-324
 *   - inference.GetAlignments -> gobwa.GoBwaSmithWaterman (mem_reg2aln on EVERY candidate)
 *                                   /root/reference/go/src/inference/lariat.go:1564-1571, gobwa.go:400-415
 *   - gobwa.SequenceConvert (nst_nt4_table)                            gobwa.go:159-167
 *
 * Built twice from this one file (oracle/Makefile):
 *   -DREPLAY_REF  : against the UNMODIFIED reference library (oracle/_ref/libbwa_ref.so)  -> kind "reference"
 *   -DREPLAY_PORT : against the plain-C restatement (oracle/liboracle_port.so)            -> kind "port"
 * Both produce the same binary dump, so `cmp` of the two files is the port's parity check and either
 * file is a golden vector for the CUDA path.
 *
 *   - the rest of inference.GetAlignments per candidate (CIGAR walk against GetSeq's reference span: matches, mismatches, indels,
 *     soft clips, mismatch positions, pos/aend)                         lariat.go:1573-1650, gobwa.go:50-80,339-371   [-x file]
 *
 * usage: replay_{ref,port} <index-prefix> <pairs.bin> <out.bin> [-t threads] [-d score_delta] [-r repeat] [-n max_pairs] [-x alnx.bin]
 * alnx.bin  : "LRX1" int32 n_pairs, then per read end: int32 n; n x {int64 pos, aend; int32 matches, mismatches, indels, indel_length,
 *             soft_clipped, soft_clipped_length, n_mm; int32 mismatchLocs[n_mm]; int32 mismatchReadLocs[n_mm]}
 * pairs.bin : "LRP1" int32 n_pairs, int32 len[2*n_pairs] (r1,r2 interleaved), then the ASCII bases back to back
 * out.bin   : "LRO1" int32 n_pairs, then per read end: int32 n; n x 88-byte alnreg; n x aln records
 *             {int64 pos; int32 rid, is_rev, NM, n_cigar; uint32 cigar[n_cigar]}
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>

#ifdef REPLAY_REF
#include "bwamem.h"
#include "bwa.h"
#include "bntseq.h"
extern mem_alnreg_v mem_align1_core(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int l_seq, char *seq, void *buf);
extern int mem_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *a, int l_ms, const uint8_t *ms, mem_alnreg_v *ma);
typedef bwaidx_t idx_t;
typedef mem_opt_t opt_t;
typedef mem_alnreg_t reg_t;
typedef mem_alnreg_v regv_t;
typedef mem_pestat_t pes_t;
typedef mem_aln_t aln_t;
#define IDX_LOAD(p) bwa_idx_load((p), BWA_IDX_ALL)
#define OPT_INIT() mem_opt_init()
#define ALIGN1(o, x, l, s) mem_align1_core((o), (x)->bwt, (x)->bns, (x)->pac, (l), (char*)(s), 0)
#define MATESW(o, x, pes, a, l, s, ma) mem_matesw((o), (x)->bns, (x)->pac, (pes), (a), (l), (s), (ma))
#define REG2ALN(o, x, l, s, ar) mem_reg2aln((o), (x)->bns, (x)->pac, (l), (const char*)(s), (ar))
#define NT4(c) nst_nt4_table[(int)(c)]
#define KIND "reference"
#define L_PAC(x) ((x)->bns->l_pac)
#define ANN_OFFSET(x, rid) ((x)->bns->anns[rid].offset)
#define FETCH_SEQ(x, beg, mid, end, rid) bns_fetch_seq((x)->bns, (x)->pac, (beg), (mid), (end), (rid))
#else
#include "port.h"
typedef port_idx_t idx_t;
typedef port_opt_t opt_t;
typedef port_reg_t reg_t;
typedef port_regv_t regv_t;
typedef port_pes_t pes_t;
typedef port_aln_t aln_t;
#define IDX_LOAD(p) port_idx_load(p)
#define OPT_INIT() port_opt_init()
#define ALIGN1(o, x, l, s) port_align1_core((o), (x), (l), (s))
#define MATESW(o, x, pes, a, l, s, ma) port_matesw((o), (x), (pes), (a), (l), (s), (ma))
#define REG2ALN(o, x, l, s, ar) port_reg2aln((o), (x), (l), (s), (ar))
#define NT4(c) port_nt4_table[(int)(c)]
#define KIND "port"
#define L_PAC(x) ((x)->l_pac)
#define ANN_OFFSET(x, rid) ((x)->anns[rid].offset)
#define FETCH_SEQ(x, beg, mid, end, rid) port_fetch_seq((x), (beg), (mid), (end), (rid))
#endif

typedef struct { uint8_t *buf; size_t n, m; } obuf_t;
static void ob_put(obuf_t *o, const void *p, size_t n)
{
	if (o->n + n > o->m) { o->m = (o->n + n) * 2 + 1024; o->buf = realloc(o->buf, o->m); }
	memcpy(o->buf + o->n, p, n); o->n += n;
}

static int g_n_pairs, *g_len, g_delta = 25, g_threads = 1, g_repeat = 1;
static int64_t *g_off;
static const char *g_bases;
static idx_t *g_idx;
static opt_t *g_opt;
static obuf_t *g_out;         /* one per pair */
static obuf_t *g_xout;        /* one per pair: the GetAlignments post-processing records (-x) */
static volatile int g_next;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static double g_stage[4];     /* align, matesw, reg2aln (thread 0 only, single-thread runs) */

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static uint8_t *convert(const char *s, int l) /* gobwa.go:159-167 */
{
	uint8_t *b = malloc(l + 1);
	for (int i = 0; i < l; ++i) b[i] = NT4((unsigned char)s[i]);
	return b;
}

static int best_of(const regv_t *v) { int b = 0; for (size_t i = 0; i < v->n; ++i) if (v->a[i].score > b) b = v->a[i].score; return b; }

/* inference.GetAlignments after GoBwaSmithWaterman, for one candidate (lariat.go:1573-1650).  `a` = mem_reg2aln's result,
 * ascii = the read as given (chain.read).  GetSeq = gobwa.go:50-80 (bns_fetch_seq of the contig-relative span, reverse-complemented
 * when `reversed`, zero bytes where the fetch was clamped); pos/aend = InterpretAlign (gobwa.go:352-363). */
static void getalignments_tail(obuf_t *xo, const reg_t *reg, const aln_t *a, int l, const char *ascii)
{
	static const char twoBitToSeq[4] = { 'A', 'C', 'G', 'T' }, twoBitToSeqComp[4] = { 'T', 'G', 'C', 'A' };
	int64_t l_pac = L_PAC(g_idx), coff = ANN_OFFSET(g_idx, reg->rid);
	int64_t chain_pos = reg->rb < l_pac ? reg->rb - coff : l_pac * 2 - 1 - reg->rb - coff;          /* gobwa.go:352-358 */
	int64_t chain_aend = reg->re < l_pac ? (int64_t)reg->re - coff : l_pac * 2 - 1 - reg->re - coff;   /* gobwa.go:359-363 */
	int reversed = a->is_rev;
	int matches = 0, indels = 0, indel_length = 0, soft_clipping = 0, soft_clipping_length = 0;
	int64_t refStart = chain_pos, refEnd = chain_aend;
	if (reversed) { refStart = chain_aend + 1; refEnd = chain_pos + 1; }
	/* GetSeq */
	int64_t n_ref = refEnd - refStart;
	char *refSeq = calloc(n_ref > 0 ? n_ref : 1, 1);
	if (n_ref > 0) {
		int64_t offstart = refStart + coff, offend = refEnd + coff;
		int rid = reg->rid;
		uint8_t *res = FETCH_SEQ(g_idx, &offstart, (offstart + offend) >> 1, &offend, &rid);
		int64_t got = offend - offstart;
		for (int64_t i = 0; i < got; ++i) {
			if (reversed) refSeq[got - i - 1] = twoBitToSeqComp[res[i]];
			else refSeq[i] = twoBitToSeq[res[i]];
		}
		free(res);
	}
	int32_t *mismatchLocs = malloc(4 * (size_t)(l + 1)), *mismatchReadLocs = malloc(4 * (size_t)(l + 1));
	int n_mm = 0;
	int64_t refSeqOffset = 0; int readOffset = 0;
	int n = a->n_cigar;
	for (int t = 0; t < n; ++t) {                              /* k from cigarStart in steps of cigarIncrement (lariat.go:1590-1596) */
		uint32_t c = a->cigar[reversed ? n - 1 - t : t];
		int op = c & 0xf, len = c >> 4;
		if (op == 0) {
			matches += len;
			for (int match = 0; match < len; ++match) {
				if (refSeqOffset + match >= n_ref) continue;
				if (readOffset + match >= l) continue;           /* the Go code panics here */
				if (refSeq[refSeqOffset + match] != ascii[readOffset + match]) {
					mismatchLocs[n_mm] = (int32_t)(reversed ? refEnd - (refSeqOffset + match) : refSeqOffset + refStart + match);
					mismatchReadLocs[n_mm] = readOffset + match;
					++n_mm;
				}
			}
			refSeqOffset += len; readOffset += len;
		} else if (op == 1) { indels += 1; indel_length += len; readOffset += len; }
		else if (op == 2) { indels += 1; indel_length += len; refSeqOffset += len; }
		else if (op == 3) { soft_clipping += 1; soft_clipping_length += len; readOffset += len; }
	}
	int mismatches = (int)a->NM - indel_length;                /* lariat.go:1633-1637 */
	matches -= mismatches;
	if (mismatches < 0) mismatches = 0;
	int64_t pos = chain_pos, aend = chain_aend;
	if (pos != -1 && reversed) { pos = chain_aend + 1; aend = chain_pos + 1; }      /* lariat.go:1645-1650 */
	int32_t h[7] = { matches, mismatches, indels, indel_length, soft_clipping, soft_clipping_length, n_mm };
	ob_put(xo, &pos, 8); ob_put(xo, &aend, 8); ob_put(xo, h, 28);
	ob_put(xo, mismatchLocs, 4 * (size_t)n_mm); ob_put(xo, mismatchReadLocs, 4 * (size_t)n_mm);
	free(mismatchLocs); free(mismatchReadLocs); free(refSeq);
}

static void dump_end(obuf_t *o, obuf_t *xo, const regv_t *v, int l, const uint8_t *seq_codes_unused, const char *ascii)
{
	int32_t n = (int32_t)v->n;
	ob_put(o, &n, 4);
	if (xo) ob_put(xo, &n, 4);
	if (n) ob_put(o, v->a, (size_t)n * sizeof(reg_t));
	for (int i = 0; i < n; ++i) {
		/* GoBwaSmithWaterman re-converts the ASCII read for every candidate (gobwa.go:401) */
		uint8_t *q = convert(ascii, l);
		aln_t a = REG2ALN(g_opt, g_idx, l, q, &v->a[i]);
		int64_t pos = a.pos; int32_t h[4] = { a.rid, (int32_t)a.is_rev, (int32_t)a.NM, a.n_cigar };
		ob_put(o, &pos, 8); ob_put(o, h, 16);
		if (a.n_cigar) ob_put(o, a.cigar, 4 * (size_t)a.n_cigar);
		if (xo) getalignments_tail(xo, &v->a[i], &a, l, ascii);
		free(a.cigar); free(q);
	}
}

static void do_pair(int p, obuf_t *o, obuf_t *xo, int timed)
{
	int l1 = g_len[2*p], l2 = g_len[2*p+1];
	const char *s1 = g_bases + g_off[2*p], *s2 = g_bases + g_off[2*p+1];
	pes_t pes[4];
	memset(pes, 0, sizeof(pes));
	pes[0].failed = 1; pes[2].failed = 1; pes[3].failed = 1;
	pes[1].low = -35; pes[1].high = 500; pes[1].failed = 0; pes[1].avg = 200.0; pes[1].std = 100.0;
	uint8_t *c1 = convert(s1, l1), *c2 = convert(s2, l2);
	regv_t r1, r2;
	memset(&r1, 0, sizeof(r1)); memset(&r2, 0, sizeof(r2));
	double t0 = timed ? now() : 0;
	if (l1 > 0) r1 = ALIGN1(g_opt, g_idx, l1, c1);
	if (l2 > 0) r2 = ALIGN1(g_opt, g_idx, l2, c2);
	double t1 = timed ? now() : 0;
	int best1 = best_of(&r1), best2 = best_of(&r2);
	/* phase 1: rescue read1 from read2's hits. The Go loop walks the PRE-rescue snapshot of read2's list. */
	int num = 0;
	for (size_t i = 0; i < r2.n && num < 50 && l1 > 0; ++i)
		if (r2.a[i].score >= best2 - g_delta) { ++num; MATESW(g_opt, g_idx, pes, &r2.a[i], l1, c1, &r1); }
	/* phase 2: anchors are read1's list AFTER phase 1, gated by the PRE-rescue best1 (gobwa.go:264,312).
	 * algns_read1 was rebuilt at gobwa.go:302-307, so its length is the post-phase-1 n and its ChainedHit
	 * pointers alias r1.a, which phase 2 does not touch. */
	num = 0;
	for (size_t i = 0; i < r1.n && num < 50 && l2 > 0; ++i)
		if (r1.a[i].score >= best1 - g_delta) { ++num; MATESW(g_opt, g_idx, pes, &r1.a[i], l2, c2, &r2); }
	double t2 = timed ? now() : 0;
	if (o) { dump_end(o, xo, &r1, l1, c1, s1); dump_end(o, xo, &r2, l2, c2, s2); }
	else { /* timing-only pass still does the reg2aln work */
		obuf_t tmp = {0,0,0}; dump_end(&tmp, 0, &r1, l1, c1, s1); dump_end(&tmp, 0, &r2, l2, c2, s2); free(tmp.buf);
	}
	double t3 = timed ? now() : 0;
	if (timed) { g_stage[0] += t1 - t0; g_stage[1] += t2 - t1; g_stage[2] += t3 - t2; }
	free(r1.a); free(r2.a); free(c1); free(c2);
}

static void *worker(void *arg)
{
	long tid = (long)arg;
	for (;;) {
		int s, e;
		pthread_mutex_lock(&g_mu);
		s = g_next; g_next += 64; /* ~ one barcode group per grab */
		pthread_mutex_unlock(&g_mu);
		if (s >= g_n_pairs) break;
		e = s + 64 < g_n_pairs ? s + 64 : g_n_pairs;
		for (int p = s; p < e; ++p) do_pair(p, g_out ? &g_out[p] : 0, g_out && g_xout ? &g_xout[p] : 0, g_threads == 1 && tid == 0);
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s <idx-prefix> <pairs.bin> <out.bin|-> [-t threads] [-d delta] [-r repeat] [-n max_pairs] [-x alnx.bin]\n", argv[0]); return 2; }
	int maxp = -1;
	const char *xpath = 0;
	for (int i = 4; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "-t")) g_threads = atoi(argv[i+1]);
		else if (!strcmp(argv[i], "-x")) xpath = argv[i+1];
		else if (!strcmp(argv[i], "-d")) g_delta = atoi(argv[i+1]);
		else if (!strcmp(argv[i], "-r")) g_repeat = atoi(argv[i+1]);
		else if (!strcmp(argv[i], "-n")) maxp = atoi(argv[i+1]);
	}
	FILE *f = fopen(argv[2], "rb");
	if (!f) { perror(argv[2]); return 1; }
	char magic[4]; int32_t n;
	if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "LRP1", 4) || fread(&n, 4, 1, f) != 1) { fprintf(stderr, "bad pairs file\n"); return 1; }
	g_len = malloc(sizeof(int) * 2 * (size_t)n);
	if (fread(g_len, 4, 2 * (size_t)n, f) != 2 * (size_t)n) return 1;
	g_off = malloc(sizeof(int64_t) * (2 * (size_t)n + 1));
	g_off[0] = 0;
	for (int i = 0; i < 2 * n; ++i) g_off[i+1] = g_off[i] + g_len[i];
	char *bases = malloc(g_off[2*n] + 1);
	if (fread(bases, 1, g_off[2*n], f) != (size_t)g_off[2*n]) return 1;
	fclose(f);
	g_bases = bases;
	g_n_pairs = (maxp >= 0 && maxp < n) ? maxp : n;
	double t0 = now();
	g_idx = IDX_LOAD(argv[1]);
	if (!g_idx) { fprintf(stderr, "index load failed\n"); return 1; }
	g_opt = OPT_INIT();
	double t_load = now() - t0;
	int want_out = strcmp(argv[3], "-") != 0;
	double best = 1e30;
	for (int rep = 0; rep < g_repeat; ++rep) {
		if (want_out && rep == 0) g_out = calloc(g_n_pairs, sizeof(obuf_t)); else g_out = 0;
		if (want_out && rep == 0 && xpath) g_xout = calloc(g_n_pairs, sizeof(obuf_t)); else g_xout = 0;
		g_next = 0;
		memset(g_stage, 0, sizeof(g_stage));
		double t1 = now();
		pthread_t *th = malloc(sizeof(pthread_t) * g_threads);
		for (long t = 0; t < g_threads; ++t) pthread_create(&th[t], 0, worker, (void*)t);
		for (int t = 0; t < g_threads; ++t) pthread_join(th[t], 0);
		free(th);
		double dt = now() - t1;
		if (dt < best) best = dt;
		if (want_out && rep == 0) {
			FILE *o = fopen(argv[3], "wb");
			int32_t np = g_n_pairs;
			fwrite("LRO1", 1, 4, o); fwrite(&np, 4, 1, o);
			for (int p = 0; p < g_n_pairs; ++p) { fwrite(g_out[p].buf, 1, g_out[p].n, o); free(g_out[p].buf); }
			fclose(o); free(g_out); g_out = 0;
			if (g_xout) {
				FILE *x = fopen(xpath, "wb");
				fwrite("LRX1", 1, 4, x); fwrite(&np, 4, 1, x);
				for (int p = 0; p < g_n_pairs; ++p) { fwrite(g_xout[p].buf, 1, g_xout[p].n, x); free(g_xout[p].buf); }
				fclose(x); free(g_xout); g_xout = 0;
			}
		}
	}
	fprintf(stderr, "{\"kind\": \"%s\", \"pairs\": %d, \"threads\": %d, \"seconds\": %.4f, \"pairs_per_s\": %.1f, \"index_load_s\": %.2f, "
			"\"stage_s\": {\"align\": %.3f, \"matesw\": %.3f, \"reg2aln\": %.3f}}\n",
			KIND, g_n_pairs, g_threads, best, g_n_pairs / best, t_load, g_stage[0], g_stage[1], g_stage[2]);
	return 0;
}
