/* port.h — plain-C restatement ("port") of the BWA-MEM path Lariat drives.   TEST INFRASTRUCTURE ONLY.
 *
 * This is the CPU oracle of the cuda-lariat repo: a from-scratch restatement of the algorithms the
 * reference executes for the hot path (reference = BWA 0.7.16a-r1181 vendored under
 * /root/reference/go/src/gobwa/bwa, "B/" below).  Each function cites the reference lines it follows.
 * It is pinned (tests/test_oracle_*.py) against (1) the reference's own PhiX known answers
 * (go/src/test/gobwa_test.go:18-25), (2) byte-identical dumps of the UNMODIFIED reference library
 * (oracle/_ref, built from /root/reference by oracle/Makefile) on synthetic data, committed as
 * tests/golden/ fixtures.  Nothing in lariat_b200/ (the product) may include, link or call this.
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- index (B/bwt.h:46-58, B/bntseq.h:42-65, B/bwa.h:20-28) ---- */
typedef struct {
	int64_t offset;
	int32_t len, n_ambs;
	uint32_t gi;
	int32_t is_alt;
	char *name, *anno;
} port_ann_t;

typedef struct {
	/* FM index of forward+reverse-complement text; bwt[] is the on-disk interleaved layout:
	 * per 128 symbols a 64-byte block { uint64 occ[4]; uint32 sym[8] } (B/bwtindex.c:151-173) */
	uint64_t primary, L2[5], seq_len, bwt_size;
	uint32_t *bwt;
	int sa_intv;
	uint64_t n_sa, *sa;
	/* packed reference + annotations */
	int64_t l_pac;
	int32_t n_seqs;
	port_ann_t *anns;
	uint8_t *pac;
} port_idx_t;

/* ---- options (B/bwamem.h:25-57; defaults B/bwamem.c:48-84) ---- */
typedef struct {
	int a, b, o_del, e_del, o_ins, e_ins, pen_unpaired, pen_clip5, pen_clip3, w, zdrop;
	uint64_t max_mem_intv;
	int T, flag, min_seed_len, min_chain_weight, max_chain_extend;
	float split_factor;
	int split_width, max_occ, max_chain_gap, n_threads, chunk_size;
	float mask_level, drop_ratio, XA_drop_ratio, mask_level_redun, mapQ_coef_len;
	int mapQ_coef_fac, max_ins, max_matesw, max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
} port_opt_t;

/* ---- alignment region: byte-compatible with mem_alnreg_t (B/bwamem.h:59-77, sizeof 88) ---- */
typedef struct {
	int64_t rb, re;
	int qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0;
	int n_comp:30, is_alt:2;
	float frac_rep;
	uint64_t hash;
} port_reg_t;
typedef struct { size_t n, m; port_reg_t *a; } port_regv_t;

typedef struct { int low, high, failed; double avg, std; } port_pes_t;       /* B/bwamem.h:81-85 */

typedef struct {                                                               /* B/bwamem.h:87-97 */
	int64_t pos;
	int rid, flag;
	uint32_t is_rev:1, is_alt:1, mapq:8, NM:22;
	int n_cigar;
	uint32_t *cigar;
	char *XA;
	int score, sub, alt_sc;
} port_aln_t;

/* ---- FM-index primitives ---- */
typedef struct { uint64_t x[3], info; } port_intv_t;                           /* B/bwt.h:60-62 */
typedef struct { size_t n, m; port_intv_t *a; } port_intvv_t;

extern const unsigned char port_nt4_table[256];

port_idx_t *port_idx_load(const char *prefix);
void port_idx_destroy(port_idx_t *ix);
port_opt_t *port_opt_init(void);
void port_fill_scmat(int a, int b, int8_t mat[25]);

void port_occ4(const port_idx_t *ix, uint64_t k, uint64_t cnt[4]);
void port_extend(const port_idx_t *ix, const port_intv_t *in, port_intv_t out[4], int is_back);
int port_smem1a(const port_idx_t *ix, int len, const uint8_t *q, int x, int min_intv, uint64_t max_intv,
				port_intvv_t *mem, port_intvv_t *scratch0, port_intvv_t *scratch1);
int port_seed_strategy1(const port_idx_t *ix, int len, const uint8_t *q, int x, int min_len, int max_intv, port_intv_t *mem);
uint64_t port_sa(const port_idx_t *ix, uint64_t k);
void port_collect_intv(const port_opt_t *opt, const port_idx_t *ix, int len, const uint8_t *seq, port_intvv_t *out);

/* ---- reference access ---- */
int port_pos2rid(const port_idx_t *ix, int64_t pos_f);
int port_intv2rid(const port_idx_t *ix, int64_t rb, int64_t re);
uint8_t *port_get_seq(const port_idx_t *ix, int64_t beg, int64_t end, int64_t *len);
uint8_t *port_fetch_seq(const port_idx_t *ix, int64_t *beg, int64_t mid, int64_t *end, int *rid);

/* ---- Smith-Waterman family ---- */
int port_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
				 int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
				 int *qle, int *tle, int *gtle, int *gscore, int *max_off);
typedef struct { int score, te, qe, score2, te2, tb, qb; } port_swr_t;         /* B/ksw.h kswr_t */
#define PORT_XBYTE  0x10000
#define PORT_XSTOP  0x20000
#define PORT_XSUBO  0x40000
#define PORT_XSTART 0x80000
port_swr_t port_sw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
						  int o_del, int e_del, int o_ins, int e_ins, int xtra);
int port_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
				 int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar);
uint32_t *port_gen_cigar2(const int8_t mat[25], int o_del, int e_del, int o_ins, int e_ins, int w_, const port_idx_t *ix,
						  int l_query, uint8_t *query, int64_t rb, int64_t re, int *score, int *n_cigar, int *NM);

/* ---- exact-order sort (B/ksort.h:176-226) ---- */
typedef int (*port_lt_fn)(const void *a, const void *b);
void port_introsort(void *base, size_t n, size_t sz, port_lt_fn lt);

/* ---- BWA-MEM proper ---- */
port_regv_t port_align1_core(const port_opt_t *opt, const port_idx_t *ix, int l_seq, uint8_t *seq);
int port_sort_dedup_patch(const port_opt_t *opt, const port_idx_t *ix, uint8_t *query, int n, port_reg_t *a);
int port_matesw(const port_opt_t *opt, const port_idx_t *ix, const port_pes_t pes[4], const port_reg_t *a,
				int l_ms, const uint8_t *ms, port_regv_t *ma);
port_aln_t port_reg2aln(const port_opt_t *opt, const port_idx_t *ix, int l_query, const uint8_t *query, const port_reg_t *ar);

/* counters for the roofline's algorithmic-work figures */
typedef struct {
	uint64_t occ_blocks;      /* 64-byte occ blocks the reference would touch in bwt_extend (1 or 2 per call) */
	uint64_t extend_calls;
	uint64_t sa_calls, lf_steps;
	uint64_t ext_calls, ext_cells;        /* ksw_extend2: sum over rows of (end-beg) */
	uint64_t sw_calls, sw_cells;          /* ksw_align2: qlen * rows, forward + reverse */
	uint64_t glb_calls, glb_cells;        /* ksw_global2: sum over rows of (end-beg) */
} port_counters_t;
extern port_counters_t port_counters;     /* not thread-safe: meaningful for single-thread runs only */

#ifdef __cplusplus
}
#endif
#endif
