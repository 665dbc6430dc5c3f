/* cuda_lariat.h — C ABI of libcuda_lariat.so: B200-native BWA-MEM candidate generation for Lariat.
 *
 * This is the drop-in boundary for the ONE hot path of 10XGenomics/lariat that is rebuilt here:
 * everything Lariat's cgo bridge (go/src/gobwa/gobwa.go, go/src/gobwa/bwa_bridge.h) asks of BWA while it
 * generates candidate alignments for the read pairs of a barcode group.  Plain C, plain pointers and sizes,
 * no torch / CUDA types.  Each entry point names the reference interface it replaces ("G/" =
 * go/src/gobwa/, "B/" = go/src/gobwa/bwa/ in the reference tree).  See INTEGRATION.md for the cgo stub.
 *
 * Threading and ownership (what Lariat's worker pool needs: I/lariat.go:348-376,473,546; G/gobwa.go:107-126):
 *   - An lc_ctx may be used by ANY number of host threads at once.  Calls of up to a few thousand pairs (one barcode group is ~100)
 *     are merged across threads into device batches and cut apart again (lc_set_coalescing); larger calls run one at a time per
 *     context, chunked over two CUDA streams.  lc_clone gives further contexts on the SAME index for callers that want several
 *     large calls in flight: the index (FM-index, dense suffix array, packed reference) is in HBM once per lc_init, never per context.
 *   - A result (*out of lc_align_packed / lc_align_pairs) belongs to the CALLER from the moment the call returns until
 *     lc_free_result: no later call — on this context, a clone, or another thread — touches it.  This is the reference's Arena:
 *     ChainedHit pointers alias the result arrays until the barcode's arena.Free().
 *   - lc_destroy of the last context of an index frees the index; results already handed out stay valid until freed.
 * Supported scoring options: the reference's defaults and anything with LC_MAX_READ_LEN * a + max(penalty) <= 32767 and every
 *   penalty <= 4096 (the DP kernels keep H and E in 16-bit cells); mat must be bwa_fill_scmat(a, b).  lc_init rejects the rest.
 * Errors: functions return 0 on success, a negative LC_E* code otherwise; lc_last_error() gives the text.
 * Nothing here ever calls exit() (the reference does: B/malloc_wrap.c, B/utils.c).
 * There is NO CPU fallback: without a CUDA device lc_init fails with LC_ENODEV.
 */
#ifndef CUDA_LARIAT_H
#define CUDA_LARIAT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_OK        0
#define LC_ENODEV   -1   /* no CUDA device / CUDA runtime error at init */
#define LC_EINDEX   -2   /* index files missing or inconsistent (reference: bwa_idx_load returns NULL, G/gobwa.go:130-135) */
#define LC_EARG     -3   /* bad argument (e.g. read longer than LC_MAX_READ_LEN) */
#define LC_ECUDA    -4   /* CUDA error while running */
#define LC_ENOMEM   -5

#define LC_MAX_READ_LEN 700  /* B/bwamem.c:598-602: the per-seed SW screen is a no-op below ~725 bp; longer reads are out of scope */

/* ---- options: layout-identical to mem_opt_t (B/bwamem.h:25-57); defaults = mem_opt_init (B/bwamem.c:48-84) ---- */
typedef struct {
	int a, b;
	int o_del, e_del;
	int o_ins, e_ins;
	int pen_unpaired;
	int pen_clip5, pen_clip3;
	int w;
	int zdrop;
	uint64_t max_mem_intv;
	int T;
	int flag;
	int min_seed_len;
	int min_chain_weight;
	int max_chain_extend;
	float split_factor;
	int split_width;
	int max_occ;
	int max_chain_gap;
	int n_threads;
	int chunk_size;
	float mask_level;
	float drop_ratio;
	float XA_drop_ratio;
	float mask_level_redun;
	float mapQ_coef_len;
	int mapQ_coef_fac;
	int max_ins;
	int max_matesw;
	int max_XA_hits, max_XA_hits_alt;
	int8_t mat[25];
} lc_opt_t;

/* ---- candidate region: layout-identical to mem_alnreg_t (B/bwamem.h:59-77; sizeof == 88) ---- */
typedef struct {
	int64_t rb, re;
	int qb, qe;
	int rid;
	int score;
	int truesc;
	int sub;
	int alt_sc;
	int csub;
	int sub_n;
	int w;
	int seedcov;
	int secondary;
	int secondary_all;
	int seedlen0;
	int n_comp:30, is_alt:2;
	float frac_rep;
	uint64_t hash;
} lc_alnreg_t;

/* ---- final alignment of one candidate: the fields of mem_aln_t Lariat reads (G/gobwa.go:449-487) ---- */
typedef struct {
	int64_t pos;        /* 0-based leftmost forward-strand position on contig `rid` (mem_aln_t.pos) */
	int32_t rid;
	int32_t is_rev;
	int32_t NM;
	int32_t n_cigar;
	uint32_t cigar_off; /* index of the first op in lc_result.cigars; BAM encoding len<<4|op, MIDSH = 01234 */
	int32_t n_mismatch; /* number of mismatching aligned bases (== NM - gap bases), for GetAlignments (I/lariat.go:1633) */
} lc_aln_t;

/* ---- what inference.GetAlignments computes per candidate AFTER mem_reg2aln (I/lariat.go:1573-1637,1645-1650): the Go loop fetches the
 * reference span with GetSeq (one cgo call per candidate), walks the CIGAR and compares bases.  With lc_set_alnx(ctx, 1) a kernel does
 * the same walk on the device and the result carries the outcome, so the shim needs no further BWA call per candidate.
 * Read and reference are compared byte for byte like the two Go slices ('A' 'C' 'G' 'T' against the read as given: a lower-case or
 * ambiguous read base differs from everything), so the lists are the Go lists. */
typedef struct {
	int64_t pos, aend;           /* Alignment.pos / .aend: contig-relative, swapped (+1) on the reverse strand (I/lariat.go:1645-1650) */
	int32_t matches;             /* M bases minus mismatches (I/lariat.go:1633-1634) */
	int32_t mismatches;          /* EditDistance - indel_length, floored at 0 (I/lariat.go:1633-1637) */
	int32_t indels;              /* number of I and D operations */
	int32_t indel_length;
	int32_t soft_clipped;        /* number of S operations */
	int32_t soft_clipped_length;
	int32_t n_mm;                /* len(mismatchLocs) == len(mismatchReadLocs) */
	uint32_t mm_off;             /* first entry of this candidate in lc_result.mm_ref_pos / mm_read_pos */
} lc_alnx_t;

/* ---- insert-size model: mem_pestat_t (B/bwamem.h:81-85). Lariat hard-codes FR, low=-35, high=500 (G/gobwa.go:229-237) ---- */
typedef struct { int low, high, failed; double avg, std; } lc_pestat_t;

/* ---- contig table: replaces the struct peeking of G/gobwa.go:27-40,136-145,435-447 ---- */
typedef struct {
	int32_t n_seqs;
	int64_t l_pac;
	const char *const *names;
	const int32_t *lens;
	const int64_t *offsets;
} lc_contigs_t;

/* ---- result of one batched call.  Read end e of pair p has index r = 2*p + e.
 * Its candidates are regs[reg_off[r] .. reg_off[r+1]) in exactly the order the reference's
 * GoBwaMemMateSW returns them; aln[i] is mem_reg2aln of regs[i] (what GetAlignments computes per candidate). */
typedef struct {
	int32_t n_reads;              /* 2 * n_pairs */
	int64_t n_regs;
	const int64_t *reg_off;       /* [n_reads + 1] */
	const lc_alnreg_t *regs;      /* [n_regs] */
	const lc_aln_t *aln;          /* [n_regs] */
	const uint32_t *cigars;
	int64_t n_cigar_words;
	/* device-side stage timings of this call, CUDA events on the library's stream (ms) */
	float ms_h2d, ms_seed, ms_sa, ms_chain, ms_extend, ms_dedup, ms_pair, ms_cigar, ms_d2h, ms_total;
	int64_t gpu_launches;         /* kernels of this library launched for the call */
	/* algorithmic work counters, filled when lc_set_counting(ctx, 1) */
	uint64_t cnt_occ_blocks, cnt_extend_calls, cnt_sa_calls, cnt_lf_steps;
	uint64_t cnt_ext_calls, cnt_ext_cells, cnt_sw_calls, cnt_sw_cells, cnt_glb_calls, cnt_glb_cells;
	/* GetAlignments post-processing, when lc_set_alnx(ctx, 1) was called (else null / 0): alnx[i] belongs to regs[i] */
	const lc_alnx_t *alnx;        /* [n_regs] */
	const int32_t *mm_ref_pos;    /* mismatchLocs: contig-relative reference positions, in the order the Go loop appends them */
	const int32_t *mm_read_pos;   /* mismatchReadLocs: positions on the read as given */
	int64_t n_mm;                 /* entries of the two arrays */
} lc_result_t;

typedef struct lc_ctx lc_ctx;

/* mem_opt_init (B/bwamem.c:48-84) — fills *opt with BWA-MEM defaults; what GoBwaAllocSettings returns (G/gobwa.go:149-153) */
void lc_opt_init(lc_opt_t *opt);

/* bwa_idx_load(prefix, BWA_IDX_ALL) (B/bwa.c:262-294; G/gobwa.go:128-147) + upload to HBM of device `device`.
 * opt == NULL means defaults.  Returns NULL on failure (see lc_last_error). */
lc_ctx *lc_init(const char *idx_prefix, int device, const lc_opt_t *opt);
void lc_destroy(lc_ctx *ctx);
const char *lc_last_error(void);

/* contig names / lengths / offsets (GetReferenceContigsInfo, EnumerateContigs: G/gobwa.go:27-40,435-447) */
const lc_contigs_t *lc_contigs(const lc_ctx *ctx);

/* GoBwaMemMateSW (G/gobwa.go:226-337) for every pair of one or more barcode groups, followed by
 * GoBwaSmithWaterman / mem_reg2aln (G/gobwa.go:400-415) on every candidate — one call, a fixed number of launches.
 * seq1[p]/seq2[p]: ASCII bases (ACGTacgt, anything else = N) of length len1[p]/len2[p]; a zero length end is
 * skipped exactly as the Go code does (G/gobwa.go:243,252,287,311).  score_delta: Lariat passes 25 (I/lariat.go:475).
 * *out is owned by the caller until lc_free_result(ctx, *out) (see "Threading and ownership" above). */
int lc_align_pairs(lc_ctx *ctx, int n_pairs, const uint8_t *const *seq1, const int32_t *len1,
                   const uint8_t *const *seq2, const int32_t *len2, int score_delta, lc_result_t **out);

/* Same, from one packed buffer: bases of read r at bases[off[r] .. off[r+1]), r = 2*pair + end (no per-read pointers). */
int lc_align_packed(lc_ctx *ctx, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out);

/* Split form used by bench.py to time the device-resident path: upload (H2D), run (kernels only, inputs already in HBM),
 * fetch (D2H + result assembly).  lc_align_packed == upload + run + fetch. */
int lc_batch_upload(lc_ctx *ctx, int n_pairs, const uint8_t *bases, const int64_t *off);
int lc_batch_run(lc_ctx *ctx, int score_delta);
int lc_batch_fetch(lc_ctx *ctx, lc_result_t **out);
/* (the result of lc_batch_fetch and the arrays of the single-read forms live in the context and are recycled by the next such call) */
/* Arena.Free (G/gobwa.go:119-126): gives a result of lc_align_packed / lc_align_pairs back.  Safe from any thread. */
void lc_free_result(lc_ctx *ctx, lc_result_t *res);
/* Another context on the index of `ctx` (same device): own CUDA stream, pools and lock, no second copy of the index.  Destroy each
 * with lc_destroy.  Replaces nothing in the reference (its bwaidx_t is read-only shared memory, G/gobwa.go:128-147) — this is how the
 * same sharing is expressed for device memory. */
lc_ctx *lc_clone(lc_ctx *ctx);
/* Merging of small calls: calls of at most max_call_pairs pairs (0 = never merge) from all threads and contexts of the index are
 * run as batches of at most max_batch_pairs pairs, up to `lanes` batches in flight; a leader may wait linger_us for company.
 * Defaults 4096 / 65536 / 3 / 0 (env LC_COALESCE_MAX, LC_COALESCE_BATCH, LC_COALESCE_LANES, LC_COALESCE_LINGER_US).  Results do not depend on it. */
int lc_set_coalescing(lc_ctx *ctx, int max_call_pairs, int max_batch_pairs, int lanes, int linger_us);

/* bns_fetch_seq as used by GoBwaReference.GetSeq (G/gobwa.go:50-80): 2-bit codes of contig `rid`, [start,end) on the
 * forward strand, clamped to the contig like the reference; returns the number of codes written to out (<= end-start). */
int64_t lc_fetch_seq(const lc_ctx *ctx, int rid, int64_t start, int64_t end, uint8_t *out);

/* ---- source-compatible single-read forms (slow: one tiny batch each; for tests and for a strictly call-by-call port
 * of the Go bridge).  Returned arrays are malloc'd; the caller frees them with free(), as the Go Arena does (G/gobwa.go:107-126). */
typedef struct { size_t n, m; lc_alnreg_t *a; } lc_alnreg_v;
/* mem_align1_core (B/bwamem.c:1048-1084; G/gobwa.go:181,244,253): seq = 2-bit codes (0..3, >=4 ambiguous) */
int lc_mem_align1_core(lc_ctx *ctx, int l_seq, const uint8_t *seq, lc_alnreg_v *out);
/* mem_matesw (B/bwamem_pair.c:111-180; G/gobwa.go:291,315): returns the number of SW performed, <0 on error */
int lc_mem_matesw(lc_ctx *ctx, const lc_pestat_t pes[4], const lc_alnreg_t *a, int l_ms, const uint8_t *ms, lc_alnreg_v *ma);
/* mem_reg2aln (B/bwamem.c:1086-1156; G/gobwa.go:404): *cigar is malloc'd (n_cigar words) */
int lc_mem_reg2aln(lc_ctx *ctx, int l_query, const uint8_t *query, const lc_alnreg_t *ar, lc_aln_t *aln, uint32_t **cigar);

/* ---- kernel-level entry points for the microbenchmark config (BASELINE.json configs[4]) and parity tests ---- */
/* ksw_extend2 (B/ksw.c:380-479) on n independent jobs; job j: query = q[qoff[j]..qoff[j+1]), target = t[toff[j]..toff[j+1]).
 * out[j*6..] = {score, qle, tle, gtle, gscore, max_off}.  *ms = device time of the kernel, *cells = sum over rows of (end-beg). */
int lc_ksw_extend_batch(lc_ctx *ctx, int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                        const int32_t *h0, int w, int end_bonus, int zdrop, int32_t *out, float *ms, uint64_t *cells);
/* ksw_align2 with mem_matesw's flags (B/bwamem_pair.c:150-151): out[j*7..] = {score, te, qe, score2, te2, tb, qb} */
int lc_ksw_align_batch(lc_ctx *ctx, int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
                       int32_t *out, float *ms, uint64_t *cells);
/* mem_collect_intv (B/bwamem.c:114-162): SMEM + re-seeding + LAST-like seeds of each read, sorted by info.
 * out_off[n+1], out = {x0, x1, x2, info} x total (caller gives capacity in intervals; returns total or <0) */
int64_t lc_collect_intv_batch(lc_ctx *ctx, int n, const uint8_t *codes, const int64_t *off, int64_t *out_off,
                              uint64_t *out, int64_t out_cap, float *ms);
/* bwt_sa (B/bwt.c:86-96) for n BWT rows */
int lc_sa_batch(lc_ctx *ctx, int64_t n, const uint64_t *rows, uint64_t *out, float *ms);

/* `bwa index` = bwa_idx_build (B/bwtindex.c:251-316) for an N-free reference, on the GPU: writes <prefix>.pac/.ann/.amb/.bwt/.sa,
 * byte for byte the files the reference writes (pack: B/bntseq.c:275-328; BWT of forward + reverse complement: B/bwtindex.c:84-121;
 * occ interleave: B/bwtindex.c:151-173; SA sampled every 32 rows: B/bwt.c:62-84; dump formats: B/bwt.c:385-407).
 * codes[i]: the lens[i] bases of contig i as 2-bit codes 0..3, one per byte (ambiguous bases are rejected: the reference draws them
 * from lrand48).  Sized for the 3.1 Gb reference: 6.2e9 suffixes are sorted in lexicographic chunks that fit one GPU's HBM. */
int lc_index_build(const char *prefix, int n_contigs, const char *const *names, const uint8_t *const *codes, const int64_t *lens, int device);

/* ---- the 9-line barcode-sorted FASTQ reader (go/src/fastqreader/reader.go:91-149 ReadOneLine, :190-260 ReadBarcodeSet; zipread.go:63-85),
 * producing barcode sets already packed for lc_align_packed.  gzip or plain text.  Grouping rules are the reference's: a set ends when
 * Barcode10X changes, after 30000 records, after 200 records of a barcode without '-' (not on the whitelist), or after 200 records of a
 * set that continues the previous set's barcode. */
enum { LC_RD_INFO = 0, LC_RD_RGID, LC_RD_BARCODE10X, LC_RD_RAWBARCODE10X, LC_RD_BARCODE10X_QUAL, LC_RD_SAMPLE_INDEX, LC_RD_SAMPLE_INDEX_QUAL,
       LC_RD_TRIM_BASES, LC_RD_TRIM_QUALS, LC_RD_NFIELDS };
typedef struct {
	int32_t n_records;        /* FastQRecords (read pairs) of the set */
	int32_t complete;         /* the bool ReadBarcodeSet returns: the set closes its barcode (or the file) */
	const uint8_t *bases;     /* read 2i = Read1 of record i AFTER trimming `trim` leading bases, read 2i + 1 = Read2: the arrays lc_align_packed takes */
	const uint8_t *quals;     /* same layout */
	const int64_t *off;       /* [2 * n_records + 1] */
	const char *text;         /* field f of record i = text[text_off[i * LC_RD_NFIELDS + f] .. text_off[i * LC_RD_NFIELDS + f + 1]) */
	const int64_t *text_off;  /* [n_records * LC_RD_NFIELDS + 1] */
} lc_readset_t;
typedef struct lc_reader lc_reader;
lc_reader *lc_reader_open(const char *path);
/* 1 = a set (valid until the next call on this reader), 0 = end of file, < 0 = error */
int lc_reader_next_set(lc_reader *rd, int trim, lc_readset_t *out);
void lc_reader_close(lc_reader *rd);

void lc_set_counting(lc_ctx *ctx, int on);
/* GetAlignments post-processing on the device (lc_alnx_t above): off by default; applies to this context and its later clones */
void lc_set_alnx(lc_ctx *ctx, int on);
/* Calls of more than 1.5 x n_pairs pairs are cut into chunks of about n_pairs pairs that alternate between two CUDA streams, so that the
 * H2D / D2H copies of one chunk run under the kernels of the other (the reference's analogue: barcode groups flowing through the
 * goroutine pool of I/lariat.go:348-350 while the BAM thread drains results).  0 = never chunk.  Default 500000 (env LC_CHUNK_PAIRS).
 * Results do not depend on it. */
int lc_set_chunk_pairs(lc_ctx *ctx, int n_pairs);
/* Device-side stopwatch for bench.py: CUDA events recorded on the library's own stream (torch.cuda.Event only sees
 * torch's current stream).  lc_timer_stop synchronises the stream and returns the milliseconds since lc_timer_start. */
int lc_timer_start(lc_ctx *ctx);
int lc_timer_stop(lc_ctx *ctx, float *ms);
/* library build info: "sm_100a", version */
const char *lc_build_info(void);

#ifdef __cplusplus
}
#endif
#endif
