// lc_mem.cuh — per-read / per-pair stage bodies of the batched BWA-MEM pipeline.
//
// One call of a stage body = the work of one read (or pair, or candidate); the CUDA kernels in lc_kernels.cu
// run them for a whole batch of barcode groups.  All inter-stage data lives in HBM pools described by LcBatch;
// variable-size outputs are carved from bump pools (one atomicAdd per read), and a stage that would overrun
// its pool raises `overflow` so the host enlarges the pool and reruns the batch — there is no CPU path.
// Reference lines: "B/" = go/src/gobwa/bwa/, "G/" = go/src/gobwa/.
#pragma once
#include "../../include/cuda_lariat.h"
#include "lc_fmindex.cuh"
#include "lc_sort.cuh"
#include "lc_sw.cuh"

struct LcIntvOut { uint64_t x0, x2; uint32_t qbeg, qend; uint32_t soff, step; };   // 32 B: one seeding interval of a read
struct LcPk { uint32_t a, b, c, m; };     // 16 B packed bi-interval of the seeding state machine (lc_seed.cuh): x0, x1, x2 < 2^34 + two read positions
struct LcSeed { int64_t rbeg; int32_t qbeg, len; };                                 // mem_seed_t (B/bwamem.c:168-172); score == len here
struct LcChainOut { int32_t rid, n; int64_t seed_off; uint32_t w; int32_t is_alt; }; // a kept chain; seeds in cseed[seed_off..+n)
// Stage 4 on reads with very many chains (repeats): the chains of a read are split into UNITS — clusters of chains whose reference
// windows overlap (lc_ext_plan) — and every unit is extended on its own.  Per chain: the next chain of its unit, its read, and
// where the regions it produced went.
struct LcSortKey { uint64_t lo; uint32_t hi, idx; };      // lc_sort_regs_keyed
struct LcChainAux { int32_t next, r, n_reg, unit_seeds; int64_t reg_off; };
struct LcChainRec { int64_t pos; int32_t rid, n, head, tail; uint32_t w; int32_t kept, first, is_alt; };
#define LC_BT_T 5
#define LC_BT_MAXK (2 * LC_BT_T - 1)
// kpos[t] = ch[key[t]].pos (a chain's position never changes once it exists): the descent compares against the node itself instead of
// hopping to the chain record for every probe (the hops were the top lines of chain_heavy_kernel, profiles/r2g)
struct LcBtNode { int32_t n, internal; int32_t key[LC_BT_MAXK]; int32_t child[LC_BT_MAXK + 1]; int64_t kpos[LC_BT_MAXK]; };

enum { LC_OVF_INTV = 1, LC_OVF_SEED = 2, LC_OVF_SCRATCH = 4, LC_OVF_CSEED = 8, LC_OVF_REG = 16, LC_OVF_FINAL = 32,
	   LC_OVF_CIGAR = 64, LC_OVF_Z = 128, LC_OVF_MEMLIST = 256, LC_OVF_CHAIN = 512, LC_OVF_PK = 1024, LC_OVF_MM = 2048 };

// packed 16x2 constants of the extension tick (lc_extsm.cuh), derived from the scoring options once on the host: as kernel
// parameters they are constant-bank operands, not per-tick arithmetic
struct LcExtConst { uint32_t c_a, n_oe_del, n_e_del, n_oe_ins, n_e1, n_e2, n_e4, lut_lo, lut_hi, kmul, dec[4]; };
static inline uint32_t lc_dup16(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }
static inline LcExtConst lc_ext_const(const LcScore &sc)
{
	LcExtConst x;
	const uint32_t qmm = 255u - (uint32_t)(sc.a + sc.b), qn = 255u - (uint32_t)(sc.a + 1);
	x.c_a = lc_dup16(sc.a - 255); x.n_oe_del = lc_dup16(-(sc.o_del + sc.e_del)); x.n_e_del = lc_dup16(-sc.e_del);
	x.n_oe_ins = lc_dup16(-(sc.o_ins + sc.e_ins)); x.n_e1 = lc_dup16(-sc.e_ins); x.n_e2 = lc_dup16(-2 * sc.e_ins); x.n_e4 = lc_dup16(-4 * sc.e_ins);
	x.lut_lo = 255u | qmm << 8 | qmm << 16 | qmm << 24; x.lut_hi = qn * 0x01010101u; x.kmul = (uint32_t)(sc.a + 1);
	for (int k = 0; k < 4; ++k) x.dec[k] = ((uint32_t)(-(2 * k + 1) * sc.e_ins) & 0xffffu) | ((uint32_t)(-(2 * k + 2) * sc.e_ins) & 0xffffu) << 16;
	return x;
}

struct LcBatch {
	// ---- inputs
	int32_t n_reads;
	const uint8_t *seq;            // 2-bit codes, 4 = ambiguous
	const int64_t *off;            // [n_reads + 1]
	LcIndex ix;
	lc_opt_t opt;
	LcScore sc;
	LcExtConst xc;                 // = lc_ext_const(sc)
	int32_t score_delta, pes_low, pes_high, max_anchors;
	LcCounters *cnt;               // null = off
	uint32_t *overflow;            // OR of LC_OVF_* raised by any thread
	unsigned int *work_ctr;        // [4] dynamic work-distribution counters of the warp-per-item kernels (zeroed per batch)
	int32_t lmax;                  // longest read of the batch
	// ---- per-worker scratch (n_workers slots)
	int32_t n_workers;
	LcIntv *w_intv;                // per worker: 3 * (lmax + 1) + memcap
	int32_t memcap;
	int32_t *w_int;                // per worker: w_int_n ints (SW rows)
	int32_t w_int_n;
	unsigned long long *w_u64;     // per worker: w_u64_n (sub-optimal row list of the local SW)
	int32_t w_u64_n;
	// ---- stage 1: seeding
	LcPk *w_pk; int32_t pk_cap;    // per lane of the seeding kernel: the read's seeds so far (pk_cap packed entries)
	int32_t lst_cap;               // per lane: entries of the shared-memory interval list
	int32_t force_wide;            // run the 64-bit interval arithmetic even if the index has < 2^32 symbols (tests)
	int32_t *redo; unsigned int *redo_cur;   // reads the state-machine kernel hands to the one-thread-per-read kernel
	LcPk *pk_pool; unsigned long long pk_pool_cap, *pk_cur;   // unsorted seeds of every read, handed to lc_seed_finish
	int64_t *r_pk_off; int32_t *r_n_pk;      // per read: its slice of pk_pool (r_n_pk < 0: redone by lc_stage_seed)
	LcIntvOut *intv; unsigned long long intv_cap, *intv_cur;
	uint64_t *seed_row;            // BWT rows, overwritten by their SA values in stage 2
	unsigned long long seed_cap, *seed_cur;
	int64_t *r_intv_off; int32_t *r_n_intv; int64_t *r_seed_off; int32_t *r_n_seed;
	// ---- stage 3: chaining (scratch pool in bytes) -> kept chains with compacted seeds
	uint8_t *scratch; unsigned long long scratch_cap, *scratch_cur;
	LcChainOut *chain; unsigned long long chain_cap, *chain_cur;
	LcSeed *cseed; unsigned long long cseed_cap, *cseed_cur;
	int64_t *r_chain_off; int32_t *r_n_chain; float *r_frac_rep; int32_t *r_n_cseed;
	// ---- stage 4: extension -> regions
	lc_alnreg_t *reg; unsigned long long reg_cap, *reg_cur, *reg_total;   // reg_total = number of regions produced
	int64_t *r_reg_off; int32_t *r_n_reg;
	// ---- stage 5: per pair dedup + mate rescue -> final candidate lists
	lc_alnreg_t *fin; unsigned long long fin_cap, *fin_cur;
	int64_t *r_fin_off; int32_t *r_n_fin;
	uint64_t *cand; unsigned long long *cand_cur;   // flat list of (read << 32 | idx) for stage 6
	// ---- stage 6: CIGARs
	lc_aln_t *aln;                 // parallel to fin
	uint32_t *cigar; unsigned long long cigar_cap, *cigar_cur;
	uint8_t *zpool; unsigned long long z_cap, *z_cur;   // traceback matrices too large for the per-worker scratch
	uint8_t *w_z; int32_t w_z_n;                         // per worker: w_z_n bytes of traceback scratch
	uint64_t *cig_jobs; unsigned long long cig_jobs_cap; // LC_CIG_NLIST job lists of cig_jobs_cap entries each (lc_cigar.cuh)
	unsigned int *cig_n;                                 // their lengths
	unsigned int *resc_n;                                // stage 5: number of pairs listed (in `redo`) for the rescue kernel
	uint64_t *wl_in, *wl_out;                            // work lists of the warp-per-item kernels (heavy reads of stage 4, rescue pairs of stage 5):
	                                                     //   entries weight << 32 | id, appended to wl_in; the backend sorts them by weight,
	                                                     //   heaviest first, into wl_out so that the long items do not start last
	// ---- stage 6b (optional): GetAlignments post-processing (lc_post.cuh)
	lc_alnx_t *alnx;                                     // parallel to fin / aln
	const uint8_t *raw;                                  // the reads as given (ASCII), kept beside the 2-bit codes: the Go loop compares BYTES
	                                                     //   (a lower-case read mismatches everywhere); null = compare codes
	int32_t *mm_ref, *mm_read; unsigned long long mm_cap, *mm_cur;   // mismatch position lists, carved per candidate
	int32_t chain_heavy; int32_t *chain_list; unsigned int *chain_heavy_n;   // stage 3: reads with more than chain_heavy seeds are listed for the
	                                                     //   warp-per-read kernel ([0] = list length, [1] = its work counter); 0 = never
	unsigned int *ext_heavy_n;                           // stage 4: [0] = reads with very many chains, listed in wl_in for lc_ext_plan
	int32_t ext_heavy_seeds, ext_heavy_chains;           // a read — and then a unit — is "heavy" above this many chained seeds / chains (0 = never)
	LcChainAux *chain_aux;                               // parallel to `chain`; filled for the chains of listed reads only
	uint32_t *ext_units; unsigned int *ext_unit_n;       // units for the lane kernel (first chain of each, index into `chain`): [0] = count of the
	                                                     //   larger units, filled from the front, [1] = work counter
	unsigned int *ext_unit_back; unsigned long long ext_units_cap;   // count of the small units, filled from the back of the same array: the
	                                                     //   lanes take the larger units first, so that none of them starts last (lc_ext_unit_at)
	uint64_t *ext_hvu_in, *ext_hvu_out;                  // heavy units for the warp-per-unit kernel: weight << 32 | first chain; sorted like wl_in -> wl_out
	unsigned int *ext_hvu_n;                             // [0] = count, [1] = work counter
	unsigned long long ext_hvu_cap;
	unsigned int *prep_heavy_n;                          // stage 5a: [0] = pairs listed in wl_out for the warp-per-pair launch, [1] = its work counter; null = never
	lc_alnreg_t *regtmp;                                 // stage 5a: record scratch of the cooperative dedup, carved with sortk (same offsets)
	LcSortKey *sortk; unsigned long long sortk_cap, *sortk_cur;      // stage 5a: key scratch of the dedup sorts, carved per read
	int32_t force_coop;                                  // tests: the cooperative dedup of stage 5a even with a one-lane executor
	int32_t small;                                       // a batch of few reads: the warp-per-read forms of seeding and extension for every read (latency, not throughput)
};

LC_HD void lc_raise(const LcBatch &b, uint32_t flag)
{
#ifdef __CUDA_ARCH__
	atomicOr(b.overflow, flag);
#else
	*b.overflow |= flag;
#endif
}
// carve `n` elements from a bump pool; returns the offset, or (u64)-1 after raising `flag`
// (warp-aggregated atomic: lc_fmindex.cuh; lc_carve_plain for the lane state machines)
LC_HD unsigned long long lc_carve(const LcBatch &b, unsigned long long *cur, unsigned long long cap, unsigned long long n, uint32_t flag)
{
	unsigned long long o = lc_atomic_add_agg(cur, n);
	if (o + n > cap) { lc_raise(b, flag); return (unsigned long long)-1; }
	return o;
}
LC_HD unsigned long long lc_carve_plain(const LcBatch &b, unsigned long long *cur, unsigned long long cap, unsigned long long n, uint32_t flag)
{
	unsigned long long o = lc_atomic_add(cur, n);
	if (o + n > cap) { lc_raise(b, flag); return (unsigned long long)-1; }
	return o;
}

// ============================================================================================================
// Stage 1 — seeding: mem_collect_intv (B/bwamem.c:114-162) and the occurrence schedule of mem_chain (B/bwamem.c:273-279)
// ============================================================================================================
// Executors of the seeding bodies below: one thread, or the 32 lanes of a warp working on ONE read (lc_kernels.cu).
struct LcSeedExec1 {
	LC_HD int lane_id() const { return 0; }
	LC_HD int n_lanes() const { return 1; }
	LC_HD bool leader() const { return true; }
	LC_HD void sync() const {}
	LC_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
	LC_HD uint64_t shfl(uint64_t v, int) const { return v; }
};

// All SMEMs through query position x (bwt_smem1a with max_intv = 0, B/bwt.c:289-351,353-356), by all lanes of `ex` together.
// mem[] receives them sorted by start; prev/curr are scratch lists of capacity len + 1 (memory every lane sees).  Returns the next x.
// The forward phase is one dependent bwt_extend per base and runs the same in every lane (the leader writes the list).  The backward
// phase extends EVERY interval of the list by the same base: those extensions are independent, one per lane, and what the reference
// then does in list order has a closed form —
//   * a hit is recorded only while the new list is empty, and the first recorded hit blocks the others of this base
//     (its start equals theirs), so only the list's first entry can be recorded, and only if it ended here;
//   * an interval that survives is pushed unless its size equals the size pushed last — and the size pushed last always equals the
//     size of the previous SURVIVOR (pushed: by assignment; not pushed: by the test), so "pushed" is a comparison between neighbours.
// In a high-copy repeat the list holds ~100 intervals for ~100 bases: ~10^4 dependent extensions per read in list order, a few
// hundred rounds this way.
template <class Exec>
LC_HD int lc_smem1_t(const LcIndex &ix, int len, const uint8_t *q, int x, int min_intv, LcIntv *mem, int &n_mem,
					 LcIntv *prev, LcIntv *curr, LcCounters *cnt, const Exec &ex)
{
	LcIntv ik, ok;
	int i, ret, n_prev, n_curr = 0;
	n_mem = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	LcCounters *cnt1 = ex.leader() ? cnt : nullptr;      // the forward extensions are the same call in every lane: counted once
	lc_set_intv(ix, q[x], ik);
	ik.info = x + 1;
	for (i = x + 1; i < len; ++i) {                 // forward: push the interval whenever its size changes
		if (q[i] > 3) { if (ex.leader()) curr[n_curr] = ik; ++n_curr; break; }
		lc_extend(ix, ik, 3 - q[i], 0, ok, cnt1);
		if (ok.x2 != ik.x2) {
			if (ex.leader()) curr[n_curr] = ik;
			++n_curr;
			if (ok.x2 < (uint64_t)min_intv) break;
		}
		ok.info = i + 1;
		ik = ok;
	}
	if (i == len) { if (ex.leader()) curr[n_curr] = ik; ++n_curr; }
	ex.sync();
	for (int a = ex.lane_id(); a < n_curr - 1 - a; a += ex.n_lanes()) { const int c2 = n_curr - 1 - a; LcIntv t = curr[a]; curr[a] = curr[c2]; curr[c2] = t; }
	ex.sync();
	ret = (int)curr[0].info;
	{ LcIntv *t = curr; curr = prev; prev = t; }
	n_prev = n_curr;
	uint64_t last_start = 0;                        // start of the hit recorded last (mem[n_mem - 1].info >> 32)
	for (i = x - 1; i >= -1; --i) {                 // backward: shrink the list base by base
		const int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		n_curr = 0;
		uint64_t carry_x2 = 0;
		bool have_carry = false;
		for (int j0 = 0; j0 < n_prev; j0 += ex.n_lanes()) {
			const int j = j0 + ex.lane_id();
			const bool act = j < n_prev;
			LcIntv p;
			p.x0 = p.x1 = p.x2 = p.info = 0; ok = p;
			if (act) { p = prev[j]; if (c >= 0) lc_extend(ix, p, c, 1, ok, cnt); }
			const bool valid = act && c >= 0 && ok.x2 >= (uint64_t)min_intv;
			if (j == 0 && !valid && (n_mem == 0 || (uint64_t)(i + 1) < last_start)) {      // (lane 0 of the first chunk)
				ik = p; ik.info |= (uint64_t)(i + 1) << 32;
				mem[n_mem] = ik;
			}
			if (j0 == 0) {
				const bool rec = (ex.ballot(j == 0 && !valid && (n_mem == 0 || (uint64_t)(i + 1) < last_start)) & 1u) != 0;
				if (rec) { ++n_mem; last_start = (uint64_t)(i + 1); }
			}
			const uint32_t vmask = ex.ballot(valid);
			const uint32_t below = vmask & ((1u << ex.lane_id()) - 1u);
			const int src = below ? 31 - lc_clz(below) : ex.lane_id();
			const uint64_t nb_x2 = ex.shfl(ok.x2, src);
			const bool push = valid && (below ? ok.x2 != nb_x2 : (!have_carry || ok.x2 != carry_x2));
			const uint32_t pmask = ex.ballot(push);
			if (push) { ok.info = p.info; curr[n_curr + lc_popc(pmask & ((1u << ex.lane_id()) - 1u))] = ok; }
			n_curr += lc_popc(pmask);
			if (vmask) { carry_x2 = ex.shfl(ok.x2, 31 - lc_clz(vmask)); have_carry = true; }
		}
		ex.sync();
		if (n_curr == 0) break;
		{ LcIntv *t = curr; curr = prev; prev = t; }
		n_prev = n_curr;
	}
	for (int a = ex.lane_id(); a < n_mem - 1 - a; a += ex.n_lanes()) { const int c2 = n_mem - 1 - a; LcIntv t = mem[a]; mem[a] = mem[c2]; mem[c2] = t; }
	ex.sync();
	return ret;
}

// mem_collect_intv (B/bwamem.c:114-162): returns the number of intervals written to mem[] (sorted by info), or -1 if memcap is too small.
// prev/curr/m1 are scratch lists of len + 1 entries.
template <class Exec>
LC_HD int lc_collect_intv_t(const LcIndex &ix, const lc_opt_t &opt, int len, const uint8_t *q, LcIntv *prev, LcIntv *curr, LcIntv *m1,
							LcIntv *mem, int memcap, LcCounters *cnt, const Exec &ex)
{
	int n = 0, n1, x;
	const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
	LcCounters *cnt1 = ex.leader() ? cnt : nullptr;
	for (x = 0; x < len;) {                                                // pass 1: all SMEMs
		if (q[x] > 3) { ++x; continue; }
		x = lc_smem1_t(ix, len, q, x, 1, m1, n1, prev, curr, cnt, ex);
		for (int i = 0; i < n1; ++i) {
			int slen = (int)((uint32_t)m1[i].info - (uint32_t)(m1[i].info >> 32));
			if (slen >= opt.min_seed_len) { if (n == memcap) return -1; if (ex.leader()) mem[n] = m1[i]; ++n; }
		}
		ex.sync();
	}
	const int old_n = n;
	for (int k = 0; k < old_n; ++k) {                                      // pass 2: re-seed inside long near-unique SMEMs
		const LcIntv p = mem[k];
		int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
		if (end - start < split_len || p.x2 > (uint64_t)opt.split_width) continue;
		lc_smem1_t(ix, len, q, (start + end) >> 1, (int)p.x2 + 1, m1, n1, prev, curr, cnt, ex);
		for (int i = 0; i < n1; ++i)
			if ((int)((uint32_t)m1[i].info - (uint32_t)(m1[i].info >> 32)) >= opt.min_seed_len) {
				if (n == memcap) return -1;
				if (ex.leader()) mem[n] = m1[i];
				++n;
			}
		ex.sync();
	}
	if (opt.max_mem_intv > 0)                                              // pass 3: LAST-like seeds (one dependent extension per base: every lane the same)
		for (x = 0; x < len;) {
			if (q[x] > 3) { ++x; continue; }
			LcIntv m;
			x = lc_seed_strategy1(ix, len, q, x, opt.min_seed_len, (int)opt.max_mem_intv, m, cnt1);
			if (m.x2 > 0) { if (n == memcap) return -1; if (ex.leader()) mem[n] = m; ++n; }
		}
	ex.sync();
	// elements with equal `info` are identical (same substring => same interval), so any correct sort reproduces B/bwamem.c:161
	if (ex.leader()) lc_introsort(mem, (long)n, [](const LcIntv &a, const LcIntv &c) { return a.info < c.info; });
	ex.sync();
	return n;
}

// Stage 1 for one read, whole: the fallback of the seeding state machine (lc_seed.cuh) for reads whose lists outgrow its slots
template <class Exec>
LC_HD void lc_stage_seed_t(const LcBatch &b, int r, int worker, const Exec &ex)
{
	const lc_opt_t &opt = b.opt;
	const int len = (int)(b.off[r + 1] - b.off[r]);
	const uint8_t *q = b.seq + b.off[r];
	if (ex.leader()) { b.r_n_intv[r] = 0; b.r_n_seed[r] = 0; b.r_intv_off[r] = 0; b.r_seed_off[r] = 0; }
	if (len < opt.min_seed_len) return;                                    // B/bwamem.c:260
	LcIntv *base = b.w_intv + (size_t)worker * (3 * (size_t)(b.lmax + 1) + b.memcap);
	LcIntv *prev = base, *curr = base + (b.lmax + 1), *m1 = base + 2 * (size_t)(b.lmax + 1), *mem = base + 3 * (size_t)(b.lmax + 1);
	const int n = lc_collect_intv_t(b.ix, opt, len, q, prev, curr, m1, mem, b.memcap, b.cnt, ex);
	if (n < 0) { if (ex.leader()) lc_raise(b, LC_OVF_MEMLIST); return; }
	// occurrence schedule: at most max_occ rows per interval, evenly strided (B/bwamem.c:278-279)
	unsigned long long S = 0;
	for (int i = 0; i < n; ++i) {
		uint64_t step = mem[i].x2 > (uint64_t)opt.max_occ ? mem[i].x2 / opt.max_occ : 1;
		uint64_t cnt = (mem[i].x2 + step - 1) / step;
		if (cnt > (uint64_t)opt.max_occ) cnt = opt.max_occ;
		S += cnt;
	}
	unsigned long long io = 0, so = 0;
	if (ex.leader()) {
		io = lc_carve(b, b.intv_cur, b.intv_cap, n, LC_OVF_INTV);
		so = lc_carve(b, b.seed_cur, b.seed_cap, S, LC_OVF_SEED);
	}
	io = ex.shfl(io, 0); so = ex.shfl(so, 0);
	if (io == (unsigned long long)-1 || so == (unsigned long long)-1) return;
	const bool dense = b.ix.sa32 || b.ix.sa64;     // with the dense suffix array resident the rows become positions right here
	unsigned long long s = 0;
	for (int i = 0; i < n; ++i) {
		const uint64_t x0 = mem[i].x0, x2 = mem[i].x2;
		const uint64_t step = x2 > (uint64_t)opt.max_occ ? x2 / opt.max_occ : 1;
		uint64_t cnt = (x2 + step - 1) / step;
		if (cnt > (uint64_t)opt.max_occ) cnt = opt.max_occ;
		if (ex.leader()) {
			LcIntvOut o;
			o.x0 = x0; o.x2 = x2; o.qbeg = (uint32_t)(mem[i].info >> 32); o.qend = (uint32_t)mem[i].info;
			o.soff = (uint32_t)s; o.step = (uint32_t)step;
			b.intv[io + i] = o;
		}
		for (uint64_t t = ex.lane_id(); t < cnt; t += ex.n_lanes()) { const uint64_t row = x0 + t * step; b.seed_row[so + s + t] = dense ? lc_sa(b.ix, row, b.cnt) : row; }
		s += cnt;
	}
	if (ex.leader()) { b.r_intv_off[r] = (int64_t)io; b.r_n_intv[r] = n; b.r_seed_off[r] = (int64_t)so; b.r_n_seed[r] = (int32_t)S; }
}
LC_HDN void lc_stage_seed(const LcBatch &b, int r, int worker) { lc_stage_seed_t(b, r, worker, LcSeedExec1()); }

// ============================================================================================================
// Stage 2 — suffix-array lookups, one BWT row each (bwt_sa inside mem_chain, B/bwamem.c:283)
// ============================================================================================================
LC_HD void lc_stage_sa(const LcBatch &b, unsigned long long g)
{
	b.seed_row[g] = lc_sa(b.ix, b.seed_row[g], b.cnt);
}

// ============================================================================================================
// Stage 3 — chaining and chain filtering: mem_chain (B/bwamem.c:251-315), mem_chain_flt (B/bwamem.c:327-385)
// ============================================================================================================
struct LcChainer {
	LcChainRec *ch; int n_ch;
	LcBtNode *nd; int n_nd, root, n_keys;
	int64_t *s_rbeg; int32_t *s_q, *s_len, *s_next;

	static LC_HD int cmp(int64_t pos, int64_t kp) { return (kp < pos) - (pos < kp); }   // chain_cmp (B/bwamem.c:186)
	LC_HD int new_node(int internal)
	{
		LcBtNode &x = nd[n_nd];
		x.n = 0; x.internal = internal;
		return n_nd++;
	}
	// __kb_getp_aux (B/kbtree.h:117-131)
	LC_HD int find(const LcBtNode &x, int64_t pos, int *r) const
	{
		int lo = 0, hi = x.n;
		if (x.n == 0) return -1;
		while (lo < hi) { int mid = (lo + hi) >> 1; if (cmp(pos, x.kpos[mid]) > 0) lo = mid + 1; else hi = mid; }
		if (lo == x.n) { if (r) *r = 1; return x.n - 1; }
		int rr = cmp(pos, x.kpos[lo]);
		if (r) *r = rr;
		return rr < 0 ? lo - 1 : lo;
	}
	// kb_intervalp, lower neighbour only (B/kbtree.h:151-168)
	LC_HD int lower(int64_t pos) const
	{
		int x = root, low = -1;
		for (;;) {
			const LcBtNode &n = nd[x];
			int r = 0, i = find(n, pos, &r);
			if (i >= 0 && r == 0) return n.key[i];
			if (i >= 0) low = n.key[i];
			if (!n.internal) return low;
			x = n.child[i + 1];
		}
	}
	// __kb_split (B/kbtree.h:176-191)
	LC_HD void split(int xi, int i, int yi)
	{
		int zi = new_node(nd[yi].internal);
		LcBtNode &x = nd[xi], &y = nd[yi], &z = nd[zi];
		z.n = LC_BT_T - 1;
		for (int t = 0; t < LC_BT_T - 1; ++t) { z.key[t] = y.key[t + LC_BT_T]; z.kpos[t] = y.kpos[t + LC_BT_T]; }
		if (y.internal) for (int t = 0; t < LC_BT_T; ++t) z.child[t] = y.child[t + LC_BT_T];
		y.n = LC_BT_T - 1;
		for (int t = x.n; t > i; --t) x.child[t + 1] = x.child[t];
		x.child[i + 1] = zi;
		for (int t = x.n; t > i; --t) { x.key[t] = x.key[t - 1]; x.kpos[t] = x.kpos[t - 1]; }
		x.key[i] = y.key[LC_BT_T - 1]; x.kpos[i] = y.kpos[LC_BT_T - 1];
		++x.n;
	}
	// kb_putp / __kb_putp_aux (B/kbtree.h:192-224)
	LC_HD void put(int key)
	{
		const int64_t pos = ch[key].pos;
		++n_keys;
		if (nd[root].n == LC_BT_MAXK) {
			int s = new_node(1);
			nd[s].child[0] = root;
			split(s, 0, root);
			root = s;
		}
		int xi = root;
		for (;;) {
			LcBtNode &x = nd[xi];
			int i = find(x, pos, 0);
			if (!x.internal) {
				for (int t = x.n - 1; t > i; --t) { x.key[t + 1] = x.key[t]; x.kpos[t + 1] = x.kpos[t]; }
				x.key[i + 1] = key; x.kpos[i + 1] = pos;
				++x.n;
				return;
			}
			++i;
			if (nd[x.child[i]].n == LC_BT_MAXK) {
				split(xi, i, x.child[i]);
				if (cmp(pos, x.kpos[i]) > 0) ++i;
			}
			xi = x.child[i];
		}
	}
	// test_and_merge (B/bwamem.c:190-211) for seed `si` against chain `ci`
	LC_HD bool merge(const lc_opt_t &opt, int64_t l_pac, int ci, int si, int rid)
	{
		LcChainRec &c = ch[ci];
		const int64_t p_r = s_rbeg[si]; const int p_q = s_q[si], p_l = s_len[si];
		const int64_t f_r = s_rbeg[c.head], l_r = s_rbeg[c.tail]; const int f_q = s_q[c.head], l_q = s_q[c.tail], l_l = s_len[c.tail];
		const int64_t qend = l_q + l_l, rend = l_r + l_l;
		if (rid != c.rid) return false;
		if (p_q >= f_q && p_q + p_l <= qend && p_r >= f_r && p_r + p_l <= rend) return true;   // contained: dropped
		if ((l_r < l_pac || f_r < l_pac) && p_r >= l_pac) return false;
		const int64_t x = p_q - l_q, y = p_r - l_r;
		if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - l_l < opt.max_chain_gap && y - l_l < opt.max_chain_gap) {
			s_next[c.tail] = si; c.tail = si; ++c.n;
			return true;
		}
		return false;
	}
	// mem_chain_weight (B/bwamem.c:213-232)
	LC_HD int weight(int ci) const
	{
		int64_t end = 0; int wq = 0, wr = 0;
		for (int s = ch[ci].head; s >= 0; s = s_next[s]) {
			if (s_q[s] >= end) wq += s_len[s];
			else if (s_q[s] + s_len[s] > end) wq += (int)(s_q[s] + s_len[s] - end);
			end = lc_max(end, (int64_t)s_q[s] + s_len[s]);
		}
		end = 0;
		for (int s = ch[ci].head; s >= 0; s = s_next[s]) {
			if (s_rbeg[s] >= end) wr += s_len[s];
			else if (s_rbeg[s] + s_len[s] > end) wr += (int)(s_rbeg[s] + s_len[s] - end);
			end = lc_max(end, s_rbeg[s] + s_len[s]);
		}
		wq = lc_min(wq, wr);
		return wq < 1 << 30 ? wq : (1 << 30) - 1;
	}
};

LC_HD size_t lc_align8(size_t x) { return (x + 7) & ~(size_t)7; }
// what mem_chain_flt's pairwise loop reads of a chain, by position in the weight-sorted order: 16 contiguous bytes instead of the
// head/tail seeds behind four pointers (in repeats the loop is quadratic in the chains of a read and was the tail of chain_kernel)
struct LcChainFlt { int32_t beg, end; int32_t w; int32_t first_alt; };      // first_alt: bit 31 = is_alt, low bits = `first` + 1 (0 = none)
LC_HD size_t lc_chain_scratch_bytes(int S)
{
	return lc_align8(sizeof(LcChainRec) * (size_t)S) + lc_align8(sizeof(LcBtNode) * ((size_t)S / (LC_BT_T - 1) + 2)) +
		   lc_align8(4 * (size_t)S) * 6 /* s_q, s_len, s_next, order, srt, kept_idx */ + lc_align8(sizeof(LcChainFlt) * (size_t)S) + 64;
}

#define LC_CHAIN_SMALL 12
#define LC_CHAIN_SMALL_BYTES 2016       // >= lc_chain_scratch_bytes(LC_CHAIN_SMALL) + 8 (checked in LcChainJob::init)

// Who runs lc_chain_job: one thread (chain_kernel, tests/hostsim) ...
struct LcChainExec1 {
	LC_HD int lane() const { return 0; }
	LC_HD int width() const { return 1; }
	LC_HD bool leader() const { return true; }
	LC_HD void sync() const {}
	LC_HD uint32_t ballot(bool p) const { return p ? 1u : 0u; }
	LC_HD int bcast(int v) const { return v; }
};
// ... or the 32 lanes of a warp (chain_heavy_kernel, lc_kernels.cu: LcChainExecWarp), for reads with thousands of seeds: in repeats
// mem_chain_flt's pairwise loop is quadratic in the chains of a read (none shadows the others: they all weigh the same), the seed lists
// are walked three times, and one thread doing it all was the tail of chain_kernel (profiles/r2b: 255 ms of the C3 step at 1.4 active
// threads per warp).  The lanes split everything that is independent per chain and the kept-chain scan of the filter; what is
// inherently ordered (the B-tree insertions, the introsort) stays with lane 0.  Same statements, same order of side effects.

struct LcChainJob {
	const LcBatch &b; int r, S, NI, n, n_kept;
	LcChainer c;
	int32_t *order, *srt, *kept_idx;
	LcChainFlt *flt;
	const LcIntvOut *iv;
	LC_HD LcChainJob(const LcBatch &b_, int r_) : b(b_), r(r_), S(b_.r_n_seed[r_]), NI(b_.r_n_intv[r_]), n(0), n_kept(0) {}

	// frac_rep: span covered by intervals more frequent than max_occ (B/bwamem.c:265-272,310); false = nothing to chain
	LC_HD bool prologue()
	{
		const int len = (int)(b.off[r + 1] - b.off[r]);
		b.r_n_chain[r] = 0; b.r_chain_off[r] = 0; b.r_frac_rep[r] = 0.f; b.r_n_cseed[r] = 0;
		if (S == 0 && NI == 0) return false;
		iv = b.intv + b.r_intv_off[r];
		int rb = 0, re = 0, l_rep = 0;
		for (int i = 0; i < NI; ++i) {
			int sb = (int)iv[i].qbeg, se = (int)iv[i].qend;
			if (iv[i].x2 <= (uint64_t)b.opt.max_occ) continue;
			if (sb > re) { l_rep += re - rb; rb = sb; re = se; } else re = lc_max(re, se);
		}
		l_rep += re - rb;
		b.r_frac_rep[r] = (float)l_rep / len;
		return S != 0;
	}
	LC_HD void init(uint8_t *p)
	{
		static_assert(sizeof(LcChainRec) * LC_CHAIN_SMALL + 8 + sizeof(LcBtNode) * (LC_CHAIN_SMALL / (LC_BT_T - 1) + 2) + 8 + (4 * LC_CHAIN_SMALL + 8) * 6 +
					  sizeof(LcChainFlt) * LC_CHAIN_SMALL + 8 + 64 + 8 <= LC_CHAIN_SMALL_BYTES, "LC_CHAIN_SMALL_BYTES too small");
		p = (uint8_t *)lc_align8((size_t)p);
		c.ch = (LcChainRec *)p; p += lc_align8(sizeof(LcChainRec) * (size_t)S);
		c.nd = (LcBtNode *)p; p += lc_align8(sizeof(LcBtNode) * ((size_t)S / (LC_BT_T - 1) + 2));
		c.s_q = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		c.s_len = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		c.s_next = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		order = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		srt = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		kept_idx = (int32_t *)p; p += lc_align8(4 * (size_t)S);
		flt = (LcChainFlt *)p;
		c.s_rbeg = (int64_t *)(b.seed_row + b.r_seed_off[r]);
	}
	// what mem_chain needs per seed before it can place it (B/bwamem.c:280-287): query span, and the contig (two binary searches over
	// the contig table) — independent per seed, so all lanes; srt[] holds the rids until sort_by_weight() takes the array over
	template <class EX> LC_HD void prep(const EX &ex)
	{
		for (int i = 0; i < NI; ++i) {
			const int slen = (int)(iv[i].qend - iv[i].qbeg), qb = (int)iv[i].qbeg, s0 = (int)iv[i].soff;
			const int cnt = (i + 1 < NI ? (int)iv[i + 1].soff : S) - s0;
			for (int t = ex.lane(); t < cnt; t += ex.width()) {
				const int si = s0 + t;
				c.s_q[si] = qb; c.s_len[si] = slen; c.s_next[si] = -1;
				const int64_t rbeg = c.s_rbeg[si];
				srt[si] = lc_intv2rid(b.ix, rbeg, rbeg + slen);
			}
		}
		ex.sync();
	}
	// mem_chain (B/bwamem.c:273-307): every seed joins the chain below it in the tree or opens a new one; then the in-order walk
	// (__kb_traverse, B/kbtree.h:336-358).  Ordered work: one thread.
	LC_HD void build()
	{
		const lc_opt_t &opt = b.opt;
		c.n_ch = 0; c.n_nd = 0; c.n_keys = 0;
		c.root = c.new_node(0);
		for (int si = 0; si < S; ++si) {                     // seeds in interval order, occurrences in schedule order
			const int rid = srt[si];
			if (rid < 0) continue;                           // spans contigs or the strand boundary (B/bwamem.c:286-287)
			const int64_t rbeg = c.s_rbeg[si];
			bool add = true;
			if (c.n_keys) {
				int low = c.lower(rbeg);
				if (low >= 0 && c.merge(opt, b.ix.l_pac, low, si, rid)) add = false;
			}
			if (add) {
				LcChainRec &nn = c.ch[c.n_ch];
				nn.pos = rbeg; nn.rid = rid; nn.n = 1; nn.head = nn.tail = si; nn.w = 0; nn.kept = 0; nn.first = -1;
				nn.is_alt = b.ix.anns[rid].is_alt ? 1 : 0;
				c.put(c.n_ch++);
			}
		}
		n = 0;
		// step t of an internal node: even t = descend into child t/2, odd t = emit key t/2
		int st_node[24], st_t[24], top = 0;
		st_node[0] = c.root; st_t[0] = 0;
		while (top >= 0) {
			const LcBtNode &x = c.nd[st_node[top]];
			if (!x.internal) { for (int i = 0; i < x.n; ++i) order[n++] = x.key[i]; --top; continue; }
			const int t = st_t[top]++;
			if (t > 2 * x.n) { --top; continue; }
			if (t & 1) order[n++] = x.key[t >> 1];
			else { ++top; st_node[top] = x.child[t >> 1]; st_t[top] = 0; }
		}
	}
	// ---- mem_chain_flt (B/bwamem.c:327-385)
	template <class EX> LC_HD void weights(const EX &ex)      // independent per chain
	{
		for (int i = ex.lane(); i < n; i += ex.width()) {
			LcChainRec &cc = c.ch[order[i]];
			cc.first = -1; cc.kept = 0;
			cc.w = (uint32_t)c.weight(order[i]);
		}
		ex.sync();
	}
	LC_HD void sort_by_weight()                               // one thread: the permutation of ks_introsort is observable (ties)
	{
		int k = 0;
		for (int i = 0; i < n; ++i) if ((int)c.ch[order[i]].w >= b.opt.min_chain_weight) srt[k++] = order[i];
		n = k;
		if (n == 0) return;
		const LcChainRec *ch = c.ch;
		lc_introsort(srt, (long)n, [ch](int a, int d) { return ch[a].w > ch[d].w; });      // flt_lt (B/bwamem.c:324-325,341)
	}
	// the pairwise loop: chain i (in weight order) against the chains kept so far, in order, until one of them shadows it.  A group of
	// `width` lanes tests `width` kept chains at a time; the first lane that would `break` cuts the scan exactly where the loop does.
	template <class EX> LC_HD void filter(const EX &ex)
	{
		const lc_opt_t &opt = b.opt;
		for (int i = ex.lane(); i < n; i += ex.width()) {      // the loop's view of chain srt[i]
			const LcChainRec &cc = c.ch[srt[i]];
			flt[i].beg = c.s_q[cc.head]; flt[i].end = c.s_q[cc.tail] + c.s_len[cc.tail]; flt[i].w = (int32_t)cc.w;
			flt[i].first_alt = cc.is_alt ? (int32_t)0x80000000u : 0;
		}
		if (ex.leader()) { c.ch[srt[0]].kept = 3; kept_idx[0] = 0; }
		n_kept = 1;
		ex.sync();
		for (int i = 1; i < n; ++i) {
			const LcChainFlt fi = flt[i];
			const int li = fi.end - fi.beg;
			int large_ovlp = 0;
			bool broke = false;
			for (int k0 = 0; k0 < n_kept && !broke; k0 += ex.width()) {
				const int k = k0 + ex.lane();
				bool sig = false, brk = false;
				int j = 0;
				LcChainFlt fj; fj.beg = fj.end = fj.w = fj.first_alt = 0;
				if (k < n_kept) {
					j = kept_idx[k];
					fj = flt[j];
					const int b_max = lc_max(fj.beg, fi.beg), e_min = lc_min(fj.end, fi.end);
					if (e_min > b_max && (fj.first_alt >= 0 || fi.first_alt < 0)) {      // (!a[j].is_alt || a[i].is_alt)
						const int lj = fj.end - fj.beg, min_l = lc_min(li, lj);
						if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
							sig = true;
							brk = fi.w < fj.w * opt.drop_ratio && fj.w - fi.w >= opt.min_seed_len << 1;
						}
					}
				}
				const uint32_t mb = ex.ballot(brk), ms = ex.ballot(sig);
				const int cut = mb ? lc_ctz(mb) : 31;                    // lanes up to and including `cut` are the iterations the loop runs
				const uint32_t upto = cut >= 31 ? 0xffffffffu : ((2u << cut) - 1u);
				if (sig && ex.lane() <= cut && (fj.first_alt & 0x7fffffff) == 0) flt[j].first_alt = fj.first_alt | (i + 1);
				if (ms & upto) large_ovlp = 1;
				broke = mb != 0;
			}
			if (!broke) {
				if (ex.leader()) { kept_idx[n_kept] = i; c.ch[srt[i]].kept = large_ovlp ? 2 : 3; }
				++n_kept;
			}
			ex.sync();
		}
	}
	LC_HD void mark()                                         // B/bwamem.c:369-383
	{
		const lc_opt_t &opt = b.opt;
		for (int i = 0; i < n_kept; ++i) { const int f = flt[kept_idx[i]].first_alt & 0x7fffffff; if (f) c.ch[srt[f - 1]].kept = 1; }
		int i, k;
		for (i = k = 0; i < n; ++i) {
			int kp = c.ch[srt[i]].kept;
			if (kp == 0 || kp == 3) continue;
			if (++k >= opt.max_chain_extend) break;
		}
		for (; i < n; ++i) if (c.ch[srt[i]].kept < 3) c.ch[srt[i]].kept = 0;
	}
	// kept chains in filter order with their seeds compacted: offsets by one thread, the copies by all
	template <class EX> LC_HD void emit(const EX &ex)
	{
		int nk = 0; unsigned long long ns = 0;
		long long co = 0, eo = 0;
		if (ex.leader()) {
			for (int i = 0; i < n; ++i) if (c.ch[srt[i]].kept) { order[nk] = (int32_t)ns; kept_idx[nk] = srt[i]; ++nk; ns += c.ch[srt[i]].n; }      // order / kept_idx are free by now
			const unsigned long long co_ = lc_carve(b, b.chain_cur, b.chain_cap, nk, LC_OVF_CHAIN);
			const unsigned long long eo_ = lc_carve(b, b.cseed_cur, b.cseed_cap, ns, LC_OVF_CSEED);
			if (co_ == (unsigned long long)-1 || eo_ == (unsigned long long)-1) nk = -1;
			else { co = (long long)co_; eo = (long long)eo_; b.r_chain_off[r] = co; b.r_n_chain[r] = nk; b.r_n_cseed[r] = (int32_t)ns; }
		}
		ex.sync();
		nk = ex.bcast(nk);
		if (nk <= 0) return;
		co = ((long long)ex.bcast((int)(co >> 32)) << 32) | (unsigned)ex.bcast((int)co);
		eo = ((long long)ex.bcast((int)(eo >> 32)) << 32) | (unsigned)ex.bcast((int)eo);
		for (int t = ex.lane(); t < nk; t += ex.width()) {
			const LcChainRec &cc = c.ch[kept_idx[t]];
			unsigned long long so = (unsigned long long)eo + (unsigned)order[t];
			LcChainOut o; o.rid = cc.rid; o.n = cc.n; o.seed_off = (int64_t)so; o.w = cc.w; o.is_alt = cc.is_alt;
			b.chain[co + t] = o;
			for (int sd_ = cc.head; sd_ >= 0; sd_ = c.s_next[sd_]) { LcSeed sd; sd.rbeg = c.s_rbeg[sd_]; sd.qbeg = c.s_q[sd_]; sd.len = c.s_len[sd_]; b.cseed[so++] = sd; }
		}
	}
	// the whole of stage 3 for one read
	template <class EX> LC_HD void run(const EX &ex)
	{
		prep(ex);
		if (ex.leader()) build();
		ex.sync();
		n = ex.bcast(n);
		weights(ex);
		if (ex.leader()) sort_by_weight();
		ex.sync();
		n = ex.bcast(n);
		if (n == 0) return;
		filter(ex);
		if (ex.leader()) mark();
		ex.sync();
		emit(ex);
	}
};

// one thread per read (chain_kernel): reads with more than b.chain_heavy seeds are listed for chain_heavy_kernel instead
LC_HDN void lc_stage_chain(const LcBatch &b, int r)
{
	LcChainJob job(b, r);
	if (!job.prologue()) return;
	const int S = job.S;
	if (b.chain_heavy > 0 && S > b.chain_heavy) { b.chain_list[lc_atomic_add_agg(b.chain_heavy_n, 1u)] = r; return; }
	// A read with a handful of seeds (the usual case outside repeats) keeps the chainer's arrays in thread-local memory: the pool is
	// cold global memory, and chaining is a string of dependent accesses to it (ncu: the kernel is latency, not work).
	uint64_t small[LC_CHAIN_SMALL_BYTES / 8];
	uint8_t *p;
	if (S <= LC_CHAIN_SMALL) p = (uint8_t *)small;
	else {
		unsigned long long so = lc_carve(b, b.scratch_cur, b.scratch_cap, lc_chain_scratch_bytes(S), LC_OVF_SCRATCH);
		if (so == (unsigned long long)-1) return;
		p = b.scratch + so;
	}
	job.init(p);
	job.run(LcChainExec1());
}
// one warp (or, in tests/hostsim, one thread) per listed read
template <class EX> LC_HD void lc_stage_chain_heavy_t(const LcBatch &b, int r, const EX &ex)
{
	LcChainJob job(b, r);
	job.iv = b.intv + b.r_intv_off[r];      // prologue() ran in lc_stage_chain
	long long so = 0;
	if (ex.leader()) so = (long long)lc_carve(b, b.scratch_cur, b.scratch_cap, lc_chain_scratch_bytes(job.S), LC_OVF_SCRATCH);
	so = ((long long)ex.bcast((int)(so >> 32)) << 32) | (unsigned)ex.bcast((int)so);
	if (so == -1) return;
	job.init(b.scratch + so);
	job.run(ex);
}

// ============================================================================================================
// Stage 4 — extension: mem_chain2aln (B/bwamem.c:632-786) for every kept chain of a read
// ============================================================================================================
LC_HD int lc_cal_max_gap(const lc_opt_t &opt, int qlen)          // B/bwamem.c:621-628
{
	// (int)((double)x / e + 1.) — for e == 1 (the default) the division is exact, so plain integer arithmetic gives the same value
	int l_del = opt.e_del == 1 ? qlen * opt.a - opt.o_del + 1 : (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
	int l_ins = opt.e_ins == 1 ? qlen * opt.a - opt.o_ins + 1 : (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
	int l = lc_max(l_del, l_ins);
	l = lc_max(l, 1);
	return lc_min(l, opt.w << 1);
}

// Executors: who runs a stage body.  LcExecThread = one thread does everything (scalar SW, per-worker scratch in HBM).
// LcExecWarp (lc_sw_warp.cuh, device only) = the 32 lanes of a warp run the control flow redundantly, lane 0 owns the
// side effects, and the SW jobs are computed cooperatively with shared-memory rows and warp shuffles.
#if defined(LC_STATS) && !defined(__CUDA_ARCH__)
struct LcStats { unsigned long long sw, skipped, onepass, hazard, full_nofix, n_sum, nmax, hv_reads, hv_chains, hv_seeds, hv_scan, hv_ext, hv_regs, hv_maxchains; };
static LcStats lc_stats;
#define LC_STAT(x) (x)
#else
#define LC_STAT(x) ((void)0)
#endif
struct LcExecThread {
	int *H, *E;
	LC_HD bool leader() const { return true; }
	LC_HD void sync() const {}
	LC_HD unsigned long long bcast(unsigned long long v) const { return v; }
	template <class QF, class TF>
	LC_HD LcExtResult extend2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int w, int end_bonus, int zdrop, int h0, LcCounters *cnt) const
	{
		return lc_extend2(qlen, q, tlen, t, sc, w, end_bonus, zdrop, h0, H, E, cnt);
	}
};

// `unit` < 0: all chains of read r, in order (the regions land in one slice of `reg`: r_reg_off / r_n_reg).
// `unit` >= 0: the chains of one unit (index of its first chain in `chain`; the read and the links come from chain_aux): the regions
// of the unit land in a slice of their own, every chain records where its regions are, and r_n_reg[r] is the sum over the read's units
// (r_reg_off[r] = -1 tells lc_pair_prepare to collect them chain by chain).
template <class Exec>
LC_HD void lc_stage_extend_t(const LcBatch &b, int r, long long unit, const Exec &ex)
{
	const lc_opt_t &opt = b.opt;
	const bool whole = unit < 0;
	if (!whole) r = b.chain_aux[unit].r;
	const int l_query = (int)(b.off[r + 1] - b.off[r]);
	const uint8_t *query = b.seq + b.off[r];
	const int NC = b.r_n_chain[r], T = whole ? b.r_n_cseed[r] : b.chain_aux[unit].unit_seeds;
	if (whole) {
		if (ex.leader()) { b.r_n_reg[r] = 0; b.r_reg_off[r] = 0; }
		if (NC == 0) return;
	}
	unsigned long long ro = 0, so = 0;
	if (ex.leader()) {
		ro = lc_carve(b, b.reg_cur, b.reg_cap, T, LC_OVF_REG);
		so = lc_carve(b, b.scratch_cur, b.scratch_cap, 8ull * T + 8, LC_OVF_SCRATCH);
	}
	ro = ex.bcast(ro); so = ex.bcast(so);
	if (ro == (unsigned long long)-1 || so == (unsigned long long)-1) return;
	lc_alnreg_t *av = b.reg + ro;
	uint64_t *srt = (uint64_t *)lc_align8((size_t)(b.scratch + so));
	const int64_t l_pac = b.ix.l_pac;
	int n_av = 0, n_av0 = 0;
	LC_STAT((++lc_stats.hv_reads, lc_stats.hv_chains += NC, lc_stats.hv_seeds += T, lc_stats.hv_maxchains = lc_stats.hv_maxchains > (unsigned long long)NC ? lc_stats.hv_maxchains : NC));
	long long cg = whole ? (long long)b.r_chain_off[r] : unit;
	for (int ci = 0;; ++ci) {
		if (ci > 0 && !whole) {                             // the chain just done: where its regions are; on to the next chain of the unit
			if (ex.leader()) { b.chain_aux[cg].reg_off = (int64_t)(ro + n_av0); b.chain_aux[cg].n_reg = n_av - n_av0; }
			cg = b.chain_aux[cg].next;
		}
		if (whole) { if (ci >= NC) break; cg = (long long)b.r_chain_off[r] + ci; }
		else if (cg < 0) break;
		n_av0 = n_av;
		const LcChainOut &c = b.chain[cg];
		const LcSeed *seeds = b.cseed + c.seed_off;
		if (c.n == 0) continue;
		int64_t r0 = l_pac << 1, r1 = 0;
		for (int i = 0; i < c.n; ++i) {                     // widest window any seed may extend into (B/bwamem.c:642-657)
			const LcSeed &t = seeds[i];
			int64_t bb = t.rbeg - (t.qbeg + lc_cal_max_gap(opt, t.qbeg));
			int tail = l_query - t.qbeg - t.len;
			int64_t ee = t.rbeg + t.len + (tail + lc_cal_max_gap(opt, tail));
			r0 = lc_min(r0, bb); r1 = lc_max(r1, ee);
		}
		r0 = lc_max(r0, (int64_t)0); r1 = lc_min(r1, l_pac << 1);
		if (r0 < l_pac && l_pac < r1) { if (seeds[0].rbeg < l_pac) r1 = l_pac; else r0 = l_pac; }
		lc_fetch_clamp(b.ix, r0, seeds[0].rbeg, r1);        // bns_fetch_seq clamps to the contig (B/bwamem.c:659)
		if (ex.leader()) {
			for (int i = 0; i < c.n; ++i) srt[i] = (uint64_t)(uint32_t)seeds[i].len << 32 | (uint32_t)i;   // score == len (B/bwamem.c:663-665)
			lc_introsort(srt, (long)c.n, [](uint64_t x, uint64_t y) { return x < y; });
		}
		ex.sync();
		for (int k = c.n - 1; k >= 0; --k) {
			const LcSeed s = seeds[(uint32_t)srt[k]];
			int i;
			for (i = 0; i < n_av; ++i) {                    // seed already explained by an earlier region? (B/bwamem.c:671-687)
				const lc_alnreg_t &p = av[i];
				int64_t rd; int qd, w, g;
				LC_STAT(++lc_stats.hv_scan);
				if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
				if (s.len - p.seedlen0 > .1 * l_query) continue;
				qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
				g = lc_cal_max_gap(opt, qd < rd ? qd : (int)rd); w = lc_min(g, p.w);
				if (qd - rd < w && rd - qd < w) break;
				qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
				g = lc_cal_max_gap(opt, qd < rd ? qd : (int)rd); w = lc_min(g, p.w);
				if (qd - rd < w && rd - qd < w) break;
			}
			if (i < n_av) {                                 // contained: extend only if an overlapping off-diagonal seed exists (B/bwamem.c:688-706)
				for (i = k + 1; i < c.n; ++i) {
					if (srt[i] == 0) continue;
					const LcSeed &t = seeds[(uint32_t)srt[i]];
					if (t.len < s.len * .95) continue;
					if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
					if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
				}
				if (i == c.n) { ex.sync(); if (ex.leader()) srt[k] = 0; ex.sync(); continue; }
			}
			lc_alnreg_t a;
			memset(&a, 0, sizeof(a));
			int aw0 = opt.w, aw1 = opt.w;
			a.w = opt.w; a.score = a.truesc = -1; a.rid = c.rid;
			if (s.qbeg) {                                   // left extension on reversed sequences (B/bwamem.c:715-744)
				LcSeqRev qf = { query + s.qbeg - 1 };
				LcRefRev tf = { &b.ix, s.rbeg - 1 };
				LcExtResult x = { 0, 0, 0, 0, 0, 0 };
				for (i = 0; i < 2; ++i) {
					int prev = a.score;
					aw0 = opt.w << i;
					x = ex.extend2(s.qbeg, qf, (int)(s.rbeg - r0), tf, b.sc, aw0, opt.pen_clip5, opt.zdrop, s.len * opt.a, b.cnt);
					a.score = x.score;
					if (a.score == prev || x.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
				}
				if (x.gscore <= 0 || x.gscore <= a.score - opt.pen_clip5) { a.qb = s.qbeg - x.qle; a.rb = s.rbeg - x.tle; a.truesc = a.score; }
				else { a.qb = 0; a.rb = s.rbeg - x.gtle; a.truesc = x.gscore; }
			} else { a.score = a.truesc = s.len * opt.a; a.qb = 0; a.rb = s.rbeg; }
			if (s.qbeg + s.len != l_query) {                // right extension (B/bwamem.c:746-771)
				const int qe = s.qbeg + s.len, sc0 = a.score;
				const int64_t re = s.rbeg + s.len;
				LcSeqFwd qf = { query + qe };
				LcRefFwd tf = { &b.ix, re };
				LcExtResult x = { 0, 0, 0, 0, 0, 0 };
				for (i = 0; i < 2; ++i) {
					int prev = a.score;
					aw1 = opt.w << i;
					x = ex.extend2(l_query - qe, qf, (int)(r1 - re), tf, b.sc, aw1, opt.pen_clip3, opt.zdrop, sc0, b.cnt);
					a.score = x.score;
					if (a.score == prev || x.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
				}
				if (x.gscore <= 0 || x.gscore <= a.score - opt.pen_clip3) { a.qe = qe + x.qle; a.re = re + x.tle; a.truesc += a.score - sc0; }
				else { a.qe = l_query; a.re = re + x.gtle; a.truesc += x.gscore - sc0; }
			} else { a.qe = l_query; a.re = s.rbeg + s.len; }
			a.seedcov = 0;
			for (i = 0; i < c.n; ++i) {
				const LcSeed &t = seeds[i];
				if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
			}
			a.w = lc_max(aw0, aw1);
			a.seedlen0 = s.len;
			a.frac_rep = b.r_frac_rep[r];
			if (ex.leader()) av[n_av] = a;
			++n_av;
			LC_STAT(++lc_stats.hv_regs);
			ex.sync();
		}
	}
	if (ex.leader()) {
		if (whole) { b.r_reg_off[r] = (int64_t)ro; b.r_n_reg[r] = n_av; }
		else if (n_av) lc_atomic_add(b.r_n_reg + r, (int32_t)n_av);
		if (n_av) lc_atomic_add(b.reg_total, (unsigned long long)n_av);
	}
}

LC_HDN void lc_stage_extend(const LcBatch &b, int r, long long unit, int worker)
{
	LcExecThread ex;
	ex.H = b.w_int + (size_t)worker * b.w_int_n; ex.E = ex.H + (b.lmax + 2);
	lc_stage_extend_t(b, r, unit, ex);
}

// ---- units of a read with very many chains (one thread per listed read)
// mem_chain2aln lets a chain see the regions of the chains before it only through the containment test (B/bwamem.c:671-687): a seed is
// skipped when it lies inside an earlier region.  A region never leaves the reference window of its chain (the extensions run on the
// window's sequence, B/bwamem.c:642-659), and neither does a seed, so two chains whose windows share no base cannot influence each
// other: the connected groups of overlapping windows — the UNITS — can be extended independently, each one running its chains in the
// read's chain order against the regions of that unit alone.  In repeats a read has hundreds of chains at hundreds of loci, i.e.
// hundreds of one-chain units: work for many lanes instead of one warp.  The regions come back chain by chain (LcChainAux), so the
// read's list is rebuilt in the reference's order whatever order the units finished in.
// Windows are taken before the strand / contig clamps (a superset: merging two units is always safe).
#define LC_EXT_UNIT_FRONT 4      // units with at least this many seeds are dealt first
LC_HD int lc_ext_n_units(const LcBatch &b) { return (int)(b.ext_unit_n[0] + b.ext_unit_back[0]); }
LC_HD uint32_t lc_ext_unit_at(const LcBatch &b, int i)
{
	const int nf = (int)b.ext_unit_n[0];
	return i < nf ? b.ext_units[i] : b.ext_units[b.ext_units_cap - 1 - (unsigned long long)(i - nf)];
}
LC_HD unsigned long long lc_ext_plan_scratch_bytes(int NC) { return 40ull * NC + 16; }
LC_HD void lc_ext_plan(const LcBatch &b, int r)
{
	const lc_opt_t &opt = b.opt;
	const int l_query = (int)(b.off[r + 1] - b.off[r]), NC = b.r_n_chain[r];
	const long long c0 = (long long)b.r_chain_off[r];
	b.r_reg_off[r] = -1; b.r_n_reg[r] = 0;
	const unsigned long long so = lc_carve(b, b.scratch_cur, b.scratch_cap, lc_ext_plan_scratch_bytes(NC), LC_OVF_SCRATCH);
	if (so == (unsigned long long)-1) return;
	uint64_t *key = (uint64_t *)lc_align8((size_t)(b.scratch + so));      // window start << 24 | chain
	int64_t *wend = (int64_t *)(key + NC);                                // window end
	int32_t *unit_of = (int32_t *)(wend + NC), *first = unit_of + NC, *last = first + NC, *tot = last + NC, *cnt = tot + NC;
	int n_used = 0;
	for (int ci = 0; ci < NC; ++ci) {
		const LcChainOut &c = b.chain[c0 + ci];
		LcChainAux &ax = b.chain_aux[c0 + ci];
		ax.next = -1; ax.r = r; ax.n_reg = 0; ax.unit_seeds = 0; ax.reg_off = 0;
		if (c.n == 0) continue;
		const LcSeed *seeds = b.cseed + c.seed_off;
		int64_t x0 = b.ix.l_pac << 1, x1 = 0;
		for (int t = 0; t < c.n; ++t) {
			const LcSeed &sd = seeds[t];
			const int tail = l_query - sd.qbeg - sd.len;
			x0 = lc_min(x0, sd.rbeg - (sd.qbeg + lc_cal_max_gap(opt, sd.qbeg)));
			x1 = lc_max(x1, sd.rbeg + sd.len + (tail + lc_cal_max_gap(opt, tail)));
		}
		key[n_used] = (uint64_t)(x0 + (1ll << 20)) << 24 | (uint32_t)ci;      // windows start above -2^20 (read length + gap allowance)
		wend[ci] = x1;
		++n_used;
	}
	lc_introsort(key, (long)n_used, [](uint64_t x, uint64_t y) { return x < y; });
	int n_units = 0;
	int64_t cur_end = -(1ll << 40);
	for (int t = 0; t < n_used; ++t) {
		const int ci = (int)(key[t] & 0xffffffu);
		const int64_t x0 = (int64_t)(key[t] >> 24) - (1ll << 20);
		if (x0 >= cur_end) { first[n_units] = last[n_units] = -1; tot[n_units] = 0; cnt[n_units] = 0; ++n_units; cur_end = wend[ci]; }
		else cur_end = lc_max(cur_end, wend[ci]);
		unit_of[ci] = n_units - 1;
	}
	for (int ci = 0; ci < NC; ++ci) {                                    // chains of a unit linked in the read's chain order
		if (b.chain[c0 + ci].n == 0) continue;
		const int u = unit_of[ci];
		if (last[u] < 0) first[u] = ci; else b.chain_aux[c0 + last[u]].next = (int32_t)(c0 + ci);
		last[u] = ci; tot[u] += b.chain[c0 + ci].n; ++cnt[u];
	}
	for (int u = 0; u < n_units; ++u) {
		const long long fc = c0 + first[u];
		b.chain_aux[fc].unit_seeds = tot[u];
		if ((b.ext_heavy_seeds > 0 && tot[u] > b.ext_heavy_seeds) || (b.ext_heavy_chains > 0 && cnt[u] > b.ext_heavy_chains)) {
			const unsigned int at = lc_atomic_add(b.ext_hvu_n, 1u);
			if (at < b.ext_hvu_cap) b.ext_hvu_in[at] = (uint64_t)(uint32_t)lc_min(tot[u], 0xffff) << 32 | (uint32_t)fc;
			else lc_raise(b, LC_OVF_CHAIN);
		} else if (tot[u] >= LC_EXT_UNIT_FRONT) b.ext_units[lc_atomic_add(b.ext_unit_n, 1u)] = (uint32_t)fc;
		else b.ext_units[b.ext_units_cap - 1 - lc_atomic_add(b.ext_unit_back, 1u)] = (uint32_t)fc;
	}
}

// ============================================================================================================
// Stage 5 — per pair: mem_sort_dedup_patch (B/bwamem.c:437-489) on each end, then the two mate-rescue phases of
// GoBwaMemMateSW (G/gobwa.go:264-324) with mem_matesw (B/bwamem_pair.c:111-180)
// ============================================================================================================
struct LcSwScratch { int *H, *E, *Hmax; unsigned long long *b; uint8_t *z_unused; };

// bwa_gen_cigar2, score only (B/bwa.c:121-207) — used by mem_patch_reg
LC_HDN int lc_global_score(const LcBatch &b, int w_, int l_query, const uint8_t *query, int64_t rb, int64_t re, int *H, int *E, bool &ok)
{
	const int64_t l_pac = b.ix.l_pac;
	ok = false;
	if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
	if (re > l_pac << 1 || rb < 0) return 0;               // bns_get_seq would return a shorter sequence (B/bwa.c:134)
	const int rlen = (int)(re - rb);
	ok = true;
	const bool rev = rb >= l_pac;
	if (l_query == rlen && w_ == 0) {
		int s = 0;
		for (int i = 0; i < l_query; ++i) s += b.sc.sub(lc_ref_base(b.ix, rb + i), query[i]);   // order-independent sum
		return s;
	}
	int mi = (int)((double)(((l_query + 1) >> 1) * b.sc.a - b.sc.o_ins) / b.sc.e_ins + 1.);
	int md = (int)((double)(((l_query + 1) >> 1) * b.sc.a - b.sc.o_del) / b.sc.e_del + 1.);
	int mg = lc_max(lc_max(mi, md), 1), dl = rlen > l_query ? rlen - l_query : l_query - rlen;
	int w = (mg + dl + 1) >> 1;
	w = lc_min(w, w_);
	w = lc_max(w, dl + 3);
	if (rev) {   // both sequences reversed so gaps are left-aligned on the forward strand (B/bwa.c:135-140)
		LcSeqRev qf = { query + l_query - 1 }; LcRefRev tf = { &b.ix, re - 1 };
		return lc_global2(l_query, qf, rlen, tf, b.sc, w, H, E, (uint8_t *)0, (uint32_t *)0, 0, (int *)0, b.cnt);
	}
	LcSeqFwd qf = { query }; LcRefFwd tf = { &b.ix, rb };
	return lc_global2(l_query, qf, rlen, tf, b.sc, w, H, E, (uint8_t *)0, (uint32_t *)0, 0, (int *)0, b.cnt);
}

// mem_patch_reg (B/bwamem.c:406-435)
LC_HDN int lc_patch_reg(const LcBatch &b, const uint8_t *query, const lc_alnreg_t &a, const lc_alnreg_t &p, int *H, int *E, int &w_out)
{
	const lc_opt_t &opt = b.opt;
	if (a.rb < b.ix.l_pac && p.rb >= b.ix.l_pac) return 0;
	if (a.qb >= p.qb || a.qe >= p.qe || a.re >= p.re) return 0;
	int w = (int)((a.re - p.rb) - (a.qe - p.qb));
	w = w > 0 ? w : -w;
	double r = (double)(a.re - p.rb) / (p.re - a.rb) - (double)(a.qe - p.qb) / (p.qe - a.qb);
	r = r > 0. ? r : -r;
	if (a.re < p.rb || a.qe < p.qb) { if (w > opt.w << 1 || r >= 0.05f) return 0; }
	else if (w > opt.w << 2 || r >= 0.05f * 2) return 0;
	w += a.w + p.w;
	w = lc_min(w, opt.w << 2);
	bool ok;
	int score = lc_global_score(b, w, p.qe - a.qb, query + a.qb, a.rb, p.re, H, E, ok);
	// when bwa_gen_cigar2 rejects the span it leaves `score` unset (B/bwa.c:132,134 / B/bwamem.c:408,428); that cannot
	// happen for two colinear same-strand regions inside one contig, which is all that reaches this point
	int q_s = (int)((double)(p.qe - a.qb) / ((p.qe - p.qb) + (a.qe - a.qb)) * (p.score + a.score) + .499);
	int r_s = (int)((double)(p.re - a.rb) / ((p.re - p.rb) + (a.re - a.rb)) * (p.score + a.score) + .499);
	if ((double)score / lc_max(q_s, r_s) < 0.90f) return 0;
	w_out = w;
	return score;
}

// mem_sort_dedup_patch (B/bwamem.c:437-489); query == null disables patching (the call from mem_matesw, B/bwamem_pair.c:175)
// The reference's introsort over regions, run on 4-byte handles when `h` (h_cap ints of scratch) can hold them: the algorithm sees the
// same comparisons in the same order, so the permutation — ties included — is the one ks_introsort leaves on the structs themselves;
// it is then applied in place, one 88-byte move per region instead of one or two per swap.  (In repeats the rescue stage re-sorts
// lists of hundreds of regions after every insertion, by a single lane: profiles/r01y, 29 % of matesw_kernel's samples.)
template <class Lt> LC_HD void lc_sort_regs(lc_alnreg_t *a, int n, Lt lt, int *h, int h_cap)
{
	if (n < 12 || !h || n > h_cap) { lc_introsort(a, (long)n, lt); return; }
	for (int i = 0; i < n; ++i) h[i] = i;
	lc_introsort(h, (long)n, [a, lt](int x, int y) { return lt(a[x], a[y]); });
	for (int i = 0; i < n; ++i) {                          // position j receives the region that was at h[j]; visited positions are marked h[j] = j
		if (h[i] == i) continue;
		const lc_alnreg_t first = a[i];
		int j = i;
		for (;;) {
			const int k = h[j];
			h[j] = j;
			if (k == i) { a[j] = first; break; }
			a[j] = a[k];
			j = k;
		}
	}
}
// The same two sorts on 16-byte {key, index} entries in a compact scratch array: a comparison reads two adjacent-ish entries instead of
// hopping through a handle into 88-byte records scattered over the list (hundreds of regions per read in repeats: the sorts of
// mem_align1_core's dedup were most of pair_prepare_kernel's tail).  which = 1: by `re` (mem_ars2); which = 2: by score (descending), rb, qb
// (mem_ars).  The keys order exactly as the comparators do, so introsort sees the same outcomes and leaves the same permutation.
LC_HD void lc_sort_regs_keyed(lc_alnreg_t *a, int n, int which, LcSortKey *ks)
{
	for (int i = 0; i < n; ++i) {
		LcSortKey k;
		if (which == 1) { k.lo = (uint64_t)a[i].re; k.hi = 0; }
		else { k.lo = (uint64_t)a[i].rb << 16 | (uint64_t)(uint16_t)a[i].qb; k.hi = (uint32_t)(0x7fffffffll - (int64_t)a[i].score); }
		k.idx = (uint32_t)i;
		ks[i] = k;
	}
	lc_introsort(ks, (long)n, [](const LcSortKey &x, const LcSortKey &y) { return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo); });
	for (int i = 0; i < n; ++i) {                          // the permutation, applied in place (as above)
		if ((int)ks[i].idx == i) continue;
		const lc_alnreg_t first = a[i];
		int j = i;
		for (;;) {
			const int k = (int)ks[j].idx;
			ks[j].idx = (uint32_t)j;
			if (k == i) { a[j] = first; break; }
			a[j] = a[k];
			j = k;
		}
	}
}
// h/h_cap: scratch for lc_sort_regs (may overlap H and E: the sorts and the patching never run at the same time)
// *clean (optional) := the call excluded nothing and merged nothing (and saw at least two regions): its output is then a fixed point of the
// query-less form — sorted by the strict order (score, rb, qb), every n_comp == 1, no redundant pair (see lc_matesw_t)
// the middle of mem_sort_dedup_patch (B/bwamem.c:441-474): n_comp, the redundancy / patch loop over the list sorted by `re`, compaction
LC_HDN int lc_dedup_core(const LcBatch &b, const uint8_t *query, int n, lc_alnreg_t *a, int *H, int *E, bool &merged)
{
	const lc_opt_t &opt = b.opt;
	int m, i, j;
	for (i = 0; i < n; ++i) a[i].n_comp = 1;
	for (i = 1; i < n; ++i) {
		lc_alnreg_t &p = a[i];
		if (p.rid != a[i - 1].rid || p.rb >= a[i - 1].re + opt.max_chain_gap) continue;
		for (j = i - 1; j >= 0 && p.rid == a[j].rid && p.rb < a[j].re + opt.max_chain_gap; --j) {
			lc_alnreg_t &q = a[j];
			int64_t orr, oq, mr, mq;
			int score, w;
			if (q.qe == q.qb) continue;
			orr = q.re - p.rb;
			oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
			mr = lc_min(q.re - q.rb, p.re - p.rb);
			mq = lc_min(q.qe - q.qb, p.qe - p.qb);
			if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
				if (p.score < q.score) { p.qe = p.qb; break; }
				else q.qe = q.qb;
			} else if (query && q.rb < p.rb && (score = lc_patch_reg(b, query, q, p, H, E, w)) > 0) {
				p.n_comp += q.n_comp + 1;
				p.seedcov = lc_max(p.seedcov, q.seedcov);
				p.sub = lc_max(p.sub, q.sub);
				p.csub = lc_max(p.csub, q.csub);
				p.qb = q.qb; p.rb = q.rb;
				p.truesc = p.score = score;
				p.w = w;
				q.qb = q.qe;
				merged = true;
			}
		}
	}
	for (i = 0, m = 0; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	return m;
}
// its tail (B/bwamem.c:476-487), on the list sorted by (score, rb, qb): identical hits marked and dropped
LC_HD int lc_dedup_identical(lc_alnreg_t *a, int n)
{
	int i, m;
	for (i = 1; i < n; ++i)
		if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
	for (i = 1, m = 1; i < n; ++i) if (a[i].qe > a[i].qb) { if (m != i) a[m] = a[i]; ++m; }
	return n < 1 ? n : m;
}

// ks (optional): n entries of scratch for the keyed sorts (regions with 0 <= qb < 2^16 and 0 <= rb < 2^48: every real one)
LC_HDN int lc_sort_dedup_patch(const LcBatch &b, const uint8_t *query, int n, lc_alnreg_t *a, int *H, int *E, int *h = nullptr, int h_cap = 0, bool *clean = nullptr,
							   LcSortKey *ks = nullptr)
{
	if (clean) *clean = false;
	if (n <= 1) return n;
	const int n_in = n;
	bool merged = false;
	if (ks && n >= 12) lc_sort_regs_keyed(a, n, 1, ks);
	else lc_sort_regs(a, n, [](const lc_alnreg_t &x, const lc_alnreg_t &y) { return x.re < y.re; }, h, h_cap);
	n = lc_dedup_core(b, query, n, a, H, E, merged);
	if (ks && n >= 12) lc_sort_regs_keyed(a, n, 2, ks);
	else lc_sort_regs(a, n, [](const lc_alnreg_t &x, const lc_alnreg_t &y) {
		return x.score > y.score || (x.score == y.score && (x.rb < y.rb || (x.rb == y.rb && x.qb < y.qb))); }, h, h_cap);
	const int m = lc_dedup_identical(a, n);
	if (clean) *clean = !merged && m == n_in;
	return m;
}

// The same function run by the lanes of `ex` together, for long lists: the keys are filled and the two permutations applied by all
// lanes (out of place, through `tmp`: n records of scratch), the introsorts and the dedup loop stay with the leader.  a[] in, a[] out.
template <class Exec>
LC_HD int lc_sort_dedup_patch_coop(const LcBatch &b, const uint8_t *query, int n, lc_alnreg_t *a, lc_alnreg_t *tmp, int *H, int *E, bool &clean,
								   LcSortKey *ks, const Exec &ex)
{
	clean = false;
	if (n <= 1) return n;
	const int n_in = n;
	for (int i = ex.lane_id(); i < n; i += ex.n_lanes()) { LcSortKey k; k.lo = (uint64_t)a[i].re; k.hi = 0; k.idx = (uint32_t)i; ks[i] = k; }
	ex.sync();
	if (ex.leader()) lc_introsort(ks, (long)n, [](const LcSortKey &x, const LcSortKey &y) { return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo); });
	ex.sync();
	for (int i = ex.lane_id(); i < n; i += ex.n_lanes()) tmp[i] = a[ks[i].idx];
	ex.sync();
	unsigned long long mg = 0;
	int m = n;
	if (ex.leader()) { bool merged = false; m = lc_dedup_core(b, query, n, tmp, H, E, merged); mg = merged ? 1 : 0; }
	ex.sync();
	n = (int)ex.bcast((unsigned long long)(unsigned)m);
	const bool merged = ex.bcast(mg) != 0;
	for (int i = ex.lane_id(); i < n; i += ex.n_lanes()) {
		LcSortKey k;
		k.lo = (uint64_t)tmp[i].rb << 16 | (uint64_t)(uint16_t)tmp[i].qb; k.hi = (uint32_t)(0x7fffffffll - (int64_t)tmp[i].score); k.idx = (uint32_t)i;
		ks[i] = k;
	}
	ex.sync();
	if (ex.leader()) lc_introsort(ks, (long)n, [](const LcSortKey &x, const LcSortKey &y) { return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo); });
	ex.sync();
	for (int i = ex.lane_id(); i < n; i += ex.n_lanes()) a[i] = tmp[ks[i].idx];
	ex.sync();
	m = n;
	if (ex.leader()) m = lc_dedup_identical(a, n);
	ex.sync();
	m = (int)ex.bcast((unsigned long long)(unsigned)m);
	clean = !merged && m == n_in;
	return m;
}

LC_HD int lc_infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t &dist)   // mem_infer_dir (B/bwamem_pair.c:23-30)
{
	int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
	int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
	dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

// The part of mem_matesw (B/bwamem_pair.c:118-147) that decides whether a Smith-Waterman is run for anchor `a`, with Lariat's
// insert model: only orientation FR (index 1) is enabled (G/gobwa.go:229-237), so at most one SW is run per call.
// Returns false when the call changes nothing; else [rb, re) is the reference window.
template <class Exec>
LC_HD bool lc_matesw_wants_sw_t(const LcBatch &b, const lc_alnreg_t &a, int l_ms, const lc_alnreg_t *ma, int n_ma, int64_t &rb, int64_t &re, const Exec &ex)
{
	const int64_t l_pac = b.ix.l_pac;
	const int low = b.pes_low, high = b.pes_high;
	bool found = false;
	for (int i = ex.lane_id(); i < n_ma && !found; i += ex.n_lanes()) {      // an existing mate hit in range satisfies the orientation (:118-124)
		int64_t dist;
		int r = lc_infer_dir(l_pac, a.rb, ma[i].rb, dist);
		if (r == 1 && dist >= low && dist <= high) found = true;
	}
	if (ex.any(found)) return false;
	// r = 1: is_rev = 1 (mate is reverse-complemented), is_larger = 1 (:130-143)
	rb = a.rb + low - l_ms; re = a.rb + high;
	if (rb < 0) rb = 0;
	if (re > l_pac << 1) re = l_pac << 1;
	int rid = -1;
	if (rb < re) rid = lc_fetch_clamp(b.ix, rb, (rb + re) >> 1, re);
	// (when rb >= re the reference compares against an uninitialised rid, :147; treated as "no window")
	return a.rid == rid && re - rb >= b.opt.min_seed_len;
}
struct LcExecOne {      // the one-lane executor of the list scans (thread-per-pair callers)
	LC_HD int lane_id() const { return 0; }
	LC_HD int n_lanes() const { return 1; }
	LC_HD bool any(bool v) const { return v; }
};
LC_HD bool lc_matesw_wants_sw(const LcBatch &b, const lc_alnreg_t &a, int l_ms, const lc_alnreg_t *ma, int n_ma, int64_t &rb, int64_t &re)
{
	return lc_matesw_wants_sw_t(b, a, l_ms, ma, n_ma, rb, re, LcExecOne());
}

// one region against another in the redundancy test of mem_sort_dedup_patch (B/bwamem.c:451-456): p follows q in the order by `re`
LC_HD bool lc_reg_redundant(const lc_opt_t &opt, const lc_alnreg_t &p, const lc_alnreg_t &q)
{
	const int64_t orr = q.re - p.rb;
	const int64_t oq = q.qb < p.qb ? q.qe - p.qb : p.qe - q.qb;
	const int64_t mr = lc_min(q.re - q.rb, p.re - p.rb);
	const int64_t mq = lc_min((int64_t)(q.qe - q.qb), (int64_t)(p.qe - p.qb));
	return orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq;
}

// a[at .. n) moves up by one slot, all lanes of the executor together (top chunk first: a chunk is read, then written)
template <class Exec>
LC_HD void lc_regs_shift_up(lc_alnreg_t *a, int at, int n, const Exec &ex)
{
	const int W = ex.n_lanes();
	for (int hi = n; hi > at; hi -= W) {
		const int i = hi - 1 - ex.lane_id();
		lc_alnreg_t t;
		if (i >= at) t = a[i];
		ex.sync();
		if (i >= at) a[i + 1] = t;
		ex.sync();
	}
}

// mem_matesw (B/bwamem_pair.c:111-180)
// `fixed` (state of the LIST, kept by the caller across calls, the same in every lane): the list is a fixed point of the query-less
// mem_sort_dedup_patch — sorted by the strict order (score, rb, qb) with distinct keys, every n_comp == 1, at least two regions, and
// no pair of regions that the redundancy test would remove.  The reference runs mem_sort_dedup_patch after EVERY Smith-Waterman,
// accepted or not (B/bwamem_pair.c:175); in repeats the lists hold hundreds of regions and a pair makes up to a hundred calls, and those
// two introsorts by a single lane were most of this stage (profiles/r01y, r2c).  On a fixed list they are not needed:
//   * nothing inserted: the call sees the same array, makes the same comparisons and leaves the same array — skipped;
//   * region x inserted: every test between two old regions comes out as it did before (the tests are pairwise; which pairs meet
//     depends only on `re` and rid, and for two regions with equal `re` the test is symmetric, so the place introsort gives to ties
//     does not matter).  What is new are the tests between x and each old region y.  If none of them fires, and x's key (score, rb, qb)
//     differs from every old key, the call's result is the old list with x at its place in the key order: one pass over the list by
//     all lanes (tests, key ties, rank of x), one cooperative shift.  The list is again a fixed point.
//     If any test fires (or a key ties) the reference's statements are replayed as they stand, by the leader lane.
template <class Exec>
LC_HD int lc_matesw_t(const LcBatch &b, const lc_alnreg_t &a, int l_ms, const uint8_t *ms, lc_alnreg_t *ma, int &n_ma, const Exec &ex, bool &fixed)
{
	const lc_opt_t &opt = b.opt;
	const int64_t l_pac = b.ix.l_pac;
	int64_t rb, re;
	if (!lc_matesw_wants_sw_t(b, a, l_ms, ma, n_ma, rb, re, ex)) return 0;
	const int xtra = LC_XSUBO | LC_XSTART | (l_ms * opt.a < 250 ? LC_XBYTE : 0) | (opt.min_seed_len * opt.a);
	LcSeqRevComp qf = { ms, l_ms };
	LcRefFwd tf = { &b.ix, rb };
	const int anchor_rid = a.rid, anchor_is_alt = a.is_alt;      // `a` may alias the list that is about to be reshuffled
	LcSwResult aln = ex.sw_align2(l_ms, qf, (int)(re - rb), tf, b.sc, xtra, b.cnt);
	ex.sync();
	aln.score = (int)ex.bcast((unsigned long long)(unsigned)aln.score); aln.score2 = (int)ex.bcast((unsigned long long)(unsigned)aln.score2);
	aln.qb = (int)ex.bcast((unsigned long long)(unsigned)aln.qb); aln.qe = (int)ex.bcast((unsigned long long)(unsigned)aln.qe);
	aln.tb = (int)ex.bcast((unsigned long long)(unsigned)aln.tb); aln.te = (int)ex.bcast((unsigned long long)(unsigned)aln.te);
	int n = n_ma;
	const bool inserted = aln.score >= opt.min_seed_len && aln.qb >= 0;
	lc_alnreg_t nb;
	if (inserted) {
		memset(&nb, 0, sizeof(nb));
		nb.rid = anchor_rid;
		nb.is_alt = anchor_is_alt;
		nb.qb = l_ms - (aln.qe + 1);
		nb.qe = l_ms - aln.qb;
		nb.rb = (l_pac << 1) - (rb + aln.te + 1);
		nb.re = (l_pac << 1) - (rb + aln.tb);
		nb.score = aln.score;
		nb.csub = aln.score2;
		nb.secondary = -1;
		nb.seedcov = (int)((nb.re - nb.rb < nb.qe - nb.qb ? nb.re - nb.rb : nb.qe - nb.qb) >> 1);
	}
	bool done = !inserted && fixed;                        // the skipped call
	LC_STAT((++lc_stats.sw, lc_stats.skipped += done, lc_stats.n_sum += n, lc_stats.nmax = lc_stats.nmax > (unsigned long long)n ? lc_stats.nmax : n));
	if (inserted && fixed) {                               // the one-pass form (see above)
		int below = 0;
		bool hazard = false;
		for (int i = ex.lane_id(); i < n; i += ex.n_lanes()) {
			const lc_alnreg_t &y = ma[i];
			const int64_t y_rb = y.rb, y_re = y.re;
			const int y_qb = y.qb, y_score = y.score;
			if (y_score > nb.score || (y_score == nb.score && (y_rb < nb.rb || (y_rb == nb.rb && y_qb < nb.qb)))) ++below;
			else if (y_score == nb.score && y_rb == nb.rb && y_qb == nb.qb) hazard = true;
			if (y.rid == nb.rid) {
				if (y_re <= nb.re && nb.rb < y_re + opt.max_chain_gap && lc_reg_redundant(opt, nb, y)) hazard = true;
				if (y_re >= nb.re && y_rb < nb.re + opt.max_chain_gap && lc_reg_redundant(opt, y, nb)) hazard = true;
			}
		}
		if (!ex.any(hazard)) {
			const int at = ex.sum(below);
			lc_regs_shift_up(ma, at, n, ex);
			if (ex.leader()) { nb.n_comp = 1; ma[at] = nb; }
			++n;
			done = true;
			LC_STAT(++lc_stats.onepass);
		} else LC_STAT(++lc_stats.hazard);
	}
	if (!done) {
		bool fx = false;
		LC_STAT(lc_stats.full_nofix += !fixed);
		if (ex.leader()) {
			if (inserted) {
				int at;
				++n;
				for (at = 0; at < n - 1; ++at) if (ma[at].score < nb.score) break;     // keep sorted by score (:165-171)
				for (int i = n - 1; i > at; --i) ma[i] = ma[i - 1];
				ma[at] = nb;
			}
			// :175; no patching without a query.  The sub-optimal-row list of the SW just done is dead by now: its memory holds the sort handles
			n = lc_sort_dedup_patch(b, (const uint8_t *)0, n, ma, (int *)0, (int *)0, (int *)ex.bl, ex.bl ? 2 * b.w_u64_n : 0, &fx);
		}
		n = (int)ex.bcast((unsigned long long)n);
		fixed = ex.bcast(fx ? 1ull : 0ull) != 0;
	}
	n_ma = n;
	ex.sync();
	return 1;
}

// Stage 5a, one THREAD per pair: the tail of mem_align1_core for both ends (mem_sort_dedup_patch, is_alt: B/bwamem.c:1070-1082),
// the final lists carved and filled, and the two rescue phases of GoBwaMemMateSW (G/gobwa.go:264-324) replayed for as long as
// every mem_matesw call is a no-op (the usual case: a properly paired hit satisfies every anchor).  The first anchor that needs
// a Smith-Waterman sends the pair to the rescue list; stage 5b (one warp per listed pair) then runs the phases for real.
// `ex`: one thread, or — for pairs with very many regions (lc_pair_prepare_defer) — the lanes of a warp: the copies and the list scans are
// shared out, the sorts and the dedup loop stay with the leader (they are the reference's statements in the reference's order).
template <class Exec>
LC_HD void lc_pair_prepare_t(const LcBatch &b, int p, int *H, int *E, const Exec &ex)
{
	int n[2], len[2], best[2] = { 0, 0 };
	bool clean[2] = { false, false };
	for (int e = 0; e < 2; ++e) { const int r = 2 * p + e; len[e] = (int)(b.off[r + 1] - b.off[r]); n[e] = b.r_n_reg[r]; }
	// capacity of the final lists: what stage 4 produced (the dedup can only shrink it) plus one region per rescue call
	const int cap0 = n[0] + (len[0] > 0 ? lc_min(b.max_anchors, n[1]) : 0);
	const int cap1 = n[1] + (len[1] > 0 ? lc_min(b.max_anchors, cap0) : 0);
	unsigned long long fo = 0;
	if (ex.leader()) {
		b.r_n_fin[2 * p] = b.r_n_fin[2 * p + 1] = 0; b.r_fin_off[2 * p] = b.r_fin_off[2 * p + 1] = 0;
		fo = lc_carve(b, b.fin_cur, b.fin_cap, (unsigned long long)cap0 + cap1, LC_OVF_FINAL);
	}
	fo = ex.bcast(fo);
	if (fo == (unsigned long long)-1) return;
	lc_alnreg_t *f0 = b.fin + fo, *f1 = b.fin + fo + cap0;
	for (int e = 0; e < 2; ++e) {
		const int r = 2 * p + e;
		lc_alnreg_t *dst = e ? f1 : f0;
		if (b.r_reg_off[r] >= 0) {                            // one slice of `reg`, in the reference's order
			const lc_alnreg_t *src = b.reg + b.r_reg_off[r];
			for (int i = ex.lane_id(); i < n[e]; i += ex.n_lanes()) dst[i] = src[i];
		} else {                                             // extended unit by unit (lc_ext_plan): chain by chain, which IS that order
			const long long c0 = (long long)b.r_chain_off[r];
			int m = 0;
			for (int ci = 0; ci < b.r_n_chain[r]; ++ci) {
				const LcChainAux &ax = b.chain_aux[c0 + ci];
				const lc_alnreg_t *src = b.reg + ax.reg_off;
				for (int i = ex.lane_id(); i < ax.n_reg; i += ex.n_lanes()) dst[m + i] = src[i];
				m += ax.n_reg;
			}
		}
		ex.sync();
		int ne = n[e];
		unsigned long long cl = 0, ko = (unsigned long long)-1;
		if (ex.leader() && ne >= 12 && b.sortk) {            // (a full pool only means the slower sort)
			ko = lc_atomic_add_agg(b.sortk_cur, (unsigned long long)ne);
			if (ko + (unsigned long long)ne > b.sortk_cap) ko = (unsigned long long)-1;
		}
		ko = ex.bcast(ko);
		if ((ex.n_lanes() > 1 || b.force_coop) && ko != (unsigned long long)-1 && b.regtmp) {      // tail of mem_align1_core (B/bwamem.c:1070), all lanes
			bool c1 = false;
			ne = lc_sort_dedup_patch_coop(b, b.seq + b.off[r], ne, dst, b.regtmp + ko, H, E, c1, b.sortk + ko, ex);
			clean[e] = c1; n[e] = ne;
		} else {
			if (ex.leader()) {
				bool c1 = false;
				ne = lc_sort_dedup_patch(b, b.seq + b.off[r], ne, dst, H, E, H, 2 * (b.lmax + 18), &c1, ko != (unsigned long long)-1 ? b.sortk + ko : nullptr);
				cl = c1 ? 1 : 0;
			}
			ex.sync();
			n[e] = (int)ex.bcast((unsigned long long)(unsigned)ne); clean[e] = ex.bcast(cl) != 0;
		}
		int bst = 0;
		for (int i = ex.lane_id(); i < n[e]; i += ex.n_lanes()) {
			if (dst[i].rid >= 0 && b.ix.anns[dst[i].rid].is_alt) dst[i].is_alt = 1;   // B/bwamem.c:1078-1082
			bst = lc_max(bst, dst[i].score);                                         // G/gobwa.go:264-283
		}
		best[e] = ex.max(bst);
		ex.sync();
	}
	if (ex.leader()) {
		b.r_fin_off[2 * p] = (int64_t)fo; b.r_n_fin[2 * p] = n[0];
		b.r_fin_off[2 * p + 1] = (int64_t)(fo + cap0); b.r_n_fin[2 * p + 1] = n[1];
	}
	bool need = false;
	int64_t rb, re;
	int num = 0;
	for (int i = 0; i < n[1] && num < b.max_anchors && len[0] > 0 && !need; ++i)
		if (f1[i].score >= best[1] - b.score_delta) { ++num; need = lc_matesw_wants_sw_t(b, f1[i], len[0], f0, n[0], rb, re, ex); }
	num = 0;
	for (int i = 0; i < n[0] && num < b.max_anchors && len[1] > 0 && !need; ++i)
		if (f0[i].score >= best[0] - b.score_delta) { ++num; need = lc_matesw_wants_sw_t(b, f0[i], len[1], f1, n[1], rb, re, ex); }
	if (need) {                                          // weight = candidates of the pair (anchors to try, lists to scan); bits 48, 49: the lists
		const int wgt = lc_min(n[0] + n[1], 0xffff);     // are fixed points already (lc_matesw_t) — a clean mem_sort_dedup_patch just made them
		if (ex.leader()) b.wl_in[lc_atomic_add_agg(b.resc_n, 1u)] = (uint64_t)((clean[0] ? 1 : 0) | (clean[1] ? 2 : 0)) << 48 | (uint64_t)(uint32_t)wgt << 32 | (uint32_t)p;
		return;
	}
	unsigned long long co = 0;
	if (ex.leader()) co = lc_atomic_add_agg(b.cand_cur, (unsigned long long)(n[0] + n[1]));   // candidate list for stage 6 (sized like fin)
	co = ex.bcast(co);
	for (int i = ex.lane_id(); i < n[0]; i += ex.n_lanes()) b.cand[co + i] = (uint64_t)(2 * p) << 32 | (uint32_t)i;
	for (int i = ex.lane_id(); i < n[1]; i += ex.n_lanes()) b.cand[co + n[0] + i] = (uint64_t)(2 * p + 1) << 32 | (uint32_t)i;
}
struct LcExecOnePair : LcExecOne {      // the thread form
	LC_HD bool leader() const { return true; }
	LC_HD void sync() const {}
	LC_HD unsigned long long bcast(unsigned long long v) const { return v; }
	LC_HD int max(int v) const { return v; }
};
// pairs with more regions than this are listed (wl_out, count in prep_heavy_n[0]) for the warp-per-pair launch of the same body
#define LC_PREP_HEAVY 96
LC_HD bool lc_pair_prepare_defer(const LcBatch &b, int p)
{
	if (!b.prep_heavy_n || b.r_n_reg[2 * p] + b.r_n_reg[2 * p + 1] <= LC_PREP_HEAVY) return false;
	b.wl_out[lc_atomic_add_agg(b.prep_heavy_n, 1u)] = (uint64_t)(uint32_t)p;
	return true;
}
LC_HD void lc_pair_prepare(const LcBatch &b, int p, int *H, int *E) { lc_pair_prepare_t(b, p, H, E, LcExecOnePair()); }

// Stage 5b, one warp (or thread) per listed pair: the two rescue phases on the lists lc_pair_prepare filled
template <class Exec>
LC_HD void lc_pair_rescue_t(const LcBatch &b, int p, int fixed_bits, const Exec &ex)
{
	int len[2], best[2] = { 0, 0 };
	const uint8_t *q[2];
	for (int e = 0; e < 2; ++e) { const int r = 2 * p + e; len[e] = (int)(b.off[r + 1] - b.off[r]); q[e] = b.seq + b.off[r]; }
	lc_alnreg_t *f0 = b.fin + b.r_fin_off[2 * p], *f1 = b.fin + b.r_fin_off[2 * p + 1];
	int n0 = b.r_n_fin[2 * p], n1 = b.r_n_fin[2 * p + 1], num = 0;
	bool fixed0 = (fixed_bits & 1) != 0, fixed1 = (fixed_bits & 2) != 0;      // see lc_matesw_t
	for (int i = 0; i < n0; ++i) best[0] = lc_max(best[0], f0[i].score);      // the pre-rescue best scores (G/gobwa.go:264-283)
	for (int i = 0; i < n1; ++i) best[1] = lc_max(best[1], f1[i].score);
	ex.sync();
	// phase 1: rescue read 1 from read 2's hits (G/gobwa.go:286-300). read 2's list is not touched in this phase.
	for (int i = 0; i < n1 && num < b.max_anchors && len[0] > 0; ++i)
		if (f1[i].score >= best[1] - b.score_delta) { ++num; lc_matesw_t(b, f1[i], len[0], q[0], f0, n0, ex, fixed0); }
	// phase 2: rescue read 2 from read 1's post-rescue list, gated by the pre-rescue best (G/gobwa.go:302-324)
	num = 0;
	for (int i = 0; i < n0 && num < b.max_anchors && len[1] > 0; ++i)
		if (f0[i].score >= best[0] - b.score_delta) { ++num; lc_matesw_t(b, f0[i], len[1], q[1], f1, n1, ex, fixed1); }
	unsigned long long co = 0;
	if (ex.leader()) {
		b.r_n_fin[2 * p] = n0; b.r_n_fin[2 * p + 1] = n1;
		co = lc_atomic_add_agg(b.cand_cur, (unsigned long long)(n0 + n1));
	}
	co = ex.bcast(co);
	for (int i = ex.lane_id(); i < n0; i += ex.n_lanes()) b.cand[co + i] = (uint64_t)(2 * p) << 32 | (uint32_t)i;
	for (int i = ex.lane_id(); i < n1; i += ex.n_lanes()) b.cand[co + n0 + i] = (uint64_t)(2 * p + 1) << 32 | (uint32_t)i;
}

struct LcExecThreadSw : LcExecThread {     // thread executor with the extra rows the local SW needs
	int *Hmax; unsigned long long *bl;
	LC_HD int lane_id() const { return 0; }
	LC_HD int n_lanes() const { return 1; }
	LC_HD bool any(bool v) const { return v; }
	LC_HD int sum(int v) const { return v; }
	template <class QF, class TF>
	LC_HD LcSwResult sw_align2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int xtra, LcCounters *cnt) const
	{
		return lc_sw_align2(qlen, q, tlen, t, sc, xtra, H, E, Hmax, bl, cnt);
	}
};

LC_HDN int lc_matesw(const LcBatch &b, const lc_alnreg_t &a, int l_ms, const uint8_t *ms, lc_alnreg_t *ma, int &n_ma, const LcSwScratch &ws)
{
	LcExecThreadSw ex;
	ex.H = ws.H; ex.E = ws.E; ex.Hmax = ws.Hmax; ex.bl = ws.b;
	bool fixed = false;
	return lc_matesw_t(b, a, l_ms, ms, ma, n_ma, ex, fixed);
}

// thread forms of the two halves of stage 5 (tests/hostsim; the device runs lc_pair_rescue_t with a warp)
LC_HDN void lc_stage_pair_prepare(const LcBatch &b, int p, int worker)
{
	int *wi = b.w_int + (size_t)worker * b.w_int_n;
	lc_pair_prepare(b, p, wi, wi + (b.lmax + 18));
}
LC_HDN void lc_stage_pair_rescue(const LcBatch &b, int p, int fixed_bits, int worker)
{
	int *wi = b.w_int + (size_t)worker * b.w_int_n;
	LcExecThreadSw ex;
	ex.H = wi; ex.E = wi + (b.lmax + 18); ex.Hmax = wi + 2 * (b.lmax + 18);
	ex.bl = b.w_u64 + (size_t)worker * b.w_u64_n;
	lc_pair_rescue_t(b, p, fixed_bits, ex);
}

// ============================================================================================================
// Stage 6 — final alignment of one candidate: mem_reg2aln (B/bwamem.c:1086-1156) -> bwa_gen_cigar2 (B/bwa.c:121-207)
// ============================================================================================================
LC_HD int lc_infer_bw(int l1, int l2, int score, int a, int q, int r)   // B/bwamem.c:792-799
{
	if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
	int w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
	int d = l1 > l2 ? l1 - l2 : l2 - l1;
	return w < d ? d : w;
}

struct LcExecThreadCigar : LcExecThread {   // thread executor for the CIGAR stage
	uint32_t *cig; uint8_t *zloc;
	LC_HD int lane_id() const { return 0; }
	LC_HD int n_lanes() const { return 1; }
	LC_HD int sum(int v) const { return v; }
	template <class QF, class TF>
	LC_HD int global2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int w, uint8_t *z, uint32_t *cigar, int cap, int *n_cigar, LcCounters *cnt) const
	{
		return lc_global2(qlen, q, tlen, t, sc, w, H, E, z, cigar, cap, n_cigar, cnt);
	}
};

// ex.cig: per-worker CIGAR build buffer (2*lmax+8 words); ex.zloc: per-worker traceback scratch of b.w_z_n bytes
template <class Exec>
LC_HD void lc_stage_cigar_t(const LcBatch &b, uint64_t cd, const Exec &ex)
{
	const lc_opt_t &opt = b.opt;
	const int r = (int)(cd >> 32), idx = (int)(uint32_t)cd;
	const int l_query = (int)(b.off[r + 1] - b.off[r]);
	const uint8_t *query = b.seq + b.off[r];
	const size_t slot = (size_t)b.r_fin_off[r] + idx;
	const lc_alnreg_t &ar = b.fin[slot];
	lc_aln_t out;
	memset(&out, 0, sizeof(out));
	if (ar.rb < 0 || ar.re < 0) { out.rid = -1; out.pos = -1; if (ex.leader()) b.aln[slot] = out; return; }
	const int64_t l_pac = b.ix.l_pac, rb = ar.rb, re = ar.re;
	const int qb = ar.qb, qe = ar.qe, lq = qe - qb, rlen = (int)(re - rb);
	const bool rev = rb >= l_pac;
	int tmp = lc_infer_bw(lq, rlen, ar.truesc, opt.a, opt.o_del, opt.e_del);
	int w2 = lc_infer_bw(lq, rlen, ar.truesc, opt.a, opt.o_ins, opt.e_ins);
	w2 = lc_max(w2, tmp);
	if (w2 > opt.w) w2 = lc_min(w2, ar.w);
	// the CIGAR is built in a per-worker buffer (an alignment of lq x rlen has at most 2*min(lq,rlen)+1 merged ops) and
	// only its final ops are carved from the output pool
	const int cig_cap = 2 * b.lmax + 8;
	uint32_t *cig = ex.cig;
	int n_cigar = 0, score = 0, last_sc = -(1 << 30), i = 0;
	const bool bad = lq <= 0 || rb >= re || (rb < l_pac && re > l_pac) || re > l_pac << 1;
	do {
		w2 = lc_min(w2, opt.w << 2);
		n_cigar = 0;
		if (bad) { score = 0; break; }                       // bwa_gen_cigar2 returns no CIGAR (B/bwa.c:132-134)
		if (lq == rlen && w2 == 0) {                         // gap-free (B/bwa.c:141-149)
			if (ex.leader()) cig[0] = (uint32_t)lq << 4;
			n_cigar = 1;
			int part = 0;
			for (int k = ex.lane_id(); k < lq; k += ex.n_lanes()) part += b.sc.sub(lc_ref_base(b.ix, rb + k), query[qb + k]);
			score = ex.sum(part);
			ex.sync();
		} else {
			int mi = (int)((double)(((lq + 1) >> 1) * b.sc.a - b.sc.o_ins) / b.sc.e_ins + 1.);
			int md = (int)((double)(((lq + 1) >> 1) * b.sc.a - b.sc.o_del) / b.sc.e_del + 1.);
			int mg = lc_max(lc_max(mi, md), 1), dl = rlen > lq ? rlen - lq : lq - rlen;
			int w = (mg + dl + 1) >> 1;
			w = lc_min(w, w2);
			w = lc_max(w, dl + 3);
			const int n_col = lq < 2 * w + 1 ? lq : 2 * w + 1;
			const unsigned long long zneed = (unsigned long long)n_col * rlen + 8;
			uint8_t *z;
			if (zneed <= (unsigned long long)b.w_z_n) z = ex.zloc;   // common case: per-worker traceback scratch
			else {
				unsigned long long zo = 0;
				if (ex.leader()) zo = lc_carve(b, b.z_cur, b.z_cap, zneed, LC_OVF_Z);
				zo = ex.bcast(zo);
				if (zo == (unsigned long long)-1) return;
				z = b.zpool + zo;
			}
			if (rev) {
				LcSeqRev qf = { query + qe - 1 }; LcRefRev tf = { &b.ix, re - 1 };
				score = ex.global2(lq, qf, rlen, tf, b.sc, w, z, cig, cig_cap, &n_cigar, b.cnt);
			} else {
				LcSeqFwd qf = { query + qb }; LcRefFwd tf = { &b.ix, rb };
				score = ex.global2(lq, qf, rlen, tf, b.sc, w, z, cig, cig_cap, &n_cigar, b.cnt);
			}
		}
		if (score == last_sc || w2 == opt.w << 2) break;
		last_sc = score;
		w2 <<= 1;
	} while (++i < 3 && score < ar.truesc - opt.a);
	// NM = mismatches + gap bases, skipping a leading/trailing deletion (B/bwa.c:169-199)
	int NM = -1, n_mm = 0;
	if (!bad && n_cigar >= 0) {
		int x = 0, y = 0, n_gap = 0, part = 0;
		for (int k = 0; k < n_cigar; ++k) {
			const int op = cig[k] & 0xf, len = (int)(cig[k] >> 4);
			if (op == 0) {
				for (int t = ex.lane_id(); t < len; t += ex.n_lanes()) {
					int qc = rev ? query[qe - 1 - (x + t)] : query[qb + x + t];
					int rc = rev ? lc_ref_base(b.ix, re - 1 - (y + t)) : lc_ref_base(b.ix, rb + y + t);
					part += qc != rc;
				}
				x += len; y += len;
			} else if (op == 2) { if (k > 0 && k < n_cigar - 1) n_gap += len; y += len; }
			else if (op == 1) { x += len; n_gap += len; }
		}
		n_mm = ex.sum(part);
		NM = n_mm + n_gap;
	}
	int64_t pos = rev ? (l_pac << 1) - 1 - (re - 1) : rb;       // bns_depos (B/bwamem.c:1122)
	if (n_cigar < 0) { if (ex.leader()) lc_raise(b, LC_OVF_CIGAR); return; }
	int cs = 0, ce = n_cigar;
	if (n_cigar > 0) {                                       // squeeze out a leading or trailing deletion (B/bwamem.c:1124-1133)
		if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; cs = 1; }
		else if ((cig[ce - 1] & 0xf) == 2) --ce;
	}
	const bool clipped = qb != 0 || qe != l_query;          // soft clips (B/bwamem.c:1134-1148)
	const int clip5 = clipped ? (rev ? l_query - qe : qb) : 0, clip3 = clipped ? (rev ? qb : l_query - qe) : 0;
	const int o = (clip5 ? 1 : 0) + (ce - cs) + (clip3 ? 1 : 0);
	unsigned long long co = 0;
	if (ex.leader()) co = lc_carve(b, b.cigar_cur, b.cigar_cap, (unsigned long long)o, LC_OVF_CIGAR);
	co = ex.bcast(co);
	if (co == (unsigned long long)-1) return;
	if (ex.leader()) {
		uint32_t *dst = b.cigar + co;
		int k = 0;
		if (clip5) dst[k++] = (uint32_t)clip5 << 4 | 3;
		for (int t = cs; t < ce; ++t) dst[k++] = cig[t];
		if (clip3) dst[k++] = (uint32_t)clip3 << 4 | 3;
	}
	out.pos = pos; out.is_rev = rev ? 1 : 0; out.NM = NM; out.n_cigar = o; out.cigar_off = (uint32_t)co; out.n_mismatch = n_mm;
	out.rid = lc_pos2rid(b.ix, pos);
	if (out.rid >= 0) out.pos = pos - b.ix.anns[out.rid].offset;
	if (ex.leader()) b.aln[slot] = out;
}

LC_HDN void lc_stage_cigar(const LcBatch &b, uint64_t cd, int worker)
{
	LcExecThreadCigar ex;
	ex.H = b.w_int + (size_t)worker * b.w_int_n; ex.E = ex.H + (b.lmax + 2);
	ex.cig = (uint32_t *)(ex.E + (b.lmax + 2)); ex.zloc = b.w_z + (size_t)worker * b.w_z_n;
	lc_stage_cigar_t(b, cd, ex);
}
