// lc_extsm.cuh — stage 4 (mem_chain2aln, B/bwamem.c:632-786, with ksw_extend2, B/ksw.c:380-479) as a per-lane state machine.
//
// The first GPU form of this stage gave one warp to each read and ran every ksw_extend2 row with 32 lanes (a prefix-max
// scan for F, band-shrink ballots, row-maximum reductions).  ncu (profiles/r01e_extend_kernel_details.txt) showed ~310
// warp-instructions per DP row for rows that are on average 44 cells wide: the fixed per-row cost of the cooperative form
// dominates, and the chain2aln control flow between the DP calls is executed 32 times over.
// Here every lane owns one read and runs the reference's scalar algorithm, resumably:
//   * the DP advances LC_XT_CELLS cells of the current row per tick, all lanes in lock step (no shuffles, no scans);
//   * the H/E row lives in shared memory as one packed word {h:16, e:16} per column, laid out [column][lane] so that the
//     32 lanes never conflict; the read is kept there too, 4 bits per base;
//   * end-of-row work (z-drop, band shrink, next target base) and the chain2aln bookkeeping between extensions are guarded
//     sections with warp-wide convergence points between them, like the seeding state machine (lc_seed.cuh).
// The code is plain scalar C++ templated on the memory stride, so tests/hostsim runs the very same state machine on the CPU
// (stride 1) against the oracle; the device kernel (lc_kernels.cu) uses stride 32.
#pragma once
#include "lc_mem.cuh"

#ifndef LC_CONVERGE
#ifdef __CUDA_ARCH__
#define LC_CONVERGE() __syncwarp()
#else
#define LC_CONVERGE() ((void)0)
#endif
#endif

#define LC_XT_CELLS 8

// ---- two 16-bit lanes per register: the DP arithmetic of a tick runs on column pairs (low half = even column).  On sm_100a each of
// these is one instruction (VIADD.16x2, VIMNMX.S16x2, VIADDMNMX.S16x2[.RELU], VIMNMX3.S16x2, VIMNMX.U16x2, PRMT); tests/hostsim
// runs the portable forms.
#ifdef __CUDA_ARCH__
LC_HD uint32_t lc_p2_add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
LC_HD uint32_t lc_p2_max(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
LC_HD uint32_t lc_p2_min(uint32_t a, uint32_t b) { return __vmins2(a, b); }
LC_HD uint32_t lc_p2_minu(uint32_t a, uint32_t b) { return __vminu2(a, b); }
LC_HD uint32_t lc_p2_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
LC_HD uint32_t lc_p2_addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }            // max(a + b, c)
LC_HD uint32_t lc_p2_addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }  // max(a + b, c, 0)
// PRMT through inline PTX: __byte_perm() masks selector bit 3 (the sign-replicate mode used for 16 -> 32 bit sign extension)
__device__ __forceinline__ uint32_t lc_prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
#else
LC_HD int lc_p2_lo(uint32_t a) { return (int)(int16_t)(a & 0xffffu); }
LC_HD int lc_p2_hi(uint32_t a) { return (int)(int16_t)(a >> 16); }
LC_HD uint32_t lc_p2_mk(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi & 0xffffu) << 16; }
LC_HD uint32_t lc_p2_add(uint32_t a, uint32_t b) { return lc_p2_mk(lc_p2_lo(a) + lc_p2_lo(b), lc_p2_hi(a) + lc_p2_hi(b)); }
LC_HD uint32_t lc_p2_max(uint32_t a, uint32_t b) { return lc_p2_mk(lc_max(lc_p2_lo(a), lc_p2_lo(b)), lc_max(lc_p2_hi(a), lc_p2_hi(b))); }
LC_HD uint32_t lc_p2_min(uint32_t a, uint32_t b) { return lc_p2_mk(lc_min(lc_p2_lo(a), lc_p2_lo(b)), lc_min(lc_p2_hi(a), lc_p2_hi(b))); }
LC_HD uint32_t lc_p2_minu(uint32_t a, uint32_t b) { return lc_min(a & 0xffffu, b & 0xffffu) | lc_min(a >> 16, b >> 16) << 16; }
LC_HD uint32_t lc_p2_max3(uint32_t a, uint32_t b, uint32_t c) { return lc_p2_max(lc_p2_max(a, b), c); }
LC_HD uint32_t lc_p2_addmax(uint32_t a, uint32_t b, uint32_t c) { return lc_p2_max(lc_p2_add(a, b), c); }
LC_HD uint32_t lc_p2_addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return lc_p2_max(lc_p2_max(lc_p2_add(a, b), c), 0u); }
LC_HD uint32_t lc_prmt(uint32_t a, uint32_t b, uint32_t sel)      // PTX prmt.b32, default mode (selector bit 3 = replicate the byte's sign)
{
	const uint64_t pool = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		const uint32_t n = (sel >> (4 * i)) & 15u;
		uint32_t byte = (uint32_t)(pool >> (8 * (n & 7u))) & 0xffu;
		if (n & 8u) byte = (byte & 0x80u) ? 0xffu : 0u;
		r |= byte << (8 * i);
	}
	return r;
}
#endif
LC_HD uint32_t lc_p2_dup(int v) { return ((uint32_t)v & 0xffffu) * 0x10001u; }      // the same 16-bit value in both halves

// ---- the H/E row in (shared) memory.  Two packings, chosen by the longest read of the batch:
//   LcCells16: one column per 32-bit word, h | e << 16                        (any read length up to LC_MAX_READ_LEN)
//   LcCells8 : two columns per word, (h | e << 8) << 16 * (column & 1)        (lmax * a <= 255: no cell of ksw_extend2 can exceed
//              the score of the whole read matching, so h and e fit a byte) — half the shared memory, twice the resident warps
// load8/store8 move the eight columns of a tick as four column pairs: h[k] = {h(col0 + 2k), h(col0 + 2k + 1)} in 16-bit halves.
template <int STRIDE> struct LcCells16 {
	static LC_HD int words(int cols) { return cols; }
	static LC_HD void load8(const uint32_t *he, int col0, uint32_t *h, uint32_t *e)
	{
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t w0 = he[(col0 + 2 * k) * STRIDE], w1 = he[(col0 + 2 * k + 1) * STRIDE];
			h[k] = lc_prmt(w0, w1, 0x5410u); e[k] = lc_prmt(w0, w1, 0x7632u);
		}
	}
	static LC_HD void store8(uint32_t *he, int col0, const uint32_t *h, const uint32_t *e)
	{
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			he[(col0 + 2 * k) * STRIDE] = lc_prmt(h[k], e[k], 0x5410u);
			he[(col0 + 2 * k + 1) * STRIDE] = lc_prmt(h[k], e[k], 0x7632u);
		}
	}
	static LC_HD void put(uint32_t *he, int col, int h, int e) { he[col * STRIDE] = (uint32_t)h | (uint32_t)e << 16; }
	static LC_HD void clear_left(uint32_t *he, int beg, const uint32_t *)      // columns [beg & ~7, beg) := {0, 0}
	{
		for (int c = beg & ~(LC_XT_CELLS - 1); c < beg; ++c) he[c * STRIDE] = 0u;
	}
};
template <int STRIDE> struct LcCells8 {
	static LC_HD int words(int cols) { return cols >> 1; }
	static LC_HD void load8(const uint32_t *he, int col0, uint32_t *h, uint32_t *e)      // col0 is a multiple of 8
	{
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const uint32_t v = he[((col0 >> 1) + k) * STRIDE];
			h[k] = lc_prmt(v, 0u, 0x4240u); e[k] = lc_prmt(v, 0u, 0x4341u);
		}
	}
	// live columns hold values < 256; a column at or right of the row's end may hold 256 and spill into its own e byte or into the
	// column to its right, which is dead as well (see tick())
	static LC_HD void store8(uint32_t *he, int col0, const uint32_t *h, const uint32_t *e)
	{
#pragma unroll
		for (int k = 0; k < 4; ++k) he[((col0 >> 1) + k) * STRIDE] = h[k] + (e[k] << 8);
	}
	static LC_HD void put(uint32_t *he, int col, int h, int e)
	{
		uint32_t *p = he + (col >> 1) * STRIDE;
		const int sh = (col & 1) << 4;
		*p = (*p & ~(0xffffu << sh)) | ((uint32_t)h | (uint32_t)e << 8) << sh;
	}
	static LC_HD void clear_left(uint32_t *he, int beg, const uint32_t *mtab)      // columns [beg & ~7, beg) := {0, 0}: four masked words, no loop
	{
		const uint32_t *m = mtab + 4 * (beg & (LC_XT_CELLS - 1));                   // masks of the columns c < (beg & 7) of the group
		uint32_t *p = he + ((beg & ~(LC_XT_CELLS - 1)) >> 1) * STRIDE;
#pragma unroll
		for (int k = 0; k < 4; ++k) p[k * STRIDE] &= ~m[k];
	}
};

// masks of the columns c < hi of a tick (hi = 0..8) as four column pairs; the device kernel keeps the table in shared memory
LC_HD void lc_ext_fill_masks(uint32_t *tab)      // 9 x 4 words
{
	for (int hi = 0; hi <= 8; ++hi)
		for (int k = 0; k < 4; ++k) tab[4 * hi + k] = (2 * k < hi ? 0xffffu : 0u) | (2 * k + 1 < hi ? 0xffff0000u : 0u);
}

enum { LC_XP_FETCH = 0, LC_XP_AFTER_LEFT, LC_XP_AFTER_RIGHT, LC_XP_RIGHT_CHECK, LC_XP_FINALIZE, LC_XP_READ_START, LC_XP_CHAIN_START, LC_XP_SEED_NEXT,
	   LC_XP_LEFT_SETUP, LC_XP_RIGHT_SETUP, LC_XP_READ_DONE,
	   LC_XP_DP, LC_XP_ROWEND, LC_XP_DONE };   // states >= LC_XP_DP: a DP row is in progress or waits for its end-of-row work (or the lane is finished)

// UNIT: the lane's work item is one unit of a read with very many chains (lc_ext_plan, lc_mem.cuh) instead of a whole read: the chains
// are followed through LcChainAux::next, every chain records where its regions went, the read's region count is a sum over its units
template <int STRIDE, class CELLS = LcCells16<STRIDE>, bool UNIT = false> struct LcExtLane {
	const LcBatch &b;
	uint32_t *he;            // H/E row of lmax + 2 columns, packed by CELLS
	uint32_t *qw;            // the read: qw[w * STRIDE] holds bases 8w..8w+7, 4 bits each, first base in the low nibble; qw[-STRIDE] is
	                         //   addressable.  Left extensions run on reversed sequences (B/bwamem.c:715-721): their columns are read from the
	                         //   same copy back to front
	int phase;
	// ---- read-level state (mem_chain2aln's locals)
	int r, l_query, NC, ci, n_av, k, cn, tryi, prev;
	long long cg; int n_av0;     // UNIT: the current chain (index into b.chain, < 0 = the unit is done), regions before it
	unsigned long long ro;
	lc_alnreg_t *av;
	uint64_t *srt;
	const LcSeed *seeds;
	int64_t r0, r1, s_rbeg, a_rb, a_re, re0;
	int s_qbeg, s_len, c_rid, a_qb, a_qe, a_score, a_truesc, aw0, aw1, qe0, sc0;
	// ---- the extension in progress (ksw_extend2's locals)
	int qlen, tlen, w, h0, zdrop, i, j, beg, end, f, hleft, hleft0, rowkey, best, best_i, best_j, g_i, gscore, max_off, tb;
	int nz_first, nz_last;   // first / last column of the row whose updated {h, e} word is non-zero (the band shrink of B/ksw.c:466-469)
	const uint32_t *mtab;    // lc_ext_fill_masks table (shared memory on the device)
	int64_t fpos; int fdir, comp;          // forward-strand coordinate of the current row's target base, its step per row, 3 on the reverse strand
	uint32_t pw, pwn; int64_t pwi;         // the 16 packed reference bases around fpos, the next such word in walking direction (prefetched), word index
	int q0, qdir;            // the extension's query column c is base q0 + c of the forward (qdir > 0) or reversed (qdir < 0) read copy
	unsigned long long cells;

	LC_HD LcExtLane(const LcBatch &b_, uint32_t *he_, uint32_t *qw_, const uint32_t *mtab_)
		: b(b_), he(he_), qw(qw_), phase(LC_XP_FETCH), l_query(0), j(0), beg(0), end(0), f(0), hleft(0), tb(0), mtab(mtab_), q0(0), qdir(1) {}

	LC_HD uint32_t pac_word(int64_t wi) const { return lc_ldg((const uint32_t *)b.ix.pac + (wi < 0 ? 0 : wi)); }
	LC_HD int target_base()              // base of the current row: bns_get_seq's decoding (B/bntseq.c:398-419) out of the cached word
	{
		if ((fpos >> 4) != pwi) { pw = pwn; pwi += fdir; pwn = pac_word(pwi + fdir); }
		return (int)((pw >> ((((int)(fpos >> 2) & 3) << 3) + ((~(int)fpos & 3) << 1))) & 3) ^ comp;
	}
	LC_HD void begin_read(int r_) { r = r_; phase = LC_XP_READ_START; }
	LC_HD void begin_unit(long long first_chain) { cg = first_chain; r = b.chain_aux[first_chain].r; phase = LC_XP_READ_START; }

	// ---- ksw_extend2 prologue (B/ksw.c:385-412) and the first row's set-up
	LC_HD void dp_setup(int qlen_, int q0_, int qdir_, int tlen_, int64_t t0_, int tdir_, int w_, int end_bonus, int h0_)
	{
		const LcScore &sc = b.sc;
		const int oe_ins = sc.o_ins + sc.e_ins;
		qlen = qlen_; q0 = q0_; qdir = qdir_; tlen = tlen_; h0 = h0_; zdrop = b.opt.zdrop;
		if (t0_ < b.ix.l_pac) { fpos = t0_; fdir = tdir_; comp = 0; }      // the window never crosses the strand boundary (B/bwamem.c:655-657)
		else { fpos = (b.ix.l_pac << 1) - 1 - t0_; fdir = -tdir_; comp = 3; }
		best = h0; best_i = best_j = -1; g_i = -1; gscore = -1; max_off = 0; beg = 0; end = qlen; cells = 0;
		{
			int v = h0 > oe_ins ? h0 - oe_ins : 0;           // row -1: H[0] = h0, H[1] = h0 - oe_ins, then -e_ins per column while positive; E = 0
			uint32_t hh[4], ee[4] = { 0u, 0u, 0u, 0u };
			for (int t0 = 0; t0 <= qlen + 1; t0 += LC_XT_CELLS) {
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					uint32_t pr = 0;
#pragma unroll
					for (int hf = 0; hf < 2; ++hf) {
						int x;
						if (t0 + 2 * k + hf == 0) x = h0;
						else { x = v; v = v > sc.e_ins ? v - sc.e_ins : 0; }
						pr |= (uint32_t)x << (16 * hf);
					}
					hh[k] = pr;
				}
				CELLS::store8(he, t0, hh, ee);
			}
		}
		{
			const int smax = sc.a > 0 ? sc.a : 0;
			const int mi = (int)((double)(qlen * smax + end_bonus - sc.o_ins) / sc.e_ins + 1.);
			const int md = (int)((double)(qlen * smax + end_bonus - sc.o_del) / sc.e_del + 1.);
			w = lc_min(w_, lc_max(mi, 1));
			w = lc_min(w, lc_max(md, 1));
		}
		i = 0; j = 0;
		if (tlen <= 0) { dp_finish(); return; }
		pwi = fpos >> 4; pw = pac_word(pwi); pwn = pac_word(pwi + fdir);
		row_start();
		phase = LC_XP_DP;
	}
	LC_HD void row_start()                                    // B/ksw.c:421-430
	{
		const LcScore &sc = b.sc;
		if (beg < i - w) {
			// the band cuts live columns off on the left (B/ksw.c:421).  tick() relies on every column left of beg holding {0, 0}
			// (true for columns dropped by the zero scan of B/ksw.c:466); clear the ones that share the first tick with beg
			beg = i - w;
			CELLS::clear_left(he, beg, mtab);
		}
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		hleft = 0;
		if (beg == 0) { hleft = h0 - (sc.o_del + sc.e_del * (i + 1)); if (hleft < 0) hleft = 0; }
		hleft0 = hleft;
		f = 0; rowkey = -1; j = beg & ~(LC_XT_CELLS - 1);      // ticks are aligned to groups of LC_XT_CELLS columns
		nz_first = 1 << 30; nz_last = -1;
		tb = target_base();
	}
	LC_HD void dp_finish()                                    // hands {best, qle, tle, gtle, gscore, max_off} to the caller's state
	{
		if (b.cnt) { lc_atomic_add(&b.cnt->ext_calls, 1ull); lc_atomic_add(&b.cnt->ext_cells, cells); }
		phase = qdir < 0 ? LC_XP_AFTER_LEFT : LC_XP_AFTER_RIGHT;
	}

	// ---- one tick: the LC_XT_CELLS columns j..j+7 of the current row (B/ksw.c:431-456), then, if the row is complete, the
	// end-of-row work (B/ksw.c:457-478).  The eight cells are computed branch-free on column pairs (two 16-bit lanes per register):
	//   M = H(i-1,j-1) ? H(i-1,j-1) + s : 0      as   min(max(diag + s, 0), diag * (a + 1))     [a negative M acts like 0 everywhere it is used]
	//   E' = max(E - e_del, M - oe_del, 0)        one VIADD + one VIADDMNMX.RELU per pair
	//   F(j+1) = max(F(j) - e_ins, M - oe_ins, 0) does not depend on H, so the eight F values are a decayed prefix maximum: a three-step
	//                                             scan over the four pairs, then the carried-in f joins with its own decay
	//   H = max(M, E, F)
	// Substitution scores come from an 8-entry byte table indexed by (query code ^ target code) through PRMT.
	// Columns left of beg hold {0, 0} (row_start) and stay {0, 0} under these recurrences, so they need no mask; columns at or right of
	// end are never read by a later row before `put` rewrites them (each row reads at most one column past the previous end), so
	// they are computed and stored as garbage, and only kept out of the row maximum and the non-zero scan (mask table mtab).
	// `row_gate`: the end-of-row work runs only on ticks where the whole warp does it (lc_kernels.cu); a lane whose row completed on
	// another tick waits in LC_XP_ROWEND (its DP part idles) until then.
	LC_HD void tick(bool row_gate = true)
	{
		const LcScore &sc = b.sc;
		const int e_del = sc.e_del, e_ins = sc.e_ins;
		const bool run = phase == LC_XP_DP;
		const int col0 = run ? j : 0;
		uint32_t h[4], e2[4], hs[4];
		{
			// query codes of the eight columns: one funnel shift over two words of the read copy; a left extension reads the read
			// back to front (column c of the reversed read = base l_query - 1 - q0 - c), so its window is nibble-reversed
			uint32_t qword;
			{
				const int qi = qdir > 0 ? q0 + col0 : l_query - LC_XT_CELLS - q0 - col0;      // first base of the window, >= -8
				const uint32_t w0 = qw[(qi >> 3) * STRIDE], w1 = qw[((qi >> 3) + 1) * STRIDE];
				uint32_t v = (uint32_t)(((uint64_t)w1 << 32 | w0) >> ((qi & 7) << 2));
				if (qdir < 0) {
					v = lc_prmt(v, 0u, 0x0123u);
					v = (v & 0x0f0f0f0fu) << 4 | (v >> 4 & 0x0f0f0f0fu);
				}
				qword = v;
			}
			// 255 - penalty per column: 0 for a match, a + b for a mismatch, a + 1 when either base is ambiguous (score -1, B/bwa.c:109-118)
			const uint32_t x = tb < 4 ? qword ^ ((uint32_t)tb * 0x11111111u) : 0x44444444u;
			const LcExtConst &xc = b.xc;
			const uint32_t lut_lo = xc.lut_lo, lut_hi = xc.lut_hi;
			const uint32_t pa = lc_prmt(lut_lo, lut_hi, x), pb = lc_prmt(lut_lo, lut_hi, x >> 16);
			uint32_t Q[4];
			Q[0] = lc_prmt(pa, 0u, 0x4140u); Q[1] = lc_prmt(pa, 0u, 0x4342u); Q[2] = lc_prmt(pb, 0u, 0x4140u); Q[3] = lc_prmt(pb, 0u, 0x4342u);
			const uint32_t c_a = xc.c_a, n_oe_del = xc.n_oe_del, n_e_del = xc.n_e_del, n_oe_ins = xc.n_oe_ins, n_e1 = xc.n_e1, n_e2 = xc.n_e2, n_e4 = xc.n_e4;
			const uint32_t kmul = xc.kmul;
			uint32_t hd[4], ev[4], me[4], A[4];
			CELLS::load8(he, col0, hd, ev);
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const uint32_t mr = lc_p2_addmax_relu(lc_p2_add(hd[k], c_a), Q[k], 0u);
				const uint32_t M = lc_p2_min(mr, hd[k] * kmul);
				e2[k] = lc_p2_addmax_relu(ev[k], n_e_del, lc_p2_add(M, n_oe_del));
				A[k] = lc_p2_addmax_relu(M, n_oe_ins, 0u);
				me[k] = lc_p2_max(M, ev[k]);
			}
			// decayed inclusive prefix maximum of A over the eight columns (A >= 0, so a mildly negative filler never wins)
			const uint32_t neg = 0xc000c000u;
			{
				const uint32_t s0 = lc_prmt(neg, A[0], 0x5432u), s1 = lc_prmt(A[0], A[1], 0x5432u), s2 = lc_prmt(A[1], A[2], 0x5432u), s3 = lc_prmt(A[2], A[3], 0x5432u);
				A[0] = lc_p2_addmax(s0, n_e1, A[0]); A[1] = lc_p2_addmax(s1, n_e1, A[1]); A[2] = lc_p2_addmax(s2, n_e1, A[2]); A[3] = lc_p2_addmax(s3, n_e1, A[3]);
			}
			A[3] = lc_p2_addmax(A[2], n_e2, A[3]); A[2] = lc_p2_addmax(A[1], n_e2, A[2]); A[1] = lc_p2_addmax(A[0], n_e2, A[1]);
			A[3] = lc_p2_addmax(A[1], n_e4, A[3]); A[2] = lc_p2_addmax(A[0], n_e4, A[2]);
			// G(c) = F entering column c + 1: the scan joined with the carried-in f, which has decayed c + 1 times by then
			const uint32_t fb = lc_p2_dup(f);
			uint32_t G[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				G[k] = lc_p2_addmax(fb, xc.dec[k], A[k]);
			}
			// F entering each column = G shifted right by one column with f in front; H = max(M, E, F)
			h[0] = lc_p2_max(me[0], lc_prmt(fb, G[0], 0x5432u));
			h[1] = lc_p2_max(me[1], lc_prmt(G[0], G[1], 0x5432u));
			h[2] = lc_p2_max(me[2], lc_prmt(G[1], G[2], 0x5432u));
			h[3] = lc_p2_max(me[3], lc_prmt(G[2], G[3], 0x5432u));
			// what the next row reads at column c is H(i, c - 1): the row shifted right by one column with hleft in front
			hs[0] = lc_prmt((uint32_t)hleft << 16, h[0], 0x5432u); hs[1] = lc_prmt(h[0], h[1], 0x5432u);
			hs[2] = lc_prmt(h[1], h[2], 0x5432u); hs[3] = lc_prmt(h[2], h[3], 0x5432u);
			if (run) CELLS::store8(he, col0, hs, e2);
			// row maximum with the reference's tie rule (the last column wins, B/ksw.c:443-444): keys h * 8 + column, dead columns -1
			int hi = end - col0;
			hi = hi < 0 ? 0 : (hi > LC_XT_CELLS ? LC_XT_CELLS : hi);
			const uint32_t *mrow = mtab + 4 * hi;
			const uint32_t m0 = mrow[0], m1 = mrow[1], m2 = mrow[2], m3 = mrow[3];
			{
				const uint32_t k0 = (h[0] * 8u + 0x00010000u) | ~m0, k1 = (h[1] * 8u + 0x00030002u) | ~m1,
					k2 = (h[2] * 8u + 0x00050004u) | ~m2, k3 = (h[3] * 8u + 0x00070006u) | ~m3;
				const uint32_t kk = lc_p2_max(lc_p2_max3(k0, k1, k2), k3);
				const int klo = (int)lc_prmt(kk, 0u, 0x9910u), khi = (int)kk >> 16;
				const int tk = klo > khi ? klo : khi;                      // -1 when no column of the tick is live
				const int tk32 = run ? (tk & ~7) * 128 + (tk & 7) + col0 : -1;   // h << 10 | column; negative for tk == -1 (columns stay below 1017)
				rowkey = rowkey > tk32 ? rowkey : tk32;
			}
			// first / last live column of the row whose stored {h, e} is non-zero
			{
				const uint32_t one = 0x00010001u;
				const uint32_t b0 = lc_p2_minu((hs[0] | e2[0]) & m0, one), b1 = lc_p2_minu((hs[1] | e2[1]) & m1, one),
					b2 = lc_p2_minu((hs[2] | e2[2]) & m2, one), b3 = lc_p2_minu((hs[3] | e2[3]) & m3, one);
				const uint32_t acc = b0 + b1 * 4u + b2 * 16u + b3 * 64u;
				const uint32_t b8 = run ? (acc | acc >> 15) & 0xffu : 0u;    // bit c = column col0 + c
				if (b8) {
					const int first = col0 + lc_ctz(b8), last = col0 + 31 - lc_clz(b8);
					nz_first = nz_first < first ? nz_first : first;
					nz_last = last;
				}
			}
			f = (int)G[3] >> 16;
			if (run) {
				j = col0 + LC_XT_CELLS;
				hleft = (int)(h[3] >> 16);
				if (j >= end) {                                    // the row is complete: keep h of its last live column (this tick's column end - 1 - col0)
					if (end > beg) {
						const int idx = end - 1 - col0;
						const uint32_t pr = idx < 4 ? (idx < 2 ? h[0] : h[1]) : (idx < 6 ? h[2] : h[3]);
						hleft = (int)((idx & 1) ? pr >> 16 : pr & 0xffffu);
					} else hleft = hleft0;
					phase = LC_XP_ROWEND;
				}
			}
		}
		LC_CONVERGE();
		if (row_gate && phase == LC_XP_ROWEND) {               // end-of-row work (B/ksw.c:457-478)
			bool stop = false;
			const int jfin = beg < end ? end : beg;            // the reference's loop variable after the column loop
			if (b.cnt) cells += (unsigned long long)(end > beg ? end - beg : 0);
			const int rowmax = rowkey < 0 ? 0 : rowkey >> 10, rowarg = rowkey < 0 ? -1 : (rowkey & 1023);
			CELLS::put(he, end, hleft, 0);
			if (jfin == qlen) {
				if (!(gscore > hleft)) g_i = i;
				gscore = lc_max(gscore, hleft);
			}
			if (rowmax == 0) stop = true;
			else if (rowmax > best) {
				best = rowmax; best_i = i; best_j = rowarg;
				int off = rowarg - i; if (off < 0) off = -off;
				max_off = lc_max(max_off, off);
			} else if (zdrop > 0) {
				const int di = i - best_i, dj = rowarg - best_j;
				if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) stop = true; }
				else { if (best - rowmax - (dj - di) * e_ins > zdrop) stop = true; }
			}
			if (!stop) {
				// shrink the band where H and E are both zero (B/ksw.c:466-469): the two scans of the reference, from the columns
				// recorded while the row was written (he[end] = {hleft, 0} is the only entry outside [beg, end))
				const int nb = nz_first < end ? nz_first : end;
				const int t = hleft ? end : (nz_last >= nb ? nz_last : nb - 1);
				beg = nb;
				end = lc_min(t + 2, qlen);
				if (++i >= tlen) stop = true;
			}
			if (stop) dp_finish();
			else { fpos += fdir; row_start(); phase = LC_XP_DP; }
		}
		LC_CONVERGE();
	}

	// ---- mem_chain2aln between the extensions; leaves phase >= LC_XP_DP or LC_XP_FETCH
	LC_HD void resolve()
	{
		const lc_opt_t &opt = b.opt;
		const int64_t l_pac = b.ix.l_pac;
		while (phase < LC_XP_DP && phase != LC_XP_FETCH) {
			if (phase == LC_XP_AFTER_LEFT) {                      // B/bwamem.c:726-744
				a_score = best;
				if (a_score == prev || max_off < (aw0 >> 1) + (aw0 >> 2) || tryi == 1) {
					if (gscore <= 0 || gscore <= a_score - opt.pen_clip5) { a_qb = s_qbeg - (best_j + 1); a_rb = s_rbeg - (best_i + 1); a_truesc = a_score; }
					else { a_qb = 0; a_rb = s_rbeg - (g_i + 1); a_truesc = gscore; }
					phase = LC_XP_RIGHT_CHECK;
				} else { tryi = 1; phase = LC_XP_LEFT_SETUP; }
			}
			if (phase == LC_XP_AFTER_RIGHT) {                     // B/bwamem.c:756-771
				a_score = best;
				if (a_score == prev || max_off < (aw1 >> 1) + (aw1 >> 2) || tryi == 1) {
					if (gscore <= 0 || gscore <= a_score - opt.pen_clip3) { a_qe = qe0 + (best_j + 1); a_re = re0 + (best_i + 1); a_truesc += a_score - sc0; }
					else { a_qe = l_query; a_re = re0 + (g_i + 1); a_truesc += gscore - sc0; }
					phase = LC_XP_FINALIZE;
				} else { tryi = 1; phase = LC_XP_RIGHT_SETUP; }
			}
			if (phase == LC_XP_RIGHT_CHECK) {                     // B/bwamem.c:746-751,771
				if (s_qbeg + s_len != l_query) { qe0 = s_qbeg + s_len; sc0 = a_score; re0 = s_rbeg + s_len; tryi = 0; phase = LC_XP_RIGHT_SETUP; }
				else { a_qe = l_query; a_re = s_rbeg + s_len; phase = LC_XP_FINALIZE; }
			}
			if (phase == LC_XP_FINALIZE) {                        // B/bwamem.c:773-783
				lc_alnreg_t a;
				memset(&a, 0, sizeof(a));
				a.rid = c_rid; a.qb = a_qb; a.qe = a_qe; a.rb = a_rb; a.re = a_re; a.score = a_score; a.truesc = a_truesc;
				a.seedcov = 0;
				for (int t = 0; t < cn; ++t) {
					const LcSeed &sd = seeds[t];
					if (sd.qbeg >= a.qb && sd.qbeg + sd.len <= a.qe && sd.rbeg >= a.rb && sd.rbeg + sd.len <= a.re) a.seedcov += sd.len;
				}
				a.w = lc_max(aw0, aw1);
				a.seedlen0 = s_len;
				a.frac_rep = b.r_frac_rep[r];
				av[n_av++] = a;
				--k;
				phase = LC_XP_SEED_NEXT;
			}
			if (phase == LC_XP_READ_START) {
				l_query = (int)(b.off[r + 1] - b.off[r]);
				NC = b.r_n_chain[r];
				const int T = b.r_n_cseed[r];
				if (!UNIT) { b.r_n_reg[r] = 0; b.r_reg_off[r] = 0; }
				n_av = 0; ci = 0; ro = 0;
				phase = LC_XP_CHAIN_START;
				if (UNIT) {
					const int TU = b.chain_aux[cg].unit_seeds;
					n_av0 = 0;
					ro = lc_carve_plain(b, b.reg_cur, b.reg_cap, (unsigned long long)TU, LC_OVF_REG);
					const unsigned long long so = lc_carve_plain(b, b.scratch_cur, b.scratch_cap, 8ull * TU + 8, LC_OVF_SCRATCH);
					if (ro == (unsigned long long)-1 || so == (unsigned long long)-1) phase = LC_XP_FETCH;
					else { av = b.reg + ro; srt = (uint64_t *)lc_align8((size_t)(b.scratch + so)); }
				}
				else if (NC == 0) phase = LC_XP_FETCH;
				else if ((b.ext_heavy_seeds > 0 && T > b.ext_heavy_seeds) || (b.ext_heavy_chains > 0 && NC > b.ext_heavy_chains)) {
					// a read with very many chains (repeats) would keep this lane — and with it the kernel — busy long after the others are
					// done: it goes to a list that a warp-per-read kernel works off (lc_stage_extend_t, cooperative rows)
					b.wl_in[lc_atomic_add(b.ext_heavy_n, 1u)] = (uint64_t)(uint32_t)lc_min(T, 0xffff) << 32 | (uint32_t)r;
					phase = LC_XP_FETCH;
				} else {
					ro = lc_carve_plain(b, b.reg_cur, b.reg_cap, (unsigned long long)T, LC_OVF_REG);
					const unsigned long long so = lc_carve_plain(b, b.scratch_cur, b.scratch_cap, 8ull * T + 8, LC_OVF_SCRATCH);
					if (ro == (unsigned long long)-1 || so == (unsigned long long)-1) phase = LC_XP_FETCH;
					else { av = b.reg + ro; srt = (uint64_t *)lc_align8((size_t)(b.scratch + so)); }
				}
			}
			if (phase == LC_XP_CHAIN_START) {                     // B/bwamem.c:642-665
				if (UNIT ? cg < 0 : ci >= NC) phase = LC_XP_READ_DONE;
				else {
					const LcChainOut &c = b.chain[UNIT ? cg : (long long)b.r_chain_off[r] + ci];
					seeds = b.cseed + c.seed_off; cn = c.n; c_rid = c.rid;
					if (cn == 0) ++ci;                            // (stays in CHAIN_START: another sweep)
					else {
						int64_t x0 = l_pac << 1, x1 = 0;
						for (int t = 0; t < cn; ++t) {            // widest window any seed may extend into
							const LcSeed &sd = seeds[t];
							const int64_t bb = sd.rbeg - (sd.qbeg + lc_cal_max_gap(opt, sd.qbeg));
							const int tail = l_query - sd.qbeg - sd.len;
							const int64_t ee = sd.rbeg + sd.len + (tail + lc_cal_max_gap(opt, tail));
							x0 = lc_min(x0, bb); x1 = lc_max(x1, ee);
						}
						x0 = lc_max(x0, (int64_t)0); x1 = lc_min(x1, l_pac << 1);
						if (x0 < l_pac && l_pac < x1) { if (seeds[0].rbeg < l_pac) x1 = l_pac; else x0 = l_pac; }
						lc_fetch_clamp(b.ix, x0, seeds[0].rbeg, x1);    // bns_fetch_seq clamps to the contig (B/bwamem.c:659)
						r0 = x0; r1 = x1;
						for (int t = 0; t < cn; ++t) srt[t] = (uint64_t)(uint32_t)seeds[t].len << 32 | (uint32_t)t;   // score == len (B/bwamem.c:663-665)
						lc_introsort(srt, (long)cn, [](uint64_t x, uint64_t y) { return x < y; });
						k = cn - 1;
						phase = LC_XP_SEED_NEXT;
					}
				}
			}
			if (phase == LC_XP_SEED_NEXT) {                       // B/bwamem.c:667-713
				if (k < 0) {
					if (UNIT) {
						LcChainAux &ax = b.chain_aux[cg];
						ax.reg_off = (int64_t)(ro + (unsigned long long)n_av0); ax.n_reg = n_av - n_av0;
						n_av0 = n_av; cg = ax.next;
					} else ++ci;
					phase = LC_XP_CHAIN_START;
				}
				else {
					const LcSeed s = seeds[(uint32_t)srt[k]];
					int t;
					for (t = 0; t < n_av; ++t) {                  // seed already explained by an earlier region? (B/bwamem.c:671-687)
						const lc_alnreg_t &p = av[t];
						int64_t rd; int qd, ww, g;
						if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
						if (s.len - p.seedlen0 > .1 * l_query) continue;
						qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
						g = lc_cal_max_gap(opt, qd < rd ? qd : (int)rd); ww = lc_min(g, p.w);
						if (qd - rd < ww && rd - qd < ww) break;
						qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
						g = lc_cal_max_gap(opt, qd < rd ? qd : (int)rd); ww = lc_min(g, p.w);
						if (qd - rd < ww && rd - qd < ww) break;
					}
					bool skip = false;
					if (t < n_av) {                               // contained: extend only if an overlapping off-diagonal seed exists (B/bwamem.c:688-706)
						for (t = k + 1; t < cn; ++t) {
							if (srt[t] == 0) continue;
							const LcSeed &o = seeds[(uint32_t)srt[t]];
							if (o.len < s.len * .95) continue;
							if (s.qbeg <= o.qbeg && s.qbeg + s.len - o.qbeg >= s.len >> 2 && o.qbeg - s.qbeg != o.rbeg - s.rbeg) break;
							if (o.qbeg <= s.qbeg && o.qbeg + o.len - s.qbeg >= s.len >> 2 && s.qbeg - o.qbeg != s.rbeg - o.rbeg) break;
						}
						if (t == cn) { srt[k] = 0; --k; skip = true; }     // (stays in SEED_NEXT: another sweep)
					}
					if (!skip) {
						s_rbeg = s.rbeg; s_qbeg = s.qbeg; s_len = s.len;
						aw0 = aw1 = opt.w; a_score = a_truesc = -1;
						if (s.qbeg) { tryi = 0; phase = LC_XP_LEFT_SETUP; }
						else { a_score = a_truesc = s.len * opt.a; a_qb = 0; a_rb = s.rbeg; phase = LC_XP_RIGHT_CHECK; }
					}
				}
			}
			if (phase == LC_XP_LEFT_SETUP) {                      // left extension on reversed sequences (B/bwamem.c:715-725)
				prev = a_score; aw0 = opt.w << tryi;
				dp_setup(s_qbeg, l_query - s_qbeg, -1, (int)(s_rbeg - r0), s_rbeg - 1, -1, aw0, opt.pen_clip5, s_len * opt.a);     // query[qbeg-1-c] = reversed[l_query-qbeg+c]
			}
			if (phase == LC_XP_RIGHT_SETUP) {                     // right extension (B/bwamem.c:752-755)
				prev = a_score; aw1 = opt.w << tryi;
				dp_setup(l_query - qe0, qe0, 1, (int)(r1 - re0), re0, 1, aw1, opt.pen_clip3, sc0);
			}
			if (phase == LC_XP_READ_DONE) {
				if (!UNIT) { b.r_reg_off[r] = (int64_t)ro; b.r_n_reg[r] = n_av; }
				else if (n_av) lc_atomic_add(b.r_n_reg + r, (int32_t)n_av);
				if (n_av) lc_atomic_add(b.reg_total, (unsigned long long)n_av);
				phase = LC_XP_FETCH;
			}
		}
	}
};

// words of the read copy (one spare word in front: a reversed window may start up to 8 bases before the read) / columns of the
// H/E row for reads up to lmax (the tick reads whole groups, hence the slack)
LC_HD int lc_ext_q_words(int lmax) { return (lmax + 7) / 8 + 3; }
LC_HD int lc_ext_he_cols(int lmax) { return ((lmax + 2 + LC_XT_CELLS - 1) & ~(LC_XT_CELLS - 1)) + LC_XT_CELLS; }

// nibble-pack read r into qw[w * STRIDE] (scalar form; the device kernel uses the whole warp)
template <int STRIDE> LC_HD void lc_ext_load_read(const LcBatch &b, int r, uint32_t *qw, int n_words)
{
	const int len = (int)(b.off[r + 1] - b.off[r]);
	const uint8_t *src = b.seq + b.off[r];
	for (int w = 0; w < n_words; ++w) {
		uint32_t v = 0;
		for (int t = 0; t < 8; ++t) {
			const int p = 8 * w + t;
			if (p < len) v |= (uint32_t)(src[p] & 15) << (4 * t);
		}
		qw[w * STRIDE] = v;
	}
}

// scalar driver (tests/hostsim): all reads, one after another, through the same state machine; cells8 selects the byte packing
template <class CELLS, bool UNIT> static inline void lc_ext_run_scalar_t(const LcBatch &b, uint32_t *he, uint32_t *qw, int idle_ticks = 0)
{
	uint32_t mtab[36];
	lc_ext_fill_masks(mtab);
	LcExtLane<1, CELLS, UNIT> ex(b, he, qw + 1, mtab);
	int next = 0;
	const int n_items = UNIT ? lc_ext_n_units(b) : b.n_reads;
	for (;;) {
		while (ex.phase == LC_XP_FETCH) {
			if (next >= n_items) { ex.phase = LC_XP_DONE; break; }
			const int it = next++;
			const int r = UNIT ? b.chain_aux[lc_ext_unit_at(b, it)].r : it;
			lc_ext_load_read<1>(b, r, qw + 1, lc_ext_q_words(b.lmax) - 1);
			if (UNIT) ex.begin_unit((long long)lc_ext_unit_at(b, it)); else ex.begin_read(r);
			ex.resolve();
		}
		if (ex.phase == LC_XP_DONE) break;
		if (idle_ticks == 0) ex.tick();
		else {                                  // the device kernel's gates: end-of-row work and resolve() wait for the warp's next round
			ex.tick(false);
			if (ex.phase == LC_XP_ROWEND) { for (int t = 0; t < idle_ticks; ++t) ex.tick(false); ex.tick(true); }
			if (ex.phase < LC_XP_DP) for (int t = 0; t < idle_ticks; ++t) ex.tick(false);
		}
		ex.resolve();
	}
}
LC_HD bool lc_ext_cells8_ok(const LcBatch &b) { return b.lmax * b.opt.a <= 255 && b.opt.a > 0; }
static inline void lc_ext_run_scalar(const LcBatch &b, uint32_t *he, uint32_t *qw, bool cells8, int idle_ticks = 0, bool units = false)
{
	if (units) { if (cells8) lc_ext_run_scalar_t<LcCells8<1>, true>(b, he, qw, idle_ticks); else lc_ext_run_scalar_t<LcCells16<1>, true>(b, he, qw, idle_ticks); }
	else if (cells8) lc_ext_run_scalar_t<LcCells8<1>, false>(b, he, qw, idle_ticks); else lc_ext_run_scalar_t<LcCells16<1>, false>(b, he, qw, idle_ticks);
}
