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

// ---- the H/E row in (shared) memory.  Two packings, chosen by the longest read of the batch:
//   LcCells16: one column per 32-bit word, h | e << 16                        (any read length up to LC_MAX_READ_LEN)
//   LcCells8 : two columns per word, (h | e << 8) << 16 * (column & 1)        (lmax * a <= 255: no cell of ksw_extend2 can exceed
//              the score of the whole read matching, so h and e fit a byte) — half the shared memory, twice the resident warps
template <int STRIDE> struct LcCells16 {
	static LC_HD int words(int cols) { return cols; }
	static LC_HD void load8(const uint32_t *he, int col0, int *h, int *e)
	{
#pragma unroll
		for (int c = 0; c < 8; ++c) { const uint32_t v = he[(col0 + c) * STRIDE]; h[c] = (int)(v & 0xffffu); e[c] = (int)(v >> 16); }
	}
	static LC_HD void store8(uint32_t *he, int col0, const int *h, const int *e)
	{
#pragma unroll
		for (int c = 0; c < 8; ++c) he[(col0 + c) * STRIDE] = (uint32_t)h[c] | (uint32_t)e[c] << 16;
	}
	static LC_HD void put(uint32_t *he, int col, int h, int e) { he[col * STRIDE] = (uint32_t)h | (uint32_t)e << 16; }
};
template <int STRIDE> struct LcCells8 {
	static LC_HD int words(int cols) { return cols >> 1; }
	static LC_HD void load8(const uint32_t *he, int col0, int *h, int *e)      // col0 is a multiple of 8
	{
#pragma unroll
		for (int c = 0; c < 8; c += 2) {
			const uint32_t v = he[((col0 + c) >> 1) * STRIDE];
			h[c] = (int)(v & 0xffu); e[c] = (int)(v >> 8 & 0xffu); h[c + 1] = (int)(v >> 16 & 0xffu); e[c + 1] = (int)(v >> 24);
		}
	}
	static LC_HD void store8(uint32_t *he, int col0, const int *h, const int *e)
	{
#pragma unroll
		for (int c = 0; c < 8; c += 2)
			he[((col0 + c) >> 1) * STRIDE] = (uint32_t)h[c] | (uint32_t)e[c] << 8 | (uint32_t)h[c + 1] << 16 | (uint32_t)e[c + 1] << 24;
	}
	static LC_HD void put(uint32_t *he, int col, int h, int e)
	{
		uint32_t *p = he + (col >> 1) * STRIDE;
		const int sh = (col & 1) << 4;
		*p = (*p & ~(0xffffu << sh)) | ((uint32_t)h | (uint32_t)e << 8) << sh;
	}
};

enum { LC_XP_FETCH = 0, LC_XP_AFTER_LEFT, LC_XP_AFTER_RIGHT, LC_XP_RIGHT_CHECK, LC_XP_FINALIZE, LC_XP_READ_START, LC_XP_CHAIN_START, LC_XP_SEED_NEXT,
	   LC_XP_LEFT_SETUP, LC_XP_RIGHT_SETUP, LC_XP_READ_DONE,
	   LC_XP_DP, LC_XP_DONE };          // states >= LC_XP_DP: a DP row is in progress (or the lane is finished)

template <int STRIDE, class CELLS = LcCells16<STRIDE> > struct LcExtLane {
	const LcBatch &b;
	uint32_t *he;            // H/E row of lmax + 2 columns, packed by CELLS
	uint32_t *qw;            // the read: qw[w * STRIDE] holds bases 8w..8w+7, 4 bits each, first base in the low nibble; qw[-STRIDE] is
	                         //   addressable.  Left extensions run on reversed sequences (B/bwamem.c:715-721): their columns are read from the
	                         //   same copy back to front
	int phase;
	// ---- read-level state (mem_chain2aln's locals)
	int r, l_query, NC, ci, n_av, k, cn, tryi, prev;
	unsigned long long ro;
	lc_alnreg_t *av;
	uint64_t *srt;
	const LcSeed *seeds;
	int64_t r0, r1, s_rbeg, a_rb, a_re, re0;
	int s_qbeg, s_len, c_rid, a_qb, a_qe, a_score, a_truesc, aw0, aw1, qe0, sc0;
	// ---- the extension in progress (ksw_extend2's locals)
	int qlen, tlen, w, h0, zdrop, i, j, beg, end, f, hleft, rowmax, rowarg, best, best_i, best_j, g_i, gscore, max_off, tb;
	int nz_first, nz_last;   // first / last column of the row whose updated {h, e} word is non-zero (the band shrink of B/ksw.c:466-469)
	int64_t fpos; int fdir, comp;          // forward-strand coordinate of the current row's target base, its step per row, 3 on the reverse strand
	uint32_t pw, pwn; int64_t pwi;         // the 16 packed reference bases around fpos, the next such word in walking direction (prefetched), word index
	int q0, qdir;            // the extension's query column c is base q0 + c of the forward (qdir > 0) or reversed (qdir < 0) read copy
	unsigned long long cells;

	LC_HD LcExtLane(const LcBatch &b_, uint32_t *he_, uint32_t *qw_) : b(b_), he(he_), qw(qw_), phase(LC_XP_FETCH), l_query(0), j(0), end(0), q0(0), qdir(1), tb(0) {}

	LC_HD uint32_t pac_word(int64_t wi) const { return lc_ldg((const uint32_t *)b.ix.pac + (wi < 0 ? 0 : wi)); }
	LC_HD int target_base()              // base of the current row: bns_get_seq's decoding (B/bntseq.c:398-419) out of the cached word
	{
		if ((fpos >> 4) != pwi) { pw = pwn; pwi += fdir; pwn = pac_word(pwi + fdir); }
		return (int)((pw >> ((((int)(fpos >> 2) & 3) << 3) + ((~(int)fpos & 3) << 1))) & 3) ^ comp;
	}
	LC_HD void begin_read(int r_) { r = r_; phase = LC_XP_READ_START; }

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
			int hh[LC_XT_CELLS], ee[LC_XT_CELLS];
			for (int t0 = 0; t0 <= qlen + 1; t0 += LC_XT_CELLS) {
#pragma unroll
				for (int c = 0; c < LC_XT_CELLS; ++c) {
					ee[c] = 0;
					if (t0 + c == 0) hh[c] = h0;
					else { hh[c] = v; v = v > sc.e_ins ? v - sc.e_ins : 0; }
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
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		hleft = 0;
		if (beg == 0) { hleft = h0 - (sc.o_del + sc.e_del * (i + 1)); if (hleft < 0) hleft = 0; }
		f = 0; rowmax = 0; rowarg = -1; j = beg & ~(LC_XT_CELLS - 1);      // ticks are aligned to groups of LC_XT_CELLS columns
		nz_first = 1 << 30; nz_last = -1;
		tb = target_base();
	}
	LC_HD void dp_finish()                                    // hands {best, qle, tle, gtle, gscore, max_off} to the caller's state
	{
		if (b.cnt) { lc_atomic_add(&b.cnt->ext_calls, 1ull); lc_atomic_add(&b.cnt->ext_cells, cells); }
		phase = qdir < 0 ? LC_XP_AFTER_LEFT : LC_XP_AFTER_RIGHT;
	}

	// ---- one tick: the LC_XT_CELLS columns j..j+7 of the current row (B/ksw.c:431-456), then, if the row is complete, the
	// end-of-row work (B/ksw.c:457-478).  The cell loop is branch-free: everything that does not depend on the running F/H
	// chain is computed for the eight columns first, the chain itself is a few instructions per cell, and cells outside
	// [beg, end) are computed but not committed.
	LC_HD void tick()
	{
		const LcScore &sc = b.sc;
		const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_del = sc.e_del, e_ins = sc.e_ins;
		const bool run = phase == LC_XP_DP;
		const int col0 = j;
		{
			// query codes of the eight columns: one funnel shift over two words of the read copy; a left extension reads the read
			// back to front (column c of the reversed read = base l_query - 1 - q0 - c), so its window is nibble-reversed
			uint32_t qword;
			{
				const int qi = qdir > 0 ? q0 + col0 : l_query - LC_XT_CELLS - q0 - col0;      // first base of the window, >= -8
				const uint32_t w0 = qw[(qi >> 3) * STRIDE], w1 = qw[((qi >> 3) + 1) * STRIDE];
				uint32_t v = (uint32_t)(((uint64_t)w1 << 32 | w0) >> ((qi & 7) << 2));
				if (qdir < 0) {
#ifdef __CUDA_ARCH__
					v = __byte_perm(v, 0, 0x0123);
#else
					v = v >> 24 | (v >> 8 & 0xff00u) | (v << 8 & 0xff0000u) | v << 24;
#endif
					v = (v & 0x0f0f0f0fu) << 4 | (v >> 4 & 0x0f0f0f0fu);
				}
				qword = v;
			}
			// substitution score without compares per base: x has a zero nibble where the base equals the target base
			const uint32_t x = qword ^ (tb < 4 ? (uint32_t)tb * 0x11111111u : 0x88888888u);
			const int pen = tb < 4 ? -sc.b : -1;
			int hd[LC_XT_CELLS], ev[LC_XT_CELLS];
			int M[LC_XT_CELLS], e2[LC_XT_CELLS], ti[LC_XT_CELLS];
			CELLS::load8(he, col0, hd, ev);
#pragma unroll
			for (int c = 0; c < LC_XT_CELLS; ++c) {
				const int diag = hd[c], e = ev[c];
				const int s = (x & (0xfu << (4 * c))) == 0 ? sc.a : ((qword & (0x4u << (4 * c))) ? -1 : pen);
				M[c] = diag ? diag + s : 0;
				e2[c] = lc_max(e - e_del, lc_max(M[c] - oe_del, 0));
				ti[c] = lc_max(M[c] - oe_ins, 0);
			}
#pragma unroll
			for (int c = 0; c < LC_XT_CELLS; ++c) {
				const int col = col0 + c;
				const bool act = run && col >= beg && col < end;
				const int h = lc_max(lc_max(M[c], ev[c]), f);
				if (act) {
					hd[c] = hleft; ev[c] = e2[c];                      // H[j] = h of the column to the left, E[j] for the next row
					if (hleft | e2[c]) { nz_first = lc_min(nz_first, col); nz_last = col; }
					if (!(rowmax > h)) rowarg = col;
					rowmax = lc_max(rowmax, h);
					f = lc_max(f - e_ins, ti[c]);
					hleft = h;
				}
			}
			if (run) CELLS::store8(he, col0, hd, ev);
			if (run) j = col0 + LC_XT_CELLS;
		}
		LC_CONVERGE();
		if (run && j >= end) {                                 // the row is complete
			bool stop = false;
			const int jfin = beg < end ? end : beg;            // the reference's loop variable after the column loop
			if (b.cnt) cells += (unsigned long long)(end > beg ? end - beg : 0);
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
			else { fpos += fdir; row_start(); }
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
				b.r_n_reg[r] = 0; b.r_reg_off[r] = 0;
				n_av = 0; ci = 0; ro = 0;
				phase = LC_XP_CHAIN_START;
				if (NC == 0) phase = LC_XP_FETCH;
				else {
					ro = lc_carve(b, b.reg_cur, b.reg_cap, (unsigned long long)T, LC_OVF_REG);
					const unsigned long long so = lc_carve(b, b.scratch_cur, b.scratch_cap, 8ull * T + 8, LC_OVF_SCRATCH);
					if (ro == (unsigned long long)-1 || so == (unsigned long long)-1) phase = LC_XP_FETCH;
					else { av = b.reg + ro; srt = (uint64_t *)lc_align8((size_t)(b.scratch + so)); }
				}
			}
			if (phase == LC_XP_CHAIN_START) {                     // B/bwamem.c:642-665
				if (ci >= NC) phase = LC_XP_READ_DONE;
				else {
					const LcChainOut &c = b.chain[b.r_chain_off[r] + ci];
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
				if (k < 0) { ++ci; phase = LC_XP_CHAIN_START; }
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
				b.r_reg_off[r] = (int64_t)ro; b.r_n_reg[r] = n_av;
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
template <class CELLS> static inline void lc_ext_run_scalar_t(const LcBatch &b, uint32_t *he, uint32_t *qw)
{
	LcExtLane<1, CELLS> ex(b, he, qw + 1);
	int next = 0;
	for (;;) {
		while (ex.phase == LC_XP_FETCH) {
			if (next >= b.n_reads) { ex.phase = LC_XP_DONE; break; }
			const int r = next++;
			lc_ext_load_read<1>(b, r, qw + 1, lc_ext_q_words(b.lmax) - 1);
			ex.begin_read(r);
			ex.resolve();
		}
		if (ex.phase == LC_XP_DONE) break;
		ex.tick();
		ex.resolve();
	}
}
LC_HD bool lc_ext_cells8_ok(const LcBatch &b) { return b.lmax * b.opt.a <= 255 && b.opt.a > 0; }
static inline void lc_ext_run_scalar(const LcBatch &b, uint32_t *he, uint32_t *qw, bool cells8)
{
	if (cells8) lc_ext_run_scalar_t<LcCells8<1> >(b, he, qw); else lc_ext_run_scalar_t<LcCells16<1> >(b, he, qw);
}
