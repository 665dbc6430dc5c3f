// lc_seed.cuh — seeding (mem_collect_intv, B/bwamem.c:114-162) as a one-extension-per-tick state machine.
//
// Why a state machine: a read's seeding is ~700 *dependent* bwt_extend calls (B/bwt.c:262-275), each a pair of random
// 32-byte FM-index lookups.  Written as the reference's nested loops with one thread per read, the 32 reads of a warp
// drift into different loop nests and the warp issues every instruction for ~4 live lanes (ncu:
// profiles/r01a_smem_kernel_details.txt).  Here every lane still owns one read, but the whole warp executes the same
// short tick — "do the pending extension (two 256-bit block loads + popcounts), then advance this read's state" — so the
// expensive part is always converged.  The three passes of mem_collect_intv, the forward/backward phases of bwt_smem1a
// (B/bwt.c:289-351) and bwt_seed_strategy1 (B/bwt.c:358-379) are states; the interval lists of bwt_smem1a live in shared
// memory as three words per interval ({x0, x2, end}: x1 is only needed while the match grows forward and stays in registers),
// laid out [word][lane] and compacted in place.  Frequent transitions are handled inline in the two `after_*` paths;
// everything rare goes through resolve(), a single ordered sweep over the bookkeeping states, which the kernel runs only
// every few ticks for all lanes that wait for it (lc_kernels.cu).
// A first design with 4 lanes per read (two per lookup) was measured at 469 warp-instructions per 8 extensions
// (profiles/r01b_*): replicated bookkeeping, not the lookups, was the cost — hence one lane per read.
//
// Output of the state machine: the read's seeds, UNSORTED, as packed entries in a pool.  lc_seed_finish (a second, small
// kernel: 4 lanes per read) sorts them by `info`, lays out the occurrence schedule of mem_chain and fetches positions.
//
// The code is plain scalar C++ (LC_HD): tests/hostsim runs the very same state machine on the CPU against the oracle.
// U = uint32_t when the index has < 2^32 BWT symbols (intervals fit 32 bits), uint64_t otherwise (< 2^34).
//
// Capacity: reads whose interval list outgrows the shared-memory slot (`lst_cap`) or whose seed list outgrows `pk_cap`
// are appended to a redo list; the host runs the state machine again over that list with larger capacities (up to three
// passes, lc_pipeline.hpp), and what still does not fit goes to the one-thread-per-read kernel (lc_stage_seed), which
// has no such limits.
#pragma once
#include "lc_mem.cuh"

#ifdef __CUDA_ARCH__
#define LC_CONVERGE() __syncwarp()
#else
#define LC_CONVERGE() ((void)0)
#endif

enum { LC_SP_FETCH = 0, LC_SP_OVERFLOW, LC_SP_FINISH, LC_SP_END_SMEM, LC_SP_P1_NEXT, LC_SP_NEXT_RESEED, LC_SP_START_SMEM, LC_SP_FWD_PREP,
	   LC_SP_START_BWD, LC_SP_BWD_PREP, LC_SP_P3_NEXT, LC_SP_P3_PREP,
	   LC_SP_FWD, LC_SP_BWD, LC_SP_P3, LC_SP_DONE };      // states >= LC_SP_FWD wait for an extension (or are final)

// ---- packed bi-interval + two read positions in 16 bytes
template <typename U> struct LcPkOps;
template <> struct LcPkOps<uint32_t> {
	static LC_HD LcPk pack(uint32_t x0, uint32_t x1, uint32_t x2, int e1, int e2) { LcPk p; p.a = x0; p.b = x1; p.c = x2; p.m = (uint32_t)e1 | (uint32_t)e2 << 16; return p; }
	static LC_HD uint32_t x0(const LcPk &p) { return p.a; }
	static LC_HD uint32_t x1(const LcPk &p) { return p.b; }
	static LC_HD uint32_t x2(const LcPk &p) { return p.c; }
	static LC_HD int e1(const LcPk &p) { return (int)(p.m & 0xffff); }
	static LC_HD int e2(const LcPk &p) { return (int)(p.m >> 16); }
};
// LC_PK_LOW_BITS: how many bits of a 64-bit interval value stay in its 32-bit word; the rest (2 bits) go to the shared word.  32 in the
// product; tests/hostsim builds a second binary with 18 so that the high fields are non-zero on a small index (same code path).
#ifndef LC_PK_LOW_BITS
#define LC_PK_LOW_BITS 32
#endif
#define LC_PK_LOW_MASK ((uint32_t)((1ull << LC_PK_LOW_BITS) - 1))
template <> struct LcPkOps<uint64_t> {      // x < 2^34, positions < 4096
	static LC_HD LcPk pack(uint64_t x0, uint64_t x1, uint64_t x2, int e1, int e2)
	{
		LcPk p;
		p.a = (uint32_t)x0 & LC_PK_LOW_MASK; p.b = (uint32_t)x1 & LC_PK_LOW_MASK; p.c = (uint32_t)x2 & LC_PK_LOW_MASK;
		p.m = (uint32_t)(x0 >> LC_PK_LOW_BITS) | (uint32_t)(x1 >> LC_PK_LOW_BITS) << 2 | (uint32_t)(x2 >> LC_PK_LOW_BITS) << 4 | (uint32_t)e1 << 8 | (uint32_t)e2 << 20;
		return p;
	}
	static LC_HD uint64_t x0(const LcPk &p) { return (uint64_t)(p.m & 3) << LC_PK_LOW_BITS | p.a; }
	static LC_HD uint64_t x1(const LcPk &p) { return (uint64_t)((p.m >> 2) & 3) << LC_PK_LOW_BITS | p.b; }
	static LC_HD uint64_t x2(const LcPk &p) { return (uint64_t)((p.m >> 4) & 3) << LC_PK_LOW_BITS | p.c; }
	static LC_HD int e1(const LcPk &p) { return (int)((p.m >> 8) & 0xfff); }
	static LC_HD int e2(const LcPk &p) { return (int)(p.m >> 20); }
};

// ---- one 32-byte block of the device FM-index, fetched with a single 256-bit load on the device.
// L1::no_allocate matters: the kernel's shared memory leaves almost no L1, and every in-flight miss of an allocating load holds an
// L1 line — measured with scripts/microbench/randsector.cu (profiles/r01_randsector.txt): random 32-byte reads from a 100 MB table
// run at 44 G sectors/s with plain ld.global.nc and at 137 G sectors/s with the no-allocate form (carve-out at 100 %).
struct LcBlk { uint32_t cnt[4]; uint64_t hi, lo; };
LC_HD LcBlk lc_ld_blk(const uint4 *p)
{
	LcBlk v;
#ifdef __CUDA_ARCH__
	uint64_t a, b;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(v.hi), "=l"(v.lo) : "l"(p));
	v.cnt[0] = (uint32_t)a; v.cnt[1] = (uint32_t)(a >> 32); v.cnt[2] = (uint32_t)b; v.cnt[3] = (uint32_t)(b >> 32);
#else
	v.cnt[0] = p[0].x; v.cnt[1] = p[0].y; v.cnt[2] = p[0].z; v.cnt[3] = p[0].w;
	v.hi = (uint64_t)p[1].y << 32 | p[1].x; v.lo = (uint64_t)p[1].w << 32 | p[1].z;
#endif
	return v;
}
// a = occ(row, c), s = sum over bases above c of occ(row, b); `r` = symbols of the block that count, ch/cl = c's bit planes
// (general form: 64-bit intervals with superblock bases)
template <typename U> LC_HD void lc_blk_counts(const LcIndex &ix, const LcBlk &v, uint64_t blk, int r, int c, uint64_t ch, uint64_t cl, U &a, U &s)
{
	const uint64_t m = lc_keep_first(r);
	const uint64_t eq = ~(v.hi ^ ch) & ~(v.lo ^ cl) & m;
	const uint64_t gt = ((v.hi & ~ch) | (~(v.hi ^ ch) & v.lo & ~cl)) & m;
	a = (U)(c == 0 ? v.cnt[0] : c == 1 ? v.cnt[1] : c == 2 ? v.cnt[2] : v.cnt[3]) + (U)lc_popcll(eq);
	s = (U)(c < 3 ? v.cnt[3] : 0u) + (U)(c < 2 ? v.cnt[2] : 0u) + (U)(c < 1 ? v.cnt[1] : 0u) + (U)lc_popcll(gt);
	if (sizeof(U) == 8) {      // superblock bases (indexes of 2^32 symbols or more)
		const uint64_t *sb = ix.sb[blk >> LC_SB_SHIFT];
		a += (U)sb[c];
		s += (U)((c < 3 ? sb[3] : 0ull) + (c < 2 ? sb[2] : 0ull) + (c < 1 ? sb[1] : 0ull));
	}
}
// x >> s with PTX semantics: a shift of 32 or more gives 0
LC_HD uint32_t lc_shr_clamp(uint32_t x, uint32_t s)
{
#ifdef __CUDA_ARCH__
	uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s)); return r;
#else
	return s >= 32 ? 0u : x >> s;
#endif
}
// among the first r symbols (1..64) of a block: e = how many equal c, g = how many are greater; chw/clw = c's two bits as word masks.
// Written on 32-bit halves (the planes keep their first symbol in bit 63) with excluded-symbol masks: ten LOP3, four POPC.
LC_HD void lc_blk_popc(const LcBlk &v, int r, uint32_t chw, uint32_t clw, uint32_t &e, uint32_t &g)
{
	const uint32_t xt = lc_shr_clamp(0xffffffffu, (uint32_t)r), xb = lc_shr_clamp(0xffffffffu, (uint32_t)(r > 32 ? r - 32 : 0));   // 1 = symbol does not count
	const uint32_t Ht = (uint32_t)(v.hi >> 32), Hb = (uint32_t)v.hi, Lt = (uint32_t)(v.lo >> 32), Lb = (uint32_t)v.lo;
	const uint32_t t1 = ~(Ht ^ chw) & ~xt, b1 = ~(Hb ^ chw) & ~xb;                  // high bit equal
	const uint32_t eqt = t1 & ~(Lt ^ clw), eqb = b1 & ~(Lb ^ clw);
	const uint32_t gtt = (Ht & ~chw & ~xt) | (t1 & Lt & ~clw), gtb = (Hb & ~chw & ~xb) | (b1 & Lb & ~clw);
	e = (uint32_t)(lc_popc(eqt) + lc_popc(eqb));
	g = (uint32_t)(lc_popc(gtt) + lc_popc(gtb));
}
// 32-bit intervals (one superblock): ak = occ(k, c), o2 = occ(l, c) - occ(k, c), sd = sum over bases above c of occ(l, b) - occ(k, b).
// The block counts enter only as differences (except ak), so the base selection is done once on the difference vector.
LC_HD void lc_blk_pair32(const LcBlk &vk, const LcBlk &vl, int rk, int rl, int c, uint32_t &ak, uint32_t &o2, uint32_t &sd)
{
	const bool c1 = (c & 1) != 0, c2 = (c & 2) != 0;
	const uint32_t chw = c2 ? ~0u : 0u, clw = c1 ? ~0u : 0u;
	uint32_t ek, gk, el, gl;
	lc_blk_popc(vk, rk, chw, clw, ek, gk);
	lc_blk_popc(vl, rl, chw, clw, el, gl);
	const uint32_t d0 = vl.cnt[0] - vk.cnt[0], d1 = vl.cnt[1] - vk.cnt[1], d2 = vl.cnt[2] - vk.cnt[2], d3 = vl.cnt[3] - vk.cnt[3];
	const uint32_t k01 = c1 ? vk.cnt[1] : vk.cnt[0], k23 = c1 ? vk.cnt[3] : vk.cnt[2];
	const uint32_t d01 = c1 ? d1 : d0, d23 = c1 ? d3 : d2;
	const uint32_t t23 = d2 + d3, s0 = c1 ? t23 : t23 + d1, s1 = c1 ? 0u : d3;
	ak = (c2 ? k23 : k01) + ek;
	o2 = (c2 ? d23 : d01) + (el - ek);
	sd = (c2 ? s1 : s0) + (gl - gk);
}

// Shared memory of the device kernel, per warp: the 32 interval lists (three words per entry) and the 32 reads (4 bits per base), both
// laid out [word][lane] — a lane only ever touches bank `lane`, whatever list index or read position it is at, so no access conflicts.
// A list entry is {x0, x2, end-on-read}: the reverse-complement interval x1 is needed only while the match grows forward (it stays in
// registers); the backward phase of bwt_smem1a and the emitted seeds use x0 and x2 alone (B/bwt.c:326-345, B/bwamem.c:277-283).
#define LC_SEED_ENT_WORDS 3
LC_HD int lc_seed_q_words(int lmax) { return (lmax + 7) / 8; }
LC_HD int lc_seed_warp_words(int lmax, int lst_cap) { return (LC_SEED_ENT_WORDS * lst_cap + lc_seed_q_words(lmax)) * 32; }

// list entry <-> three words
template <typename U> struct LcEntOps;
template <> struct LcEntOps<uint32_t> {
	static LC_HD void pack(uint32_t x0, uint32_t x2, int e, uint32_t &w0, uint32_t &w1, uint32_t &w2) { w0 = x0; w1 = x2; w2 = (uint32_t)e; }
	static LC_HD void unpack(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &x0, uint32_t &x2, int &e) { x0 = w0; x2 = w1; e = (int)w2; }
};
template <> struct LcEntOps<uint64_t> {      // x < 2^34
	static LC_HD void pack(uint64_t x0, uint64_t x2, int e, uint32_t &w0, uint32_t &w1, uint32_t &w2)
	{ w0 = (uint32_t)x0 & LC_PK_LOW_MASK; w1 = (uint32_t)x2 & LC_PK_LOW_MASK; w2 = (uint32_t)(x0 >> LC_PK_LOW_BITS) | (uint32_t)(x2 >> LC_PK_LOW_BITS) << 2 | (uint32_t)e << 4; }
	static LC_HD void unpack(uint32_t w0, uint32_t w1, uint32_t w2, uint64_t &x0, uint64_t &x2, int &e)
	{ x0 = (uint64_t)(w2 & 3) << LC_PK_LOW_BITS | w0; x2 = (uint64_t)((w2 >> 2) & 3) << LC_PK_LOW_BITS | w1; e = (int)(w2 >> 4); }
};

template <typename U, bool COUNT = true, int STRIDE = 1> struct LcSeedLane {      // COUNT = false: the work counters are compiled out; STRIDE: words between
	                                                                                   // consecutive list/read words of this lane (32 on the device)
	typedef LcPkOps<U> PK;
	const LcBatch &b;
	typedef LcEntOps<U> EN;
	uint32_t *lst; int C;    // interval list of the current bwt_smem1a call: entry t = words lst[(3t + 0..2) * STRIDE] (shared memory on the device)
	const uint32_t *qw;      // the current read: qw[w * STRIDE] holds bases 8w..8w+7, 4 bits each, first base in the low nibble
	LcPk *memv; int M;       // seeds of the read so far, unsorted (global scratch of this lane)
	// ---- state
	int phase, r, len;
	U src, oth, x2;          // source interval of the pending extension; src = the side that is extended (x1 going forward, x0 going backward)
	int pend;                // its end on the read (bwtintv_t.info & 0xffffffff)
	int c;                   // base to extend with
	int i, x, min_intv;
	int n_fwd, n_prev, n_curr, j, last_start, ret;
	U last_x2;
	int n_mem, old_n, k2, pass;
	unsigned int cnt_ext, cnt_blk;   // algorithmic work of this read, flushed when the read completes

	LC_HD LcSeedLane(const LcBatch &b_, uint32_t *lst_, int C_, const uint32_t *qw_, LcPk *memv_, int M_)
		: b(b_), lst(lst_), C(C_), qw(qw_), memv(memv_), M(M_), phase(LC_SP_FETCH), r(0), len(0), src(1), oth(1), x2(0), c(0) {}

	LC_HD int qat(int t) const { return (int)(qw[(t >> 3) * STRIDE] >> ((t & 7) << 2)) & 15; }
	LC_HD void lst_put(int t, U x0_, U x2_, int e)
	{
		uint32_t w0, w1, w2;
		EN::pack(x0_, x2_, e, w0, w1, w2);
		uint32_t *p = lst + (size_t)(LC_SEED_ENT_WORDS * t) * STRIDE;
		p[0] = w0; p[STRIDE] = w1; p[2 * STRIDE] = w2;
	}
	LC_HD void set_intv(int base) { src = (U)(b.ix.L2[3 - base] + 1); oth = (U)(b.ix.L2[base] + 1); x2 = (U)(b.ix.L2[base + 1] - b.ix.L2[base]); }   // bwt_set_intv, forward roles
	LC_HD void push_fwd()            // forward phase: (x0, x1) = (oth, src)
	{
		if (n_fwd == C) { phase = LC_SP_OVERFLOW; return; }
		lst_put(n_fwd++, oth, x2, pend);
		ret = pend;
	}
	// backward phase: prev[j] and curr[j] of B/bwt.c:326-345 are both entry n_fwd - 1 - j (longest match first; curr is written
	// behind the read position, so one array serves both); src = x0, and x1 (`oth`) is dead from here on
	LC_HD void load_entry(int jj)
	{
		const uint32_t *p = lst + (size_t)(LC_SEED_ENT_WORDS * (n_fwd - 1 - jj)) * STRIDE;
		EN::unpack(p[0], p[STRIDE], p[2 * STRIDE], src, x2, pend);
	}
	LC_HD void emit(int start)       // a maximal match [start, pend) of the current call (B/bwt.c:334-338); backward roles
	{
		last_start = start;
		if (pend - start < b.opt.min_seed_len) return;            // B/bwamem.c:128,139
		if (n_mem == M) { phase = LC_SP_OVERFLOW; return; }
		memv[n_mem++] = PK::pack(src, 0, x2, start, pend);
	}

	// ---- the pending extension (bwt_extend, B/bwt.c:262-275; only the child for base c).  Executed by every lane on every tick;
	// lanes without a live read extend a harmless dummy interval.
	LC_HD void tick()
	{
		const LcIndex &ix = b.ix;
		const U prim = (U)ix.primary;
		U k = src - 1, l = k + x2;      // src >= 1 for every valid interval, so the row -1 of bwt_occ4 never occurs
		const bool kp = k >= prim, lp = l >= prim, adj = lp && !kp;      // adj: the interval contains the row of '$' (B/bwt.c:270)
		k -= (U)kp; l -= (U)lp;
		if (COUNT && b.cnt && phase >= LC_SP_FWD && phase < LC_SP_DONE) { ++cnt_ext; cnt_blk += (k >> 7) != (l >> 7) ? 2u : 1u; }   // 64-byte blocks of the REFERENCE layout (bwt_2occ4, B/bwt.c:194)
		const uint64_t bk = (uint64_t)(k >> 6), bl = (uint64_t)(l >> 6);
		// both blocks are requested before either is used, so their latencies overlap (same block: the second request hits L1)
		const LcBlk vk = lc_ld_blk(ix.occ + 2 * bk);
		const LcBlk vl = lc_ld_blk(ix.occ + 2 * bl);
		U o2, sn, on;
		if (sizeof(U) == 4) {
			uint32_t ak, d, sd;
			lc_blk_pair32(vk, vl, (int)(k & 63) + 1, (int)(l & 63) + 1, c, ak, d, sd);
			o2 = (U)d; sn = (U)((uint32_t)ix.L2[c] + 1u + ak); on = oth + (U)adj + (U)sd;
		} else {
			const uint64_t ch = (c & 2) ? ~0ull : 0ull, cl = (c & 1) ? ~0ull : 0ull;
			U ak, sk, al, sl;
			lc_blk_counts<U>(ix, vk, bk, (int)(k & 63) + 1, c, ch, cl, ak, sk);
			lc_blk_counts<U>(ix, vl, bl, (int)(l & 63) + 1, c, ch, cl, al, sl);
			o2 = al - ak; sn = (U)(ix.L2[c] + 1) + ak; on = oth + (U)adj + (sl - sk);
		}
		// The bookkeeping is a flat sequence of short guarded sections with a warp-wide convergence point between them
		// (LC_CONVERGE = __syncwarp on the device).  Without the barriers the compiler threads the jumps of the nested form and
		// every sub-path of a diverged warp runs its own copy of the common tail (ncu: profiles/r01d, ~4 live lanes there).
		const int ph = phase;
		const bool isb = ph == LC_SP_BWD, isf = ph == LC_SP_FWD || ph == LC_SP_P3, p3 = ph == LC_SP_P3;
		LC_CONVERGE();
		if (isb) {                                                    // backward extension of prev[j] done (B/bwt.c:330-343)
			if (o2 < (U)min_intv) {
				if (n_curr == 0 && i + 1 < last_start) emit(i + 1);
			} else if (n_curr == 0 || o2 != last_x2) {
				lst_put(n_fwd - 1 - n_curr, sn, o2, pend);
				last_x2 = o2; ++n_curr;
			}
		}
		LC_CONVERGE();
		if (isb && phase == LC_SP_BWD && ++j == n_prev) {             // end of a backward round
			if (n_curr == 0) phase = LC_SP_END_SMEM;
			else {
				n_prev = n_curr; n_curr = 0; j = 0; --i;
				const int qc = i >= 0 ? qat(i) : 4;
				if (qc > 3) phase = LC_SP_BWD_PREP; else c = qc;
			}
		}
		LC_CONVERGE();
		if (isb && phase == LC_SP_BWD) load_entry(j);
		LC_CONVERGE();
		bool adv = isf;
		if (isf) {                                                    // forward: bwt_smem1a (B/bwt.c:307-316) or bwt_seed_strategy1 (B/bwt.c:368-376)
			if (p3) {
				if (o2 < (U)b.opt.max_mem_intv && i - x >= b.opt.min_seed_len) {
					adv = false;
					if (o2 > 0 && n_mem == M) phase = LC_SP_OVERFLOW;
					else {
						if (o2 > 0) memv[n_mem++] = PK::pack(on, 0, o2, x, i + 1);      // B/bwamem.c:148
						x = i + 1; phase = LC_SP_P3_NEXT;
					}
				}
			} else if (o2 != x2) {
				push_fwd();
				if (phase == LC_SP_OVERFLOW) adv = false;
				else if (o2 < (U)min_intv) { phase = LC_SP_START_BWD; adv = false; }
			}
		}
		LC_CONVERGE();
		if (adv) {
			src = sn; oth = on; x2 = o2; pend = i + 1; ++i;
			// next forward step inline; anything unusual (end of read, ambiguous base, dead interval) goes through resolve()
			const int qc = i < len ? qat(i) : 4;
			if (qc > 3 || (p3 && o2 == 0)) phase = p3 ? LC_SP_P3_PREP : LC_SP_FWD_PREP;
			else c = 3 - qc;
		}
		LC_CONVERGE();
	}

	// ---- begin read r_ (its codes are at q_[0..len_), len_ >= min_seed_len)
	LC_HD void begin_read(int r_, int len_)
	{
		r = r_; len = len_;
		n_mem = 0; pass = 1; ret = 0; cnt_ext = cnt_blk = 0;
		phase = LC_SP_P1_NEXT;
	}

	// ---- bookkeeping states, one ordered sweep (a backward jump costs another sweep); leaves phase >= LC_SP_FWD or LC_SP_FETCH
	LC_HD void resolve()
	{
		const lc_opt_t &opt = b.opt;
		while (phase < LC_SP_FWD && phase != LC_SP_FETCH) {
			if (phase == LC_SP_OVERFLOW) {
				b.redo[lc_atomic_add(b.redo_cur, 1u)] = r;
				b.r_n_pk[r] = -1;
				phase = LC_SP_FETCH;
				break;
			}
			if (phase == LC_SP_FINISH) {                          // hand the unsorted seeds to lc_seed_finish
				if (COUNT && b.cnt) { lc_atomic_add(&b.cnt->extend_calls, (unsigned long long)cnt_ext); lc_atomic_add(&b.cnt->occ_blocks, (unsigned long long)cnt_blk); }
#if defined(LC_STATS) && !defined(__CUDA_ARCH__)
				if (COUNT && b.cnt) fprintf(stderr, "SEEDEXT %d %u %d\n", r, cnt_ext, n_mem);
#endif
				unsigned long long po = n_mem ? lc_carve_plain(b, b.pk_cur, b.pk_pool_cap, (unsigned long long)n_mem, LC_OVF_PK) : 0ull;
				if (po == (unsigned long long)-1) { po = 0; n_mem = 0; }
				for (int t = 0; t < n_mem; ++t) b.pk_pool[po + t] = memv[t];
				b.r_pk_off[r] = (int64_t)po; b.r_n_pk[r] = n_mem;
				phase = LC_SP_FETCH;
				break;
			}
			if (phase == LC_SP_END_SMEM) phase = pass == 1 ? LC_SP_P1_NEXT : LC_SP_NEXT_RESEED;
			if (phase == LC_SP_P1_NEXT) {                         // pass 1 (B/bwamem.c:121-131): next SMEM start, skipping ambiguous bases
				int xx = ret;
				while (xx < len && qat(xx) > 3) ++xx;
				if (xx < len) { x = xx; min_intv = 1; phase = LC_SP_START_SMEM; }
				else { pass = 2; old_n = n_mem; k2 = 0; phase = LC_SP_NEXT_RESEED; }
			}
			if (phase == LC_SP_NEXT_RESEED) {                     // pass 2 (B/bwamem.c:133-142): re-seed inside long near-unique SMEMs
				const int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
				phase = LC_SP_P3_NEXT; x = 0;
				while (k2 < old_n) {
					const LcPk e = memv[k2++];
					const int start = PK::e1(e), end = PK::e2(e);
					const U w = PK::x2(e);
					if (end - start < split_len || w > (U)opt.split_width) continue;
					x = (start + end) >> 1; min_intv = (int)w + 1; phase = LC_SP_START_SMEM;
					break;
				}
			}
			if (phase == LC_SP_START_SMEM) {                      // bwt_smem1a prologue (B/bwt.c:296-303); q[x] is a real base here
				set_intv(qat(x));
				pend = x + 1; n_fwd = 0; i = x + 1;
				phase = LC_SP_FWD_PREP;
			}
			if (phase == LC_SP_FWD_PREP) {
				if (i == len || qat(i) > 3) {                    // B/bwt.c:317-322
					push_fwd();
					if (phase != LC_SP_OVERFLOW) phase = LC_SP_START_BWD;
				} else { c = 3 - qat(i); phase = LC_SP_FWD; }
			}
			if (phase == LC_SP_START_BWD) {                       // B/bwt.c:323-326; ret = end of the longest forward match
				n_prev = n_fwd; i = x - 1; j = 0; n_curr = 0; last_start = 1 << 30;
				phase = LC_SP_BWD_PREP;
			}
			if (phase == LC_SP_BWD_PREP) {
				const int qc = i < 0 ? 4 : qat(i);
				load_entry(j);
				if (qc > 3) {
					// read start or ambiguous base: every entry stops here and only the first (longest, j == 0) can be a new
					// maximal match (B/bwt.c:332-338 with curr->n == 0 throughout)
					phase = LC_SP_END_SMEM;
					if (i + 1 < last_start) emit(i + 1);
				} else { c = qc; phase = LC_SP_BWD; }
			}
			if (phase == LC_SP_P3_NEXT) {                         // pass 3 (B/bwamem.c:144-159)
				if (!(opt.max_mem_intv > 0)) phase = LC_SP_FINISH;
				else {
					while (x < len && qat(x) > 3) ++x;
					if (x >= len) phase = LC_SP_FINISH;
					else { set_intv(qat(x)); i = x + 1; phase = LC_SP_P3_PREP; }      // B/bwt.c:364-366
				}
			}
			if (phase == LC_SP_P3_PREP) {
				if (i >= len) { x = len; phase = LC_SP_P3_NEXT; }                  // B/bwt.c:378
				else if (qat(i) > 3) { x = i + 1; phase = LC_SP_P3_NEXT; }           // B/bwt.c:377
				else if (x2 == 0) {
					// an empty interval stays empty, but the reference keeps extending it until the length test of B/bwt.c:371
					// passes (the resulting seed has x2 == 0 and is dropped, B/bwamem.c:148).  No lookups needed.
					int ii = i, xx;
					unsigned int n_ext = 0;
					for (;;) {
						if (ii >= len) { xx = len; break; }
						if (qat(ii) > 3) { xx = ii + 1; break; }
						++n_ext;
						if (ii - x >= opt.min_seed_len) { xx = ii + 1; break; }
						++ii;
					}
					cnt_ext += n_ext; cnt_blk += n_ext;           // k == l: one block each
					x = xx; phase = LC_SP_P3_NEXT;
				} else { c = 3 - qat(i); phase = LC_SP_P3; }
			}
		}
	}
};

// scalar driver (tests/hostsim): all reads, one after another, through the same state machine.  The device kernel does not run
// resolve() after every tick (lc_kernels.cu: lanes wait, ticking on their stale interval, until the warp's next housekeeping
// round); `idle_ticks` such ticks are inserted here before every resolve() so that the tests cover their harmlessness.
// `list` (n_items read ids) = a second pass over the reads an earlier pass handed back (larger lists); null = reads 0..n_items-1.
template <typename U> static inline void lc_seed_run_scalar(const LcBatch &b, uint32_t *lst, int C, uint32_t *qbuf, LcPk *memv, int M, int idle_ticks = 0,
															const int32_t *list = nullptr, int n_items = -1)
{
	LcSeedLane<U> sd(b, lst, C, qbuf, memv, M);
	int next = 0;
	if (n_items < 0) n_items = b.n_reads;
	for (;;) {
		while (sd.phase == LC_SP_FETCH) {
			if (next >= n_items) { sd.phase = LC_SP_DONE; break; }
			const int r = list ? list[next] : next, len = (int)(b.off[r + 1] - b.off[r]);
			++next;
			if (len < b.opt.min_seed_len) { b.r_n_pk[r] = 0; b.r_pk_off[r] = 0; continue; }      // B/bwamem.c:260
			for (int t = 0; t < (len + 7) / 8; ++t) qbuf[t] = 0;
			for (int t = 0; t < len; ++t) qbuf[t >> 3] |= (uint32_t)(b.seq[b.off[r] + t] & 15) << ((t & 7) << 2);
			sd.begin_read(r, len);
			for (int t = 0; t < idle_ticks; ++t) sd.tick();
			sd.resolve();
		}
		if (sd.phase == LC_SP_DONE) break;
		sd.tick();
		if (sd.phase < LC_SP_FWD) for (int t = 0; t < idle_ticks; ++t) sd.tick();
		sd.resolve();
	}
}

// ---- second step: sort, schedule, fetch positions -----------------------------------------------------------------
// Coop = a small lane group (device: 4 lanes, lc_kernels.cu) or one lane (hostsim).
struct LcCoopScalar {
	LC_HD bool leader() const { return true; }
	LC_HD int sub() const { return 0; }
	LC_HD int width() const { return 1; }
	LC_HD void sync() const {}
	LC_HD unsigned long long bcast(unsigned long long v) const { return v; }
};

// Sort the read's seeds by `info` (B/bwamem.c:161; equal keys are identical intervals, so any order of ties is the
// reference's), lay out the occurrence schedule of mem_chain (B/bwamem.c:273-279) and write the seed rows — or, when the
// dense suffix array is resident, directly the reference positions.
template <typename U, class Coop> LC_HD void lc_seed_finish(const LcBatch &b, const Coop &co, int r)
{
	typedef LcPkOps<U> PK;
	const int n = b.r_n_pk[r];
	if (n < 0) return;                                                // redone by lc_stage_seed, which writes the final form itself
	const LcPk *memv = b.pk_pool + b.r_pk_off[r];
	const uint64_t max_occ = (uint64_t)b.opt.max_occ;
	auto occ_step = [max_occ](uint64_t w) -> uint64_t { return w > max_occ ? w / max_occ : 1; };
	auto occ_count = [max_occ, occ_step](uint64_t w) -> uint32_t {
		if (w <= max_occ) return (uint32_t)w;                 // step 1: every occurrence (and no 64-bit division in the common case)
		const uint64_t st = occ_step(w); uint64_t cnt = (w + st - 1) / st; return (uint32_t)(cnt > max_occ ? max_occ : cnt); };
	unsigned long long S = 0;
	for (int f = 0; f < n; ++f) S += occ_count(PK::x2(memv[f]));
	unsigned long long io = 0, so = 0;
	if (co.leader()) {
		io = lc_carve(b, b.intv_cur, b.intv_cap, (unsigned long long)n, LC_OVF_INTV);
		so = lc_carve(b, b.seed_cur, b.seed_cap, S, LC_OVF_SEED);
	}
	io = co.bcast(io); so = co.bcast(so);
	if (io == (unsigned long long)-1 || so == (unsigned long long)-1) {
		if (co.leader()) { b.r_n_intv[r] = 0; b.r_n_seed[r] = 0; b.r_intv_off[r] = 0; b.r_seed_off[r] = 0; }
		return;
	}
	for (int e = co.sub(); e < n; e += co.width()) {              // rank sort: each lane places its own entries
		const LcPk me = memv[e];
		const uint32_t key = (uint32_t)PK::e1(me) << 12 | (uint32_t)PK::e2(me);
		int rank = 0;
		uint32_t soff = 0;
		for (int f = 0; f < n; ++f) {
			const LcPk o = memv[f];
			const uint32_t kf = (uint32_t)PK::e1(o) << 12 | (uint32_t)PK::e2(o);
			if (kf < key || (kf == key && f < e)) { ++rank; soff += occ_count(PK::x2(o)); }
		}
		LcIntvOut v;
		v.x0 = PK::x0(me); v.x2 = PK::x2(me); v.qbeg = (uint32_t)PK::e1(me); v.qend = (uint32_t)PK::e2(me);
		v.soff = soff; v.step = (uint32_t)occ_step(v.x2);
		b.intv[io + rank] = v;
	}
	co.sync();
	const bool dense = b.ix.sa32 || b.ix.sa64;
	for (int t = 0; t < n; ++t) {
		const LcIntvOut v = b.intv[io + t];
		const uint32_t cnt = occ_count(v.x2);
		for (uint32_t u = (uint32_t)co.sub(); u < cnt; u += (uint32_t)co.width()) {
			const uint64_t row = v.x0 + (uint64_t)u * v.step;
			b.seed_row[so + v.soff + u] = dense ? lc_sa(b.ix, row, b.cnt) : row;
		}
	}
	if (co.leader()) { b.r_intv_off[r] = (int64_t)io; b.r_n_intv[r] = n; b.r_seed_off[r] = (int64_t)so; b.r_n_seed[r] = (int32_t)S; }
}
