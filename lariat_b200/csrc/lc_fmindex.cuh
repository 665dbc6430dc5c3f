// lc_fmindex.cuh — device-resident FM-index, packed reference and the walks over them.
//
// Everything here is `__host__ __device__` so that the per-read logic can also be compiled by a plain C++
// compiler for the CPU-side logic tests in tests/hostsim (never linked into libcuda_lariat.so; the product
// has no CPU execution path).  Reference lines ("B/" = go/src/gobwa/bwa/) are cited per function.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LC_HD __host__ __device__ __forceinline__
#define LC_HDN inline __host__ __device__ __noinline__
#else
#define LC_HD inline
#define LC_HDN inline
struct uint4 { unsigned x, y, z, w; };
#endif

LC_HD int lc_popc(unsigned v)
{
#ifdef __CUDA_ARCH__
	return __popc(v);
#else
	return __builtin_popcount(v);
#endif
}
LC_HD uint4 lc_ld16(const uint4 *p)
{
#ifdef __CUDA_ARCH__
	return __ldg(p);
#else
	return *p;
#endif
}
template <typename T> LC_HD T lc_ldg(const T *p)
{
#ifdef __CUDA_ARCH__
	return __ldg(p);
#else
	return *p;
#endif
}
template <typename T> LC_HD T lc_atomic_add(T *p, T v)
{
#ifdef __CUDA_ARCH__
	return atomicAdd(p, v);
#else
	T o = *p; *p += v; return o;   // hostsim runs the "threads" one after another
#endif
}
template <typename T> LC_HD T lc_max(T a, T b) { return a > b ? a : b; }
template <typename T> LC_HD T lc_min(T a, T b) { return a < b ? a : b; }

// optional algorithmic-work counters; null pointer = off
struct LcCounters {
	unsigned long long occ_blocks, extend_calls, sa_calls, lf_steps, ext_calls, ext_cells, sw_calls, sw_cells, glb_calls, glb_cells;
};

struct LcAnn { int64_t offset; int32_t len; int32_t is_alt; };

// The index as it lives in HBM.  bwt = the .bwt payload unchanged: one 64-byte block per 128 BWT symbols,
// { u64 occ[4]; u32 sym[8] } (2 bits/base, first symbol in the top bits)  (B/bwt.h:72-73, B/bwtindex.c:151-173),
// viewed as 4 x 16-byte vectors so a block is fetched with 128-bit loads.
struct LcIndex {
	uint64_t primary, L2[5], seq_len;
	const uint4 *bwt;
	const uint64_t *sa;      // sampled suffix array; sa[0] = (u64)-1 (B/bwt.c:83,437)
	uint64_t sa_mask;        // sa_intv - 1
	int sa_shift;            // log2(sa_intv)
	const uint8_t *pac;      // forward strand, 2 bits/base
	int64_t l_pac;
	const LcAnn *anns;
	int n_seqs;
};

struct LcIntv { uint64_t x0, x1, x2, info; };   // bwtintv_t (B/bwt.h:60-62): x0 fwd, x1 revcomp, x2 size

// per-base counts among the top `n` 2-bit cells of a 32-bit BWT word (n <= 0: none, n >= 16: all)
LC_HD void lc_count_word(unsigned w, int n, unsigned &c0, unsigned &c1, unsigned &c2, unsigned &c3)
{
	if (n <= 0) return;
	unsigned valid = n >= 16 ? 0x55555555u : (0x55555555u & ~((1u << ((16 - n) << 1)) - 1u));
	unsigned hi = (w >> 1) & valid, lo = w & valid;
	unsigned both = hi & lo;
	unsigned n3 = lc_popc(both), n2 = lc_popc(hi) - n3, n1 = lc_popc(lo) - n3;
	c3 += n3; c2 += n2; c1 += n1;
	c0 += (unsigned)(n >= 16 ? 16 : n) - n3 - n2 - n1;
}

// occ(k, .) for the four bases = number of each base in BWT[0..k]   (bwt_occ4, B/bwt.c:169-186)
LC_HD void lc_occ4(const LcIndex &ix, uint64_t k, uint64_t c[4])
{
	if (k == (uint64_t)-1) { c[0] = c[1] = c[2] = c[3] = 0; return; }
	k -= (k >= ix.primary);
	const uint4 *b = ix.bwt + ((k >> 7) << 2);
	uint4 q0 = lc_ld16(b), q1 = lc_ld16(b + 1), q2 = lc_ld16(b + 2);
	int r = (int)(k & 127) + 1;
	unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
	lc_count_word(q2.x, r, c0, c1, c2, c3);
	lc_count_word(q2.y, r - 16, c0, c1, c2, c3);
	lc_count_word(q2.z, r - 32, c0, c1, c2, c3);
	lc_count_word(q2.w, r - 48, c0, c1, c2, c3);
	if (r > 64) {
		uint4 q3 = lc_ld16(b + 3);
		lc_count_word(q3.x, r - 64, c0, c1, c2, c3);
		lc_count_word(q3.y, r - 80, c0, c1, c2, c3);
		lc_count_word(q3.z, r - 96, c0, c1, c2, c3);
		lc_count_word(q3.w, r - 112, c0, c1, c2, c3);
	}
	c[0] = ((uint64_t)q0.y << 32 | q0.x) + c0;
	c[1] = ((uint64_t)q0.w << 32 | q0.z) + c1;
	c[2] = ((uint64_t)q1.y << 32 | q1.x) + c2;
	c[3] = ((uint64_t)q1.w << 32 | q1.z) + c3;
}

// the symbol at BWT row x ('$' removed)   (bwt_B0, B/bwt.h:78)
LC_HD int lc_bwt_sym(const LcIndex &ix, uint64_t x)
{
	const unsigned *w = (const unsigned *)(ix.bwt + ((x >> 7) << 2) + 2);
	unsigned v = lc_ldg(w + ((x & 127) >> 4));
	return (v >> ((~x & 15) << 1)) & 3;
}

LC_HD void lc_set_intv(const LcIndex &ix, int c, LcIntv &ik)   // bwt_set_intv (B/bwt.h:80)
{
	ik.x0 = ix.L2[c] + 1; ik.x2 = ix.L2[c + 1] - ix.L2[c]; ik.x1 = ix.L2[3 - c] + 1; ik.info = 0;
}

// Extend a bi-interval by base c on one side; only the child for base c is produced   (bwt_extend, B/bwt.c:262-275).
// is_back = 1 prepends c to the match (uses x0), is_back = 0 appends the complement's complement (uses x1).
LC_HD void lc_extend(const LcIndex &ix, const LcIntv &in, int c, int is_back, LcIntv &out, LcCounters *cnt)
{
	uint64_t tk[4], tl[4];
	uint64_t src = is_back ? in.x0 : in.x1, oth = is_back ? in.x1 : in.x0;
	uint64_t k = src - 1, l = src - 1 + in.x2;
	lc_occ4(ix, k, tk);
	lc_occ4(ix, l, tl);
	if (cnt) {
		uint64_t k_ = k - (k >= ix.primary), l_ = l - (l >= ix.primary);
		lc_atomic_add(&cnt->extend_calls, 1ull);
		lc_atomic_add(&cnt->occ_blocks, (k == (uint64_t)-1 || l == (uint64_t)-1 || (k_ >> 7) != (l_ >> 7)) ? 2ull : 1ull);
	}
	uint64_t o = oth + (src <= ix.primary && src + in.x2 - 1 >= ix.primary);
	for (int b = 3; b > c; --b) o += tl[b] - tk[b];
	uint64_t s = ix.L2[c] + 1 + tk[c];
	out.x2 = tl[c] - tk[c];
	if (is_back) { out.x0 = s; out.x1 = o; } else { out.x1 = s; out.x0 = o; }
}

// All SMEMs through query position x (bwt_smem1a with max_intv = 0, B/bwt.c:289-351,353-356).
// mem[] receives them sorted by start; prev/curr are scratch lists of capacity len + 1.  Returns the next x.
LC_HDN int lc_smem1(const LcIndex &ix, int len, const uint8_t *q, int x, int min_intv, LcIntv *mem, int &n_mem,
					LcIntv *prev, LcIntv *curr, LcCounters *cnt)
{
	LcIntv ik, ok;
	int i, ret, n_prev, n_curr = 0;
	n_mem = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	lc_set_intv(ix, q[x], ik);
	ik.info = x + 1;
	for (i = x + 1; i < len; ++i) {                 // forward: push the interval whenever its size changes
		if (q[i] > 3) { curr[n_curr++] = ik; break; }
		lc_extend(ix, ik, 3 - q[i], 0, ok, cnt);
		if (ok.x2 != ik.x2) {
			curr[n_curr++] = ik;
			if (ok.x2 < (uint64_t)min_intv) break;
		}
		ok.info = i + 1;
		ik = ok;
	}
	if (i == len) curr[n_curr++] = ik;
	for (int a = 0, b = n_curr - 1; a < b; ++a, --b) { LcIntv t = curr[a]; curr[a] = curr[b]; curr[b] = t; }
	ret = (int)curr[0].info;
	{ LcIntv *t = curr; curr = prev; prev = t; }
	n_prev = n_curr;
	for (i = x - 1; i >= -1; --i) {                 // backward: shrink the list base by base
		int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		n_curr = 0;
		for (int j = 0; j < n_prev; ++j) {
			const LcIntv p = prev[j];
			if (c >= 0) lc_extend(ix, p, c, 1, ok, cnt);
			if (c < 0 || ok.x2 < (uint64_t)min_intv) {
				if (n_curr == 0 && (n_mem == 0 || (uint64_t)(i + 1) < (mem[n_mem - 1].info >> 32))) {
					ik = p; ik.info |= (uint64_t)(i + 1) << 32;
					mem[n_mem++] = ik;
				}
			} else if (n_curr == 0 || ok.x2 != curr[n_curr - 1].x2) {
				ok.info = p.info;
				curr[n_curr++] = ok;
			}
		}
		if (n_curr == 0) break;
		{ LcIntv *t = curr; curr = prev; prev = t; }
		n_prev = n_curr;
	}
	for (int a = 0, b = n_mem - 1; a < b; ++a, --b) { LcIntv t = mem[a]; mem[a] = mem[b]; mem[b] = t; }
	return ret;
}

// LAST-like forward seed (bwt_seed_strategy1, B/bwt.c:358-379)
LC_HDN int lc_seed_strategy1(const LcIndex &ix, int len, const uint8_t *q, int x, int min_len, int max_intv, LcIntv &mem, LcCounters *cnt)
{
	LcIntv ik, ok;
	mem.x0 = mem.x1 = mem.x2 = mem.info = 0;
	if (q[x] > 3) return x + 1;
	lc_set_intv(ix, q[x], ik);
	for (int i = x + 1; i < len; ++i) {
		if (q[i] > 3) return i + 1;
		lc_extend(ix, ik, 3 - q[i], 0, ok, cnt);
		if (ok.x2 < (uint64_t)max_intv && i - x >= min_len) {
			mem = ok;
			mem.info = (uint64_t)x << 32 | (uint32_t)(i + 1);
			return i + 1;
		}
		ik = ok;
	}
	return len;
}

// Suffix-array value of BWT row k: LF-walk to a sampled row (bwt_sa / bwt_invPsi / bwt_occ, B/bwt.c:53-59,86-129)
LC_HD uint64_t lc_sa(const LcIndex &ix, uint64_t k, LcCounters *cnt)
{
	uint64_t steps = 0;
	while (k & ix.sa_mask) {
		++steps;
		if (k == ix.primary) { k = 0; continue; }
		uint64_t x = k - (k > ix.primary);
		// symbol and its rank come from the same 64-byte block
		const uint4 *b = ix.bwt + ((x >> 7) << 2);
		uint4 q2 = lc_ld16(b + 2), q3 = lc_ld16(b + 3), q0 = lc_ld16(b), q1 = lc_ld16(b + 1);
		unsigned ws[8] = { q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w };
		int r = (int)(x & 127) + 1, c = 0;
		{
			int wi = (int)((x & 127) >> 4);
			unsigned v = 0;
#pragma unroll
			for (int t = 0; t < 8; ++t) if (t == wi) v = ws[t];
			c = (v >> ((~x & 15) << 1)) & 3;
		}
		unsigned n = 0;
#pragma unroll
		for (int t = 0; t < 8; ++t) {
			int m = r - 16 * t;
			if (m > 0) {
				unsigned valid = m >= 16 ? 0x55555555u : (0x55555555u & ~((1u << ((16 - m) << 1)) - 1u));
				unsigned w = ws[t];
				unsigned hi = (c & 2) ? (w >> 1) : ~(w >> 1), lo = (c & 1) ? w : ~w;
				n += lc_popc(hi & lo & valid);
			}
		}
		uint64_t base = c == 0 ? ((uint64_t)q0.y << 32 | q0.x) : c == 1 ? ((uint64_t)q0.w << 32 | q0.z)
					  : c == 2 ? ((uint64_t)q1.y << 32 | q1.x) : ((uint64_t)q1.w << 32 | q1.z);
		k = ix.L2[c] + base + n;
	}
	if (cnt) { lc_atomic_add(&cnt->sa_calls, 1ull); lc_atomic_add(&cnt->lf_steps, (unsigned long long)steps); }
	return steps + lc_ldg(ix.sa + (k >> ix.sa_shift));
}

// ---- packed reference ------------------------------------------------------------------------------
// base at doubled-coordinate p in [0, 2*l_pac): forward strand, then its reverse complement  (bns_get_seq, B/bntseq.c:398-419)
LC_HD int lc_ref_base(const LcIndex &ix, int64_t p)
{
	if (p >= ix.l_pac) { int64_t f = (ix.l_pac << 1) - 1 - p; return 3 - ((lc_ldg(ix.pac + (f >> 2)) >> ((~f & 3) << 1)) & 3); }
	return (lc_ldg(ix.pac + (p >> 2)) >> ((~p & 3) << 1)) & 3;
}

LC_HD int lc_pos2rid(const LcIndex &ix, int64_t pos_f)     // bns_pos2rid (B/bntseq.c:349-363)
{
	int lo = 0, mid = 0, hi = ix.n_seqs;
	if (pos_f >= ix.l_pac) return -1;
	while (lo < hi) {
		mid = (lo + hi) >> 1;
		if (pos_f >= ix.anns[mid].offset) {
			if (mid == ix.n_seqs - 1 || pos_f < ix.anns[mid + 1].offset) break;
			lo = mid + 1;
		} else hi = mid;
	}
	return mid;
}
LC_HD int64_t lc_depos(const LcIndex &ix, int64_t pos, int &is_rev)   // bns_depos (B/bntseq.h:87-90)
{
	is_rev = pos >= ix.l_pac;
	return is_rev ? (ix.l_pac << 1) - 1 - pos : pos;
}
LC_HD int lc_intv2rid(const LcIndex &ix, int64_t rb, int64_t re)      // bns_intv2rid (B/bntseq.c:365-373)
{
	int r;
	if (rb < ix.l_pac && re > ix.l_pac) return -2;
	int a = lc_pos2rid(ix, lc_depos(ix, rb, r));
	int b = rb < re ? lc_pos2rid(ix, lc_depos(ix, re - 1, r)) : a;
	return a == b ? a : -1;
}
// the clamping part of bns_fetch_seq (B/bntseq.c:421-446): [beg,end) is cut to the contig (and strand) holding `mid`
LC_HD int lc_fetch_clamp(const LcIndex &ix, int64_t &beg, int64_t mid, int64_t &end)
{
	int is_rev;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	int rid = lc_pos2rid(ix, lc_depos(ix, mid, is_rev));
	int64_t fb = ix.anns[rid].offset, fe = fb + ix.anns[rid].len;
	if (is_rev) { int64_t t = fb; fb = (ix.l_pac << 1) - fe; fe = (ix.l_pac << 1) - t; }
	if (beg < fb) beg = fb;
	if (end > fe) end = fe;
	return rid;
}
