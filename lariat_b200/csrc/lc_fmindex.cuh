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
// Bump pools and work lists hang off a handful of counters: one atomic per read on the SAME word serialises in the L2 (~1 ns each,
// 2-6 M per step: it was half of seed_finish_kernel, chain_kernel, cigar_plan_kernel and pair_prepare_kernel).  In this form the
// lanes of a warp that arrive together with the same counter send ONE atomic for their sum and share the range out by lane order.
// For the thread-per-item kernels; the lane state machines keep the plain form (measured: the extension kernel loses 4 ms of 31
// to the extra instructions, its atomics are spread over the whole kernel anyway).
template <typename T> LC_HD T lc_atomic_add_agg(T *p, T v)
{
#ifdef __CUDA_ARCH__
	const unsigned act = __activemask();
	const unsigned grp = __match_any_sync(act, (unsigned long long)(uintptr_t)p);
	const int lane = threadIdx.x & 31;
	if (grp == (1u << lane)) return atomicAdd(p, v);
	T pre = 0, tot = 0;
	for (unsigned m = grp; m; m &= m - 1) {
		const int L = __ffs(m) - 1;
		const T o = __shfl_sync(grp, v, L);
		if (L < lane) pre += o;
		tot += o;
	}
	const int leader = __ffs(grp) - 1;
	T base = 0;
	if (lane == leader) base = atomicAdd(p, tot);
	return __shfl_sync(grp, base, leader) + pre;
#else
	return lc_atomic_add(p, v);
#endif
}
template <typename T> LC_HD T lc_max(T a, T b) { return a > b ? a : b; }
template <typename T> LC_HD T lc_min(T a, T b) { return a < b ? a : b; }

// optional algorithmic-work counters; null pointer = off
struct LcCounters {
	unsigned long long occ_blocks, extend_calls, sa_calls, lf_steps, ext_calls, ext_cells, sw_calls, sw_cells, glb_calls, glb_cells;
};

struct LcAnn { int64_t offset; int32_t len; int32_t is_alt; };

// The FM-index as it lives in HBM — NOT the .bwt file layout.  `bwa index` interleaves {u64 occ[4]; u32 sym[8]} per 128
// symbols (64 B; B/bwt.h:72-73, B/bwtindex.c:151-173), so one occ lookup touches two 32-byte sectors and needs up to
// eight masked word popcounts.  At load time the file payload is re-laid on the device (lc_occ_from_bwa) into one
// 32-byte block per 64 symbols:
//     { u32 cnt[4]; u64 hi; u64 lo }      cnt[c] = occurrences of c before the block, relative to its 2^32-symbol
//                                          superblock (absolute bases in LcIndex::sb); hi/lo = the high/low bit planes
//                                          of the 64 symbols, first symbol in bit 63
// Same bytes per base (0.5), but a lookup is ONE sector and at most three 64-bit popcounts.  Results are identical:
// occ(k, c) is a property of the BWT string, not of its packing.
#ifndef LC_SB_SHIFT              // tests/hostsim builds a second binary with a small value: several superblocks on a small index
#define LC_SB_SHIFT 26          // blocks per superblock = 2^26 (2^32 symbols)
#endif
#define LC_MAX_SB 4             // up to 2^34 symbols (the 3.1 Gb reference doubled is 6.2e9)
struct LcIndex {
	uint64_t primary, L2[5], seq_len;
	const uint4 *occ;        // 2 x uint4 per block: [0] = cnt[0..3], [1] = {hi.lo32, hi.hi32, lo.lo32, lo.hi32}
	uint64_t sb[LC_MAX_SB][4];
	const uint64_t *sa;      // sampled suffix array as in the .sa file; sa[0] = (u64)-1 (B/bwt.c:83,437)
	uint64_t sa_mask;        // sa_intv - 1
	int sa_shift;            // log2(sa_intv)
	const uint32_t *sa32;    // dense suffix array (every row), built on the device at load time when it fits; else null.
	const uint64_t *sa64;    //   sa32 when seq_len < 2^32 - 1, sa64 otherwise
	const uint8_t *pac;      // forward strand, 2 bits/base
	int64_t l_pac;
	const LcAnn *anns;
	int n_seqs;
};

struct LcIntv { uint64_t x0, x1, x2, info; };   // bwtintv_t (B/bwt.h:60-62): x0 fwd, x1 revcomp, x2 size

LC_HD int lc_popcll(uint64_t v)
{
#ifdef __CUDA_ARCH__
	return __popcll(v);
#else
	return __builtin_popcountll(v);
#endif
}
LC_HD int lc_ctz(uint32_t v)      // v != 0
{
#ifdef __CUDA_ARCH__
	return __ffs((int)v) - 1;
#else
	return __builtin_ctz(v);
#endif
}
LC_HD int lc_clz(uint32_t v)      // v != 0
{
#ifdef __CUDA_ARCH__
	return __clz((int)v);
#else
	return __builtin_clz(v);
#endif
}
// even bits of a 32-bit word gathered into 16 bits (bit 2i -> bit i)
LC_HD uint32_t lc_even_bits(uint32_t x)
{
	x &= 0x55555555u;
	x = (x | x >> 1) & 0x33333333u;
	x = (x | x >> 2) & 0x0f0f0f0fu;
	x = (x | x >> 4) & 0x00ff00ffu;
	x = (x | x >> 8) & 0x0000ffffu;
	return x;
}
// absolute counts before device block j (64 symbols), straight from the .bwt payload.  `bwa` = the u32 words after the
// 40-byte file header, zero-padded to whole 64-byte blocks.
LC_HD void lc_bwa_counts_at(const uint32_t *bwa, uint64_t j, uint64_t cnt[4])
{
	const uint32_t *o = bwa + 16 * (j >> 1);
	for (int c = 0; c < 4; ++c) cnt[c] = (uint64_t)o[2 * c] | (uint64_t)o[2 * c + 1] << 32;
	if (j & 1)      // add the first 64 symbols of the 128-symbol file block
		for (int t = 0; t < 4; ++t) {
			const uint32_t w = o[8 + t], hi = (w >> 1) & 0x55555555u, lo = w & 0x55555555u;
			const int n3 = lc_popc(hi & lo), n2 = lc_popc(hi) - n3, n1 = lc_popc(lo) - n3;
			cnt[3] += n3; cnt[2] += n2; cnt[1] += n1; cnt[0] += 16 - n3 - n2 - n1;
		}
}
// device block j from the .bwt payload; sb_base = absolute counts at the start of j's superblock
LC_HD void lc_occ_from_bwa(const uint32_t *bwa, uint64_t j, const uint64_t sb_base[4], uint4 out[2])
{
	const uint32_t *o = bwa + 16 * (j >> 1) + 8 + 4 * (j & 1);
	uint64_t cnt[4], hi = 0, lo = 0;
	lc_bwa_counts_at(bwa, j, cnt);
	for (int t = 0; t < 4; ++t) {
		hi = hi << 16 | lc_even_bits(o[t] >> 1);
		lo = lo << 16 | lc_even_bits(o[t]);
	}
	out[0].x = (uint32_t)(cnt[0] - sb_base[0]); out[0].y = (uint32_t)(cnt[1] - sb_base[1]);
	out[0].z = (uint32_t)(cnt[2] - sb_base[2]); out[0].w = (uint32_t)(cnt[3] - sb_base[3]);
	out[1].x = (uint32_t)hi; out[1].y = (uint32_t)(hi >> 32); out[1].z = (uint32_t)lo; out[1].w = (uint32_t)(lo >> 32);
}
LC_HD uint64_t lc_keep_first(int r) { return r >= 64 ? ~0ull : ~(~0ull >> r); }     // mask of the first r symbols (r = 0: none)

// occ(k, .) for the four bases = number of each base in BWT[0..k]   (bwt_occ4, B/bwt.c:169-186)
LC_HD void lc_occ4(const LcIndex &ix, uint64_t k, uint64_t c[4])
{
	if (k == (uint64_t)-1) { c[0] = c[1] = c[2] = c[3] = 0; return; }
	k -= (k >= ix.primary);
	const uint64_t blk = k >> 6;
	const uint4 q0 = lc_ld16(ix.occ + 2 * blk), q1 = lc_ld16(ix.occ + 2 * blk + 1);
	const int r = (int)(k & 63) + 1;
	const uint64_t m = lc_keep_first(r);
	const uint64_t hi = ((uint64_t)q1.y << 32 | q1.x) & m, lo = ((uint64_t)q1.w << 32 | q1.z) & m;
	const int n3 = lc_popcll(hi & lo), n2 = lc_popcll(hi) - n3, n1 = lc_popcll(lo) - n3;
	const uint64_t *sb = ix.sb[blk >> LC_SB_SHIFT];
	c[0] = sb[0] + q0.x + (uint64_t)(r - n3 - n2 - n1);
	c[1] = sb[1] + q0.y + (uint64_t)n1;
	c[2] = sb[2] + q0.z + (uint64_t)n2;
	c[3] = sb[3] + q0.w + (uint64_t)n3;
}

// the symbol at BWT position x ('$' removed)   (bwt_B0, B/bwt.h:78)
LC_HD int lc_bwt_sym(const LcIndex &ix, uint64_t x)
{
	const uint4 q1 = lc_ld16(ix.occ + 2 * (x >> 6) + 1);
	const int bit = 63 - (int)(x & 63);
	const uint64_t hi = (uint64_t)q1.y << 32 | q1.x, lo = (uint64_t)q1.w << 32 | q1.z;
	return (int)((hi >> bit) & 1) << 1 | (int)((lo >> bit) & 1);
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

// One LF step: the BWT row of the suffix that starts one text position earlier   (bwt_invPsi + bwt_occ, B/bwt.c:53-59,107-129).
// The symbol and its rank come from the same 32-byte block.
LC_HD uint64_t lc_lf(const LcIndex &ix, uint64_t k)
{
	if (k == ix.primary) return 0;
	const uint64_t x = k - (k > ix.primary);
	const uint64_t blk = x >> 6;
	const uint4 q0 = lc_ld16(ix.occ + 2 * blk), q1 = lc_ld16(ix.occ + 2 * blk + 1);
	const int off = (int)(x & 63), bit = 63 - off;
	const uint64_t hi = (uint64_t)q1.y << 32 | q1.x, lo = (uint64_t)q1.w << 32 | q1.z;
	const int c = (int)((hi >> bit) & 1) << 1 | (int)((lo >> bit) & 1);
	const uint64_t eq = ~(hi ^ ((c & 2) ? ~0ull : 0ull)) & ~(lo ^ ((c & 1) ? ~0ull : 0ull)) & lc_keep_first(off + 1);
	const uint32_t base = c == 0 ? q0.x : c == 1 ? q0.y : c == 2 ? q0.z : q0.w;
	return ix.L2[c] + ix.sb[blk >> LC_SB_SHIFT][c] + base + (uint64_t)lc_popcll(eq);
}

// Suffix-array value of BWT row k (bwt_sa, B/bwt.c:86-96).  With the dense array it is one load; otherwise the reference's
// LF-walk to a sampled row.  The counters always report the walk the REFERENCE would do.
LC_HD uint64_t lc_sa_walk(const LcIndex &ix, uint64_t k, uint64_t &steps)
{
	steps = 0;
	while (k & ix.sa_mask) { ++steps; k = lc_lf(ix, k); }
	return steps + lc_ldg(ix.sa + (k >> ix.sa_shift));
}
LC_HD uint64_t lc_sa(const LcIndex &ix, uint64_t k, LcCounters *cnt)
{
	uint64_t steps = 0, v;
	if ((ix.sa32 || ix.sa64) && !cnt) {
		if (ix.sa32) { const uint32_t t = lc_ldg(ix.sa32 + k); return t == 0xffffffffu ? (uint64_t)-1 : t; }
		return lc_ldg(ix.sa64 + k);
	}
	v = lc_sa_walk(ix, k, steps);
	if (cnt) { lc_atomic_add(&cnt->sa_calls, 1ull); lc_atomic_add(&cnt->lf_steps, (unsigned long long)steps); }
	return v;
}


// Dense suffix array, one sampled row at a time: SA[LF(k)] = SA[k] - 1, so the walk from sampled row j*intv labels every row it
// passes until it meets the next sampled row.  Over all samples every BWT row is labelled exactly once (LF is a permutation).
// T = uint32_t when seq_len < 2^32 - 1, else uint64_t; row 0 keeps the reference's sentinel (u64)-1 (B/bwt.c:83).
template <typename T> LC_HD void lc_sa_fill_sample(const LcIndex &ix, uint64_t j, T *out)
{
	const uint64_t k0 = j << ix.sa_shift;
	uint64_t v = j == 0 ? ix.seq_len : lc_ldg(ix.sa + j), k = k0;
	out[k0] = j == 0 ? (T)~(T)0 : (T)v;
	for (;;) {
		k = lc_lf(ix, k);
		if ((k & ix.sa_mask) == 0) break;
		--v;
		out[k] = (T)v;
	}
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
