// lc_sw_warp.cuh — warp-cooperative Smith-Waterman (device only).
//
// One warp per job.  Rows (reference bases) are processed in order, exactly like the reference, because the
// band pruning, the z-drop test and every tie-break are defined row by row.  Inside a row the 32 lanes take 32
// consecutive query columns per step.  The only loop-carried dependency along a row is F (a gap in the query
// direction); it is a max-plus linear recurrence
//        F(j+1) = max(F(j) - e_ins, T(j)),   T(j) = max(M(j) - o_ins - e_ins, 0)
// so it is evaluated for the 32 columns at once with a 5-step warp prefix-max of T(j) + j*e_ins.  H/E rows and the
// query live in shared memory, the row maximum / argmax and the band-shrink scans are warp reductions and ballots.
// Results are bit-identical to lc_extend2 / lc_sw_pass / lc_global2 (scalar forms, lc_sw.cuh), which follow the
// reference line by line; tests compare all three against the oracle.
#pragma once
#include "lc_sw.cuh"

#ifdef __CUDACC__
#define LC_FULL 0xffffffffu

__device__ __forceinline__ int lc_warp_prefix_max_incl(int v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		int o = __shfl_up_sync(LC_FULL, v, d);
		if (lane >= d) v = max(v, o);
	}
	return v;
}
__device__ __forceinline__ uint32_t lc_prmt_dev(uint32_t a, uint32_t b, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
__device__ __forceinline__ int lc_warp_max(int v)
{
#pragma unroll
	for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(LC_FULL, v, d));
	return v;
}

// ksw_extend2 (B/ksw.c:380-479), one warp.  H, E: shared memory, qlen + 2 ints each.  qs: shared memory, qlen bytes.
template <class QF, class TF>
__device__ __noinline__ LcExtResult lc_extend2_warp_smem(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int end_bonus, int zdrop,
														 int h0, int *H, int *E, uint8_t *qs, int lane, LcCounters *cnt)
{
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_ins = sc.e_ins, e_del = sc.e_del;
	int best = h0, best_i = -1, best_j = -1, g_i = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
	unsigned long long cells = 0;
	// row -1 (B/ksw.c:395-397): H[0] = h0, H[j] = max(h0 - oe_ins - (j-1) e_ins, 0), E = 0
	for (int j = lane; j <= qlen + 1; j += 32) {
		int v = 0;
		if (j == 0) v = h0;
		else if (j <= qlen) { v = h0 - oe_ins - (j - 1) * e_ins; if (v < 0) v = 0; }
		H[j] = v; E[j] = 0;
		if (j < qlen) qs[j] = (uint8_t)query(j);
	}
	{
		int smax = sc.a > 0 ? sc.a : 0;
		int mi = (int)((double)(qlen * smax + end_bonus - sc.o_ins) / sc.e_ins + 1.);
		int md = (int)((double)(qlen * smax + end_bonus - sc.o_del) / sc.e_del + 1.);
		w = min(w, max(mi, 1));
		w = min(w, max(md, 1));
	}
	__syncwarp();
	for (int i = 0; i < tlen; ++i) {
		const int tb = target(i);
		if (beg < i - w) beg = i - w;
		if (end > i + w + 1) end = i + w + 1;
		if (end > qlen) end = qlen;
		int h_first = 0;
		if (beg == 0) { h_first = h0 - (sc.o_del + sc.e_del * (i + 1)); if (h_first < 0) h_first = 0; }
		int f_carry = 0, hl_carry = h_first, key = -1;
		for (int cs = beg; cs < end; cs += 32) {
			const int j = cs + lane;
			const bool act = j < end;
			const int diag = act ? H[j] : 0, e = act ? E[j] : 0;
			int M = 0;
			if (act && diag) M = diag + sc.sub(tb, qs[j]);
			const int t = max(M - oe_ins, 0);
			const int pin = lc_warp_prefix_max_incl(t + j * e_ins, lane);
			const int pex = __shfl_up_sync(LC_FULL, pin, 1);
			int F = f_carry - lane * e_ins;
			if (lane > 0) F = max(F, pex - (j - 1) * e_ins);
			int h = max(M, e);
			h = max(h, F);
			const int e2 = max(e - e_del, max(M - oe_del, 0));
			int hl = __shfl_up_sync(LC_FULL, h, 1);
			if (lane == 0) hl = hl_carry;
			if (act) { H[j] = hl; E[j] = e2; key = max(key, (h << 10) | j); }
			const int last = min(31, end - 1 - cs);
			hl_carry = __shfl_sync(LC_FULL, h, last);
			f_carry = max(__shfl_sync(LC_FULL, F, 31) - e_ins, __shfl_sync(LC_FULL, t, 31));
		}
		cells += (unsigned long long)(end > beg ? end - beg : 0);
		const int jfin = beg < end ? end : beg;
		__syncwarp();
		if (lane == 0) { H[end] = hl_carry; E[end] = 0; }
		__syncwarp();
		if (jfin == qlen) {
			if (!(gscore > hl_carry)) g_i = i;
			gscore = max(gscore, hl_carry);
		}
		key = lc_warp_max(key);
		const int rowmax = key < 0 ? 0 : key >> 10, rowarg = key < 0 ? -1 : key & 1023;
		if (rowmax == 0) break;
		if (rowmax > best) {
			best = rowmax; best_i = i; best_j = rowarg;
			max_off = max(max_off, abs(rowarg - i));
		} else if (zdrop > 0) {
			const int di = i - best_i, dj = rowarg - best_j;
			if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) break; }
			else { if (best - rowmax - (dj - di) * e_ins > zdrop) break; }
		}
		// shrink the band where H and E are both zero (B/ksw.c:466-469)
		int nb = end;
		for (int cs = beg; cs < end; cs += 32) {
			const int j = cs + lane;
			const unsigned m = __ballot_sync(LC_FULL, j < end && (H[j] != 0 || E[j] != 0));
			if (m) { nb = cs + __ffs(m) - 1; break; }
		}
		beg = nb;
		int jj = beg - 1;
		for (int top = end; top >= beg; top -= 32) {
			const int j = top - lane;
			const unsigned m = __ballot_sync(LC_FULL, j >= beg && (H[j] != 0 || E[j] != 0));
			if (m) { jj = top - (__ffs(m) - 1); break; }
		}
		end = min(jj + 2, qlen);
	}
	__syncwarp();
	if (cnt && lane == 0) { atomicAdd(&cnt->ext_calls, 1ull); atomicAdd(&cnt->ext_cells, cells); }
	LcExtResult r = { best, best_j + 1, best_i + 1, g_i + 1, gscore, max_off };
	return r;
}

// ksw_extend2 with the H/E row in REGISTERS (same layout idea as lc_sw_pass_warp_reg below): lane L owns the K consecutive columns
// L*K .. L*K+K-1 of the row arrays (K * 32 >= qlen + 2).  The row arrays hold, as in the reference, H(i-1, j-1) at index j, so the
// diagonal needs no shuffle; what moves between lanes per row is the previous column's new h (one shuffle), the F scan across lanes
// (skipped when no lane hands a positive F on), the row maximum (REDUX) and the two band-shrink scans (two ballots).  Columns outside
// [beg, end) keep their values, contribute T = 0 to the F chain (so F enters `beg` as 0, B/ksw.c:424) and stay out of the maxima.
template <int K, class QF, class TF>
__device__ __noinline__ LcExtResult lc_extend2_warp_reg(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int end_bonus, int zdrop,
														int h0, int lane, LcCounters *cnt)
{
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_ins = sc.e_ins, e_del = sc.e_del;
	int best = h0, best_i = -1, best_j = -1, g_i = -1, gscore = -1, max_off = 0, beg = 0, end = qlen;
	unsigned long long cells = 0;
	const int j0 = lane * K;
	int Hs[K], Es[K];
	uint32_t stab[K];            // byte t = score of target base t (0..3) against this column's query base
#pragma unroll
	for (int k = 0; k < K; ++k) {                       // row -1 (B/ksw.c:395-397)
		const int j = j0 + k;
		int v = 0;
		if (j == 0) v = h0;
		else if (j <= qlen) { v = h0 - oe_ins - (j - 1) * e_ins; if (v < 0) v = 0; }
		Hs[k] = v; Es[k] = 0;
		uint32_t tw = 0;
		if (j < qlen) {
			const int q = query(j);
#pragma unroll
			for (int t = 0; t < 4; ++t) tw |= ((uint32_t)sc.sub(t, q) & 0xffu) << (8 * t);
		}
		stab[k] = tw;
	}
	{
		int smax = sc.a > 0 ? sc.a : 0;
		int mi = (int)((double)(qlen * smax + end_bonus - sc.o_ins) / sc.e_ins + 1.);
		int md = (int)((double)(qlen * smax + end_bonus - sc.o_del) / sc.e_del + 1.);
		w = min(w, max(mi, 1));
		w = min(w, max(md, 1));
	}
	bool stop = false;
	for (int i0 = 0; i0 < tlen && !stop; i0 += 32) {
		const int tmine = i0 + lane < tlen ? target(i0 + lane) : 4;
		const int nrow = min(32, tlen - i0);
		for (int ii = 0; ii < nrow; ++ii) {
			const int i = i0 + ii;
			const int tb = __shfl_sync(LC_FULL, tmine, ii);
			if (beg < i - w) beg = i - w;
			if (end > i + w + 1) end = i + w + 1;
			if (end > qlen) end = qlen;
			int h_first = 0;
			if (beg == 0) { h_first = h0 - (sc.o_del + sc.e_del * (i + 1)); if (h_first < 0) h_first = 0; }
			const uint32_t tsel = (uint32_t)(tb > 3 ? 4 : tb), sel = tsel | (tsel | 8u) << 4 | (tsel | 8u) << 8 | (tsel | 8u) << 12;
			int M[K], t[K], g = 0;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int j = j0 + k;
				const bool act = j >= beg && j < end;
				const int s = (int)lc_prmt_dev(stab[k], 0xffu, sel);      // ambiguous target base: -1 (B/bwa.c:109-118)
				const int m = (act && Hs[k]) ? Hs[k] + s : 0;
				M[k] = m;
				t[k] = __viaddmax_s32(m, -oe_ins, 0);
				g = k == 0 ? t[0] : __viaddmax_s32(g, -e_ins, t[k]);
			}
			int f = 0;
			if (__any_sync(LC_FULL, g > 0)) {
				const int pin = lc_warp_prefix_max_incl(g + (j0 + K - 1) * e_ins, lane);
				const int pex = __shfl_up_sync(LC_FULL, pin, 1);
				f = lane > 0 ? max(pex - (j0 - 1) * e_ins, 0) : 0;
			}
			int h[K], e2[K], key = -1, cand = 0;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int j = j0 + k;
				const bool act = j >= beg && j < end;
				const int hk = max(max(M[k], Es[k]), f);
				f = __viaddmax_s32(f, -e_ins, t[k]);
				e2[k] = __viaddmax_s32_relu(M[k], -oe_del, Es[k] - e_del);
				h[k] = hk;
				if (act) key = max(key, (hk << 10) | j);
				if (j == end - 1) cand = hk;
			}
			const bool nonempty = beg < end;
			int hl_carry = __shfl_sync(LC_FULL, cand, nonempty ? (end - 1) / K : 0);        // h of the row's last column
			if (!nonempty) hl_carry = h_first;
			const int hprev0 = __shfl_up_sync(LC_FULL, h[K - 1], 1);
			unsigned nz = 0;                                                                  // bit k: column j0 + k of the new row arrays is non-zero
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int j = j0 + k;
				const int hp = k == 0 ? hprev0 : h[k - 1];
				if (j >= beg && j < end) { Hs[k] = j == beg ? h_first : hp; Es[k] = e2[k]; }
				else if (j == end) { Hs[k] = hl_carry; Es[k] = 0; }
				if ((Hs[k] | Es[k]) != 0) nz |= 1u << k;
			}
			cells += (unsigned long long)(nonempty ? end - beg : 0);
			const int jfin = nonempty ? end : beg;
			if (jfin == qlen) {
				if (!(gscore > hl_carry)) g_i = i;
				gscore = max(gscore, hl_carry);
			}
			key = __reduce_max_sync(LC_FULL, key);
			const int rowmax = key < 0 ? 0 : key >> 10, rowarg = key < 0 ? -1 : key & 1023;
			if (rowmax == 0) { stop = true; break; }
			if (rowmax > best) {
				best = rowmax; best_i = i; best_j = rowarg;
				max_off = max(max_off, abs(rowarg - i));
			} else if (zdrop > 0) {
				const int di = i - best_i, dj = rowarg - best_j;
				if (di > dj) { if (best - rowmax - (di - dj) * e_del > zdrop) { stop = true; break; } }
				else { if (best - rowmax - (dj - di) * e_ins > zdrop) { stop = true; break; } }
			}
			// shrink the band where H and E are both zero (B/ksw.c:466-469)
			{
				unsigned m1 = 0, m2 = 0;                    // my columns in [beg, end) / in [beg, end] that are non-zero
#pragma unroll
				for (int k = 0; k < K; ++k) {
					const int j = j0 + k;
					if ((nz >> k & 1u) && j >= beg && j < end) m1 |= 1u << k;
					if ((nz >> k & 1u) && j >= beg && j <= end) m2 |= 1u << k;
				}
				const unsigned b1 = __ballot_sync(LC_FULL, m1 != 0);
				int nb = end;
				if (b1) { const int L = __ffs(b1) - 1; nb = L * K + (__ffs(__shfl_sync(LC_FULL, m1, L)) - 1); }
				// the last non-zero column in [nb, end]: columns of m2 left of nb are zero by construction of nb (they are in [beg, nb))
				const unsigned b2 = __ballot_sync(LC_FULL, m2 != 0);
				int jj = nb - 1;
				if (b2) { const int L = 31 - __clz(b2); jj = L * K + (31 - __clz(__shfl_sync(LC_FULL, m2, L))); }
				beg = nb;
				end = min(jj + 2, qlen);
			}
		}
	}
	if (cnt && lane == 0) { atomicAdd(&cnt->ext_calls, 1ull); atomicAdd(&cnt->ext_cells, cells); }
	LcExtResult r = { best, best_j + 1, best_i + 1, g_i + 1, gscore, max_off };
	return r;
}

template <class QF, class TF>
__device__ __forceinline__ LcExtResult lc_extend2_warp(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int end_bonus, int zdrop,
													   int h0, int *H, int *E, uint8_t *qs, int lane, LcCounters *cnt)
{
	if (qlen + 2 <= 128) return lc_extend2_warp_reg<4>(qlen, query, tlen, target, sc, w, end_bonus, zdrop, h0, lane, cnt);
	if (qlen + 2 <= 160) return lc_extend2_warp_reg<5>(qlen, query, tlen, target, sc, w, end_bonus, zdrop, h0, lane, cnt);
	if (qlen + 2 <= 256) return lc_extend2_warp_reg<8>(qlen, query, tlen, target, sc, w, end_bonus, zdrop, h0, lane, cnt);
	return lc_extend2_warp_smem(qlen, query, tlen, target, sc, w, end_bonus, zdrop, h0, H, E, qs, lane, cnt);
}

// One pass of the local SW with the reference's striped-SSE2 semantics (ksw_u8 / ksw_i16, B/ksw.c:111-334), one warp.
// Same contract as lc_sw_pass (lc_sw.cuh).  H, E, Hmax: shared memory rows of P ints; qs: P bytes; bl: global scratch list.
template <class QF, class TF>
__device__ __noinline__ LcSwResult lc_sw_pass_warp_smem(int qlen, QF query, int tlen, TF target, const LcScore &sc, int size, int xtra,
														int *H, int *E, int *Hmax, uint8_t *qs, unsigned long long *bl, int lane, LcCounters *cnt)
{
	const int lanes = size == 1 ? 16 : 8, P = (qlen + lanes - 1) / lanes * lanes;
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_ins = sc.e_ins, e_del = sc.e_del;
	const int shift = size == 1 ? max(sc.b, 1) : 0;
	const int minsc = (xtra & LC_XSUBO) ? (xtra & 0xffff) : 0x10000, endsc = (xtra & LC_XSTOP) ? (xtra & 0xffff) : 0x10000;
	LcSwResult r = { 0, -1, -1, -1, -1, -1, -1 };
	int gmax = 0, te = -1, n_b = 0, last_i = -2, last_max = 0, rows = 0;
	for (int j = lane; j < P; j += 32) { H[j] = 0; E[j] = 0; Hmax[j] = 0; qs[j] = j < qlen ? (uint8_t)query(j) : (uint8_t)4; }
	__syncwarp();
	for (int i = 0; i < tlen; ++i) {
		const int tb = target(i);
		int f_carry = 0, diag_carry = 0, imax = 0;
		++rows;
		for (int cs = 0; cs < P; cs += 32) {
			const int j = cs + lane;
			const bool act = j < P;
			const int old = act ? H[j] : 0, e = act ? E[j] : 0;
			int diag = __shfl_up_sync(LC_FULL, old, 1);
			if (lane == 0) diag = diag_carry;
			diag_carry = __shfl_sync(LC_FULL, old, 31);
			const int s = (act && j < qlen) ? sc.sub(tb, qs[j]) : 0;
			int hnf = max(diag + s, 0);
			hnf = max(hnf, e);
			const int t = max(hnf - oe_ins, 0);
			const int pin = lc_warp_prefix_max_incl(t + j * e_ins, lane);
			const int pex = __shfl_up_sync(LC_FULL, pin, 1);
			int F = max(f_carry - lane * e_ins, 0);
			if (lane > 0) F = max(F, pex - (j - 1) * e_ins);
			const int h = max(hnf, F);
			const int e2 = max(max(e - e_del, h - oe_del), 0);
			if (act) { H[j] = h; E[j] = e2; imax = max(imax, h); }
			f_carry = max(__shfl_sync(LC_FULL, F, 31) - e_ins, __shfl_sync(LC_FULL, t, 31));
		}
		imax = lc_warp_max(imax);
		__syncwarp();
		if (imax >= minsc) {                         // run-length list of rows scoring >= minsc (B/ksw.c:192-200)
			if (n_b == 0 || last_i + 1 != i) {
				if (lane == 0) bl[n_b] = (unsigned long long)imax << 32 | (unsigned)i;
				++n_b; last_i = i; last_max = imax;
			} else if (last_max < imax) {
				if (lane == 0) bl[n_b - 1] = (unsigned long long)imax << 32 | (unsigned)i;
				last_i = i; last_max = imax;
			}
		}
		if (imax > gmax) {
			gmax = imax; te = i;
			for (int j = lane; j < P; j += 32) Hmax[j] = H[j];
			if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
		}
	}
	__syncwarp();
	if (cnt && lane == 0) { atomicAdd(&cnt->sw_calls, 1ull); atomicAdd(&cnt->sw_cells, (unsigned long long)rows * qlen); }
	r.score = (size == 1 && gmax + shift >= 255) ? 255 : gmax;
	r.te = te;
	if (!(size == 1 && r.score == 255)) {
		int key = -1;                                 // smallest column holding the row maximum (B/ksw.c:211-216)
		for (int j = lane; j < P; j += 32) key = max(key, (Hmax[j] << 10) | (1023 - j));
		key = lc_warp_max(key);
		if (key >= 0) r.qe = 1023 - (key & 1023);
		if (n_b) {
			const int rad = (r.score + sc.a - 1) / sc.a, low = te - rad, high = te + rad;
			for (int i = 0; i < n_b; ++i) {
				const unsigned long long v = bl[i];
				const int e = (int)(unsigned)v;
				if ((e < low || e > high) && (int)(v >> 32) > r.score2) { r.score2 = (int)(v >> 32); r.te2 = e; }
			}
		}
	}
	__syncwarp();
	return r;
}

// The same pass with the H/E row in REGISTERS: lane L owns the K consecutive columns L*K .. L*K+K-1 (K * 32 >= P, so reads up to 255
// bases), a row is one sweep — no shared-memory traffic, one shuffle for the diagonal, one 5-step warp scan for F across lanes and one
// warp reduction for the row maximum, instead of all of that per 32-column chunk.  F inside a lane's block is the scalar recurrence
// F(j+1) = max(F(j) - e_ins, T(j)); the value entering the block comes from the scan over the blocks' outgoing values.  The row that
// last raised the maximum is not copied (Hmax): its smallest maximal column is tracked with the row maximum as one key.
// Cell arithmetic uses the DPX three-input forms (VIADDMNMX[.RELU]); substitution scores come from one PRMT on a per-column byte table.
// Target bases are fetched 32 rows at a time (one per lane) and handed out by shuffle, so no load sits on the row-to-row chain.
template <int K, class QF, class TF>
__device__ __noinline__ LcSwResult lc_sw_pass_warp_reg(int qlen, QF query, int tlen, TF target, const LcScore &sc, int size, int xtra,
													   unsigned long long *bl, int lane, LcCounters *cnt)
{
	const int lanes = size == 1 ? 16 : 8, P = (qlen + lanes - 1) / lanes * lanes;
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_ins = sc.e_ins, e_del = sc.e_del;
	const int shift = size == 1 ? max(sc.b, 1) : 0;
	const int minsc = (xtra & LC_XSUBO) ? (xtra & 0xffff) : 0x10000, endsc = (xtra & LC_XSTOP) ? (xtra & 0xffff) : 0x10000;
	LcSwResult r = { 0, -1, -1, -1, -1, -1, -1 };
	int gmax = 0, te = -1, qe_key = -1, n_b = 0, last_i = -2, last_max = 0, rows = 0;
	const int j0 = lane * K;
	int h[K], e[K];
	uint32_t stab[K];            // byte t = score of target base t against this column's query base (t = 0..3); padding columns score 0
#pragma unroll
	for (int k = 0; k < K; ++k) {
		h[k] = 0; e[k] = 0;
		const int j = j0 + k;
		uint32_t w = 0;
		if (j < qlen) {
			const int q = query(j);
#pragma unroll
			for (int t = 0; t < 4; ++t) w |= ((uint32_t)sc.sub(t, q) & 0xffu) << (8 * t);
		}
		stab[k] = w;
	}
	for (int i0 = 0; i0 < tlen; i0 += 32) {
		const int tmine = i0 + lane < tlen ? target(i0 + lane) : 4;
		const int nrow = min(32, tlen - i0);
		bool stop = false;
		for (int ii = 0; ii < nrow; ++ii) {
			const int i = i0 + ii;
			const int tb = __shfl_sync(LC_FULL, tmine, ii);
			++rows;
			// PRMT selector: byte tb of the column's table, sign-extended; an ambiguous target base selects byte 4 = the low byte of the
			// second operand (-1 on real columns, 0 on padding columns)
			const uint32_t tsel = (uint32_t)(tb > 3 ? 4 : tb), sel = tsel | (tsel | 8u) << 4 | (tsel | 8u) << 8 | (tsel | 8u) << 12;
			int dprev = __shfl_up_sync(LC_FULL, h[K - 1], 1);          // H(i-1, j0-1)
			if (lane == 0) dprev = 0;
			int hnf[K], t[K];
			int g = 0;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int diag = k == 0 ? dprev : h[k - 1];
				const int s = (int)lc_prmt_dev(stab[k], (j0 + k < qlen) ? 0xffu : 0u, sel);
				hnf[k] = __viaddmax_s32_relu(diag, s, e[k]);                 // max(diag + s, e, 0)
				t[k] = __viaddmax_s32(hnf[k], -oe_ins, 0);
				g = k == 0 ? t[0] : __viaddmax_s32(g, -e_ins, t[k]);          // F leaving column k if nothing entered the block
			}
			// F entering each lane's block: exclusive prefix maximum over the blocks' outgoing values, position-normalised
			// (g is what a block hands to the next lane's first column: when it is 0 in every lane — no cell of the row above the gap-open
			// cost, the usual case away from the alignment — nothing positive enters any block and the five-step scan is skipped)
			int f = 0;
			if (__any_sync(LC_FULL, g > 0)) {
				const int pin = lc_warp_prefix_max_incl(g + (j0 + K - 1) * e_ins, lane);
				const int pex = __shfl_up_sync(LC_FULL, pin, 1);
				f = lane > 0 ? max(pex - (j0 - 1) * e_ins, 0) : 0;
			}
			int key = -1;
#pragma unroll
			for (int k = 0; k < K; ++k) {
				const int hk = max(hnf[k], f);
				f = __viaddmax_s32(f, -e_ins, t[k]);
				h[k] = hk;
				e[k] = __viaddmax_s32_relu(hk, -oe_del, e[k] - e_del);        // max(h - oe_del, e - e_del, 0)
				if (j0 + k < P) key = max(key, (hk << 10) | (1023 - (j0 + k)));
			}
			key = __reduce_max_sync(LC_FULL, key);       // REDUX: one instruction instead of five shuffle rounds
			const int imax = key < 0 ? 0 : key >> 10;
			if (imax >= minsc) {                         // run-length list of rows scoring >= minsc (B/ksw.c:192-200)
				if (n_b == 0 || last_i + 1 != i) {
					if (lane == 0) bl[n_b] = (unsigned long long)imax << 32 | (unsigned)i;
					++n_b; last_i = i; last_max = imax;
				} else if (last_max < imax) {
					if (lane == 0) bl[n_b - 1] = (unsigned long long)imax << 32 | (unsigned)i;
					last_i = i; last_max = imax;
				}
			}
			if (imax > gmax) {
				gmax = imax; te = i; qe_key = key;
				if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) { stop = true; break; }
			}
		}
		if (stop) break;
	}
	__syncwarp();
	if (cnt && lane == 0) { atomicAdd(&cnt->sw_calls, 1ull); atomicAdd(&cnt->sw_cells, (unsigned long long)rows * qlen); }
	r.score = (size == 1 && gmax + shift >= 255) ? 255 : gmax;
	r.te = te;
	if (!(size == 1 && r.score == 255)) {
		if (qe_key >= 0) r.qe = 1023 - (qe_key & 1023);      // smallest column holding the maximum of the row that set it (B/ksw.c:211-216)
		if (n_b) {
			const int rad = (r.score + sc.a - 1) / sc.a, low = te - rad, high = te + rad;
			for (int i = 0; i < n_b; ++i) {
				const unsigned long long v = bl[i];
				const int ee = (int)(unsigned)v;
				if ((ee < low || ee > high) && (int)(v >> 32) > r.score2) { r.score2 = (int)(v >> 32); r.te2 = ee; }
			}
		}
	}
	__syncwarp();
	return r;
}

template <class QF, class TF>
__device__ __forceinline__ LcSwResult lc_sw_pass_warp(int qlen, QF query, int tlen, TF target, const LcScore &sc, int size, int xtra,
													  int *H, int *E, int *Hmax, uint8_t *qs, unsigned long long *bl, int lane, LcCounters *cnt)
{
	const int lanes = size == 1 ? 16 : 8, P = (qlen + lanes - 1) / lanes * lanes;
	if (P <= 128) return lc_sw_pass_warp_reg<4>(qlen, query, tlen, target, sc, size, xtra, bl, lane, cnt);
	if (P <= 160) return lc_sw_pass_warp_reg<5>(qlen, query, tlen, target, sc, size, xtra, bl, lane, cnt);
	if (P <= 256) return lc_sw_pass_warp_reg<8>(qlen, query, tlen, target, sc, size, xtra, bl, lane, cnt);
	return lc_sw_pass_warp_smem(qlen, query, tlen, target, sc, size, xtra, H, E, Hmax, qs, bl, lane, cnt);
}

// ksw_align2 as mem_matesw calls it (B/ksw.c:343-365), one warp
template <class QF, class TF>
__device__ __forceinline__ LcSwResult lc_sw_align2_warp(int qlen, QF query, int tlen, TF target, const LcScore &sc, int xtra,
													   int *H, int *E, int *Hmax, uint8_t *qs, unsigned long long *bl, int lane, LcCounters *cnt)
{
	const int size = (xtra & LC_XBYTE) ? 1 : 2;
	LcSwResult r = lc_sw_pass_warp(qlen, query, tlen, target, sc, size, xtra, H, E, Hmax, qs, bl, lane, cnt);
	if ((xtra & LC_XSTART) == 0 || ((xtra & LC_XSUBO) && r.score < (xtra & 0xffff))) return r;
	LcPrefixRev<QF> rq = { query, r.qe };
	LcPrefixRev<TF> rt = { target, r.te };
	LcSwResult rr = lc_sw_pass_warp(r.qe + 1, rq, tlen, rt, sc, size, LC_XSTOP | r.score, H, E, Hmax, qs, bl, lane, cnt);
	if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
	return r;
}

// Banded global alignment with traceback (ksw_global2, B/ksw.c:504-606), one warp: the fill is cooperative, the
// pointer-chasing traceback is done by lane 0.  Same contract as lc_global2 (lc_sw.cuh).
template <class QF, class TF>
__device__ __noinline__ int lc_global2_warp(int qlen, QF query, int tlen, TF target, const LcScore &sc, int w, int *H, int *E, uint8_t *qs,
											uint8_t *z, uint32_t *cigar, int cigar_cap, int *n_cigar, int lane, LcCounters *cnt)
{
	const int oe_del = sc.o_del + sc.e_del, oe_ins = sc.o_ins + sc.e_ins, e_ins = sc.e_ins, e_del = sc.e_del;
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	unsigned long long cells = 0;
	for (int j = lane; j <= qlen; j += 32) {
		H[j] = j == 0 ? 0 : (j <= w ? -(sc.o_ins + sc.e_ins * j) : LC_NEG_INF);
		E[j] = LC_NEG_INF;
		if (j < qlen) qs[j] = (uint8_t)query(j);
	}
	__syncwarp();
	for (int i = 0; i < tlen; ++i) {
		const int tb = target(i);
		const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int f_carry = LC_NEG_INF, hl_carry = beg == 0 ? -(sc.o_del + sc.e_del * (i + 1)) : LC_NEG_INF;
		uint8_t *zi = z ? z + (size_t)i * n_col : (uint8_t *)0;
		for (int cs = beg; cs < end; cs += 32) {
			const int j = cs + lane;
			const bool act = j < end;
			const int diag = act ? H[j] : LC_NEG_INF, e = act ? E[j] : LC_NEG_INF;
			const int M = diag + (act ? sc.sub(tb, qs[j]) : 0);
			const int t = M - oe_ins;
			const int pin = lc_warp_prefix_max_incl(t + j * e_ins, lane);
			const int pex = __shfl_up_sync(LC_FULL, pin, 1);
			int F = f_carry - lane * e_ins;
			if (lane > 0) F = max(F, pex - (j - 1) * e_ins);
			unsigned d = M >= e ? 0u : 1u;
			int h = M >= e ? M : e;
			d = h >= F ? d : 2u;
			h = h >= F ? h : F;
			const int t2 = M - oe_del, ee = e - e_del;
			if (ee > t2) d |= 1u << 2;
			const int e2 = ee > t2 ? ee : t2;
			if (F - e_ins > t) d |= 2u << 4;
			int hl = __shfl_up_sync(LC_FULL, h, 1);
			if (lane == 0) hl = hl_carry;
			if (act) { H[j] = hl; E[j] = e2; if (zi) zi[j - beg] = (uint8_t)d; }
			const int last = min(31, end - 1 - cs);
			hl_carry = __shfl_sync(LC_FULL, h, last);
			f_carry = max(__shfl_sync(LC_FULL, F, 31) - e_ins, __shfl_sync(LC_FULL, t, 31));
		}
		cells += (unsigned long long)(end > beg ? end - beg : 0);
		__syncwarp();
		if (lane == 0) { H[end] = hl_carry; E[end] = LC_NEG_INF; }
		__syncwarp();
	}
	const int score = H[qlen];
	if (cnt && lane == 0) { atomicAdd(&cnt->glb_calls, 1ull); atomicAdd(&cnt->glb_cells, cells); }
	if (n_cigar) {
		int n = 0;
		if (z && cigar && lane == 0) {
			int state = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
			bool ovf = false;
			while (i >= 0 && k >= 0) {
				if (n + 2 > cigar_cap) { ovf = true; break; }
				int col = k - (i > w ? i - w : 0);
				col = col < 0 ? 0 : (col >= n_col ? n_col - 1 : col);
				state = z[(size_t)i * n_col + col] >> (state << 1) & 3;
				if (state == 0) { n = lc_cigar_push(cigar, n, 0, 1); --i; --k; }
				else if (state == 1) { n = lc_cigar_push(cigar, n, 2, 1); --i; }
				else { n = lc_cigar_push(cigar, n, 1, 1); --k; }
			}
			if (!ovf && n + 2 > cigar_cap) ovf = true;
			if (!ovf) {
				if (i >= 0) n = lc_cigar_push(cigar, n, 2, i + 1);
				if (k >= 0) n = lc_cigar_push(cigar, n, 1, k + 1);
				for (int a = 0; a < n >> 1; ++a) { uint32_t tt = cigar[a]; cigar[a] = cigar[n - 1 - a]; cigar[n - 1 - a] = tt; }
			} else n = -1;
		}
		n = __shfl_sync(LC_FULL, n, 0);
		*n_cigar = n;
	}
	__syncwarp();
	return score;
}

struct LcExecWarp {
	int *H, *E;          // shared memory rows of this warp
	uint8_t *qs;         // shared memory copy of the query side of the current job
	int lane;
	int *Hmax;           // third shared row (local SW only)
	unsigned long long *bl;   // global scratch: sub-optimal row list of the local SW
	uint32_t *cig;       // global scratch: CIGAR build buffer (CIGAR stage only)
	uint8_t *zloc;       // global scratch: traceback matrix (CIGAR stage only)
	__device__ __forceinline__ int sum(int v) const
	{
#pragma unroll
		for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(LC_FULL, v, d);
		return v;
	}
	template <class QF, class TF>
	__device__ __forceinline__ int global2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int w, uint8_t *z, uint32_t *cigar, int cap,
										   int *n_cigar, LcCounters *cnt) const
	{
		return lc_global2_warp(qlen, q, tlen, t, sc, w, H, E, qs, z, cigar, cap, n_cigar, lane, cnt);
	}
	__device__ __forceinline__ int lane_id() const { return lane; }
	__device__ __forceinline__ int n_lanes() const { return 32; }
	__device__ __forceinline__ bool any(bool v) const { return __any_sync(LC_FULL, v) != 0; }
	__device__ __forceinline__ int max(int v) const { return __reduce_max_sync(LC_FULL, v); }
	template <class QF, class TF>
	__device__ __forceinline__ LcSwResult sw_align2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int xtra, LcCounters *cnt) const
	{
		return lc_sw_align2_warp(qlen, q, tlen, t, sc, xtra, H, E, Hmax, qs, bl, lane, cnt);
	}
	__device__ __forceinline__ bool leader() const { return lane == 0; }
	__device__ __forceinline__ void sync() const { __syncwarp(); }
	__device__ __forceinline__ unsigned long long bcast(unsigned long long v) const { return __shfl_sync(LC_FULL, v, 0); }
	template <class QF, class TF>
	__device__ __forceinline__ LcExtResult extend2(int qlen, QF q, int tlen, TF t, const LcScore &sc, int w, int end_bonus, int zdrop, int h0,
												   LcCounters *cnt) const
	{
		return lc_extend2_warp(qlen, q, tlen, t, sc, w, end_bonus, zdrop, h0, H, E, qs, lane, cnt);
	}
};

// bytes of dynamic shared memory one warp needs for reads up to lmax: `nrows` int rows of lmax+18 (rounded) + the query bytes
__host__ __device__ inline int lc_warp_row_ints(int lmax) { return (lmax + 18 + 31) & ~31; }
__host__ __device__ inline int lc_warp_smem_bytes(int lmax, int nrows = 2) { return nrows * 4 * lc_warp_row_ints(lmax) + lc_warp_row_ints(lmax); }
#endif
