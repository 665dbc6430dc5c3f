// lc_cigar.cuh — stage 6 (mem_reg2aln -> bwa_gen_cigar2 -> ksw_global2: B/bwamem.c:1086-1156, B/bwa.c:121-207, B/ksw.c:504-606)
// with one THREAD per candidate and the DP band held in registers.
//
// The first GPU form of this stage gave a warp to every candidate (lc_stage_cigar_t + lc_global2_warp): ncu
// (profiles/r01j_global_cigar_kernel_details.txt) showed 11 k warp-instructions per candidate, almost all of it the fixed
// per-row cost of the cooperative row (prefix-max scan, shuffles) spent on bands that are only 23-47 cells wide.
// Here the stage is split by what a candidate needs:
//   * lc_cig_plan (one thread per candidate): candidates whose alignment is gap-free (B/bwa.c:141-149; ~70 % on
//     C2-like data) are finished on the spot — a mismatch count; the others are binned by band width into job lists;
//   * lc_cig_band<NB> (one thread per job of a list): ksw_global2 in *diagonal* coordinates.  For row i the band is the
//     columns j = i - w + d, d in [0, 2w]; with d as the array index the cell (i, j) reads H(i-1, j-1) from the SAME slot d,
//     E(i, j) from slot d + 1 and F from slot d - 1, so H[] and E[] are fixed-size arrays with compile-time indices — they
//     live in registers, there is no shared memory and no per-row band bookkeeping.  The traceback matrix is 4 bits per
//     cell (the reference's byte holds three 2-bit fields of which only 4 bits carry information), one 32-bit word per 8
//     cells, in a per-job slice of a global pool; the same thread walks it back, merges the ops, counts mismatches and
//     writes the final CIGAR;
//   * everything unusual — bands wider than 80 cells, the rare band-doubling retry of mem_reg2aln (B/bwamem.c:1111-1119),
//     CIGARs with more than LC_CIG_MAXOPS merged ops, degenerate regions — goes to a fallback list that the warp-per-candidate
//     body (lc_stage_cigar_t) processes from scratch.
// Band values are 16-bit, two per register, with -16384 for the reference's MINUS_INF: exact under the range check lc_cig_fits16 (below).
// All bodies are LC_HD: tests/hostsim runs them serially on the CPU against the oracle.
#pragma once
#include "lc_mem.cuh"
#include "lc_extsm.cuh"      // the two-halves-per-register helpers (lc_p2_*, lc_prmt)

#define LC_CIG_NCLASS 5
#define LC_CIG_FALLBACK LC_CIG_NCLASS      // index of the fallback list
#define LC_CIG_NLIST (LC_CIG_NCLASS + 1)
#define LC_CIG_MAXOPS 24                   // merged CIGAR ops a thread keeps; longer CIGARs take the fallback

// 80 cells: a 150-base candidate whose score deficit comes from mismatches (diverged repeat copies) asks for w = 35-36 (B/bwa.c:152-157);
// that band no longer fits the register file, the compiler keeps part of it in (thread-interleaved, L1-resident) local memory — still
// an order of magnitude cheaper than the warp-per-candidate fallback (profiles/r2e: 139 ms of 166 on the 3.1 Gb config)
LC_HD int lc_cig_class_cells(int c) { return c == 0 ? 24 : c == 1 ? 32 : c == 2 ? 48 : c == 3 ? 64 : 80; }

struct LcCigJob {      // a candidate, decoded (B/bwamem.c:1086-1110)
	int r, l_query, qb, qe, lq, rlen, w2, truesc;
	int64_t rb, re;
	bool rev, bad;
	size_t slot;
	const uint8_t *query;
};

LC_HD void lc_cig_decode(const LcBatch &b, uint64_t cd, LcCigJob &j)
{
	const lc_opt_t &opt = b.opt;
	j.r = (int)(cd >> 32);
	const int idx = (int)(uint32_t)cd;
	j.l_query = (int)(b.off[j.r + 1] - b.off[j.r]);
	j.query = b.seq + b.off[j.r];
	j.slot = (size_t)b.r_fin_off[j.r] + idx;
	const lc_alnreg_t &ar = b.fin[j.slot];
	const int64_t l_pac = b.ix.l_pac;
	j.rb = ar.rb; j.re = ar.re; j.qb = ar.qb; j.qe = ar.qe; j.truesc = ar.truesc;
	j.lq = j.qe - j.qb; j.rlen = (int)(j.re - j.rb);
	j.rev = j.rb >= l_pac;
	int tmp = lc_infer_bw(j.lq, j.rlen, ar.truesc, opt.a, opt.o_del, opt.e_del);
	int w2 = lc_infer_bw(j.lq, j.rlen, ar.truesc, opt.a, opt.o_ins, opt.e_ins);
	w2 = lc_max(w2, tmp);
	if (w2 > opt.w) w2 = lc_min(w2, ar.w);
	j.w2 = lc_min(w2, opt.w << 2);                          // first statement of the retry loop (B/bwamem.c:1112)
	j.bad = j.rb < 0 || j.re < 0 || j.lq <= 0 || j.rb >= j.re || (j.rb < l_pac && j.re > l_pac) || j.re > l_pac << 1;
}

// band half-width of bwa_gen_cigar2 (B/bwa.c:152-157) for the first try
LC_HD int lc_cig_band_w(const LcBatch &b, const LcCigJob &j)
{
	int mi = (int)((double)(((j.lq + 1) >> 1) * b.sc.a - b.sc.o_ins) / b.sc.e_ins + 1.);
	int md = (int)((double)(((j.lq + 1) >> 1) * b.sc.a - b.sc.o_del) / b.sc.e_del + 1.);
	int mg = lc_max(lc_max(mi, md), 1), dl = j.rlen > j.lq ? j.rlen - j.lq : j.lq - j.rlen;
	int w = (mg + dl + 1) >> 1;
	w = lc_min(w, j.w2);
	w = lc_max(w, dl + 3);
	return w;
}

LC_HD void lc_cig_push_job(const LcBatch &b, int list, uint64_t cd)
{
	const unsigned int at = lc_atomic_add_agg(b.cig_n + list, 1u);
	b.cig_jobs[(size_t)list * b.cig_jobs_cap + at] = cd;
}

// oriented sequences as bwa_gen_cigar2 hands them to ksw_global2: on the reverse strand both are reversed (B/bwa.c:135-140)
LC_HD int lc_cig_q(const LcCigJob &j, int x) { return j.rev ? j.query[j.qe - 1 - x] : j.query[j.qb + x]; }
LC_HD int lc_cig_t(const LcBatch &b, const LcCigJob &j, int y) { return j.rev ? lc_ref_base(b.ix, j.re - 1 - y) : lc_ref_base(b.ix, j.rb + y); }

// NM, position, clipping and the final CIGAR of a candidate whose oriented alignment is cig[0..n_cigar) with n_mm mismatching
// aligned bases (B/bwa.c:169-199, B/bwamem.c:1122-1148).  Returns false when the CIGAR pool is full (flag raised).
LC_HD bool lc_cig_emit(const LcBatch &b, const LcCigJob &j, const uint32_t *cig, int n_cigar, int n_mm)
{
	const int64_t l_pac = b.ix.l_pac;
	int n_gap = 0;
	for (int k = 0; k < n_cigar; ++k) {
		const int op = cig[k] & 0xf, len = (int)(cig[k] >> 4);
		if (op == 2) { if (k > 0 && k < n_cigar - 1) n_gap += len; }
		else if (op == 1) n_gap += len;
	}
	int64_t pos = j.rev ? (l_pac << 1) - 1 - (j.re - 1) : j.rb;       // bns_depos (B/bwamem.c:1122)
	int cs = 0, ce = n_cigar;
	if (n_cigar > 0) {                                               // squeeze out a leading or trailing deletion (B/bwamem.c:1124-1133)
		if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; cs = 1; }
		else if ((cig[ce - 1] & 0xf) == 2) --ce;
	}
	const bool clipped = j.qb != 0 || j.qe != j.l_query;              // soft clips (B/bwamem.c:1134-1148)
	const int clip5 = clipped ? (j.rev ? j.l_query - j.qe : j.qb) : 0, clip3 = clipped ? (j.rev ? j.qb : j.l_query - j.qe) : 0;
	const int o = (clip5 ? 1 : 0) + (ce - cs) + (clip3 ? 1 : 0);
	const unsigned long long co = lc_carve(b, b.cigar_cur, b.cigar_cap, (unsigned long long)o, LC_OVF_CIGAR);
	if (co == (unsigned long long)-1) return false;
	uint32_t *dst = b.cigar + co;
	int k = 0;
	if (clip5) dst[k++] = (uint32_t)clip5 << 4 | 3;
	for (int t = cs; t < ce; ++t) dst[k++] = cig[t];
	if (clip3) dst[k++] = (uint32_t)clip3 << 4 | 3;
	lc_aln_t out;
	memset(&out, 0, sizeof(out));
	out.pos = pos; out.is_rev = j.rev ? 1 : 0; out.NM = n_mm + n_gap; out.n_cigar = o; out.cigar_off = (uint32_t)co; out.n_mismatch = n_mm;
	out.rid = lc_pos2rid(b.ix, pos);
	if (out.rid >= 0) out.pos = pos - b.ix.anns[out.rid].offset;
	b.aln[j.slot] = out;
	return true;
}

// ---- 16-bit cells.  The band kernel keeps two band slots per register (s16x2: VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2 / PRMT on
// sm_100a, the portable forms of lc_extsm.cuh in tests/hostsim) and uses -16384 for the reference's MINUS_INF (-2^30).  That is exact, not
// approximate: every value of the recurrence is either FINITE (the score of a real partial alignment, or H(i,-1)) or DERIVED FROM
// MINUS_INF (MINUS_INF plus a few scores/penalties).  Sums keep the category, a maximum of two values of one category compares them
// the same way under any constant shift, and between the categories the finite value wins in both arithmetics as long as
//     finite values  >  -16384 + a * rlen  (the highest a derived value can climb)      and      derived values  >=  -32768.
// lc_cig_fits16 checks exactly that for the job; jobs that fail (not reachable with Lariat's scoring and reads <= 700 bp) take the fallback.
#define LC_CIG_NEG16 (-16384)
LC_HD bool lc_cig_fits16(const LcBatch &b, const LcCigJob &j)
{
	const LcScore &sc = b.sc;
	const int oe_max = lc_max(sc.o_del + sc.e_del, sc.o_ins + sc.e_ins);
	const int pmax = lc_max(lc_max(sc.b, oe_max), 1);                 // the most one step of a path can cost
	const long long fin_lo = -((long long)(j.rlen + j.lq + 2) * pmax + 2LL * oe_max);      // paths are <= rlen + lq steps; + the t = M - oe intermediates
	const long long der_hi = (long long)LC_CIG_NEG16 + (long long)lc_max(sc.a, 0) * j.rlen;
	const long long der_lo = (long long)LC_CIG_NEG16 - (long long)(j.rlen + 2 + 96) * pmax - 2LL * oe_max;      // (+ F decaying across a whole row of slots that are not live)
	return fin_lo > der_hi + 64 && der_lo >= -32768 + 64 && (long long)j.lq * lc_max(sc.a, 0) < 32000;
}
// ---- one thread per candidate: finish the gap-free ones, bin the rest
LC_HD void lc_cig_plan(const LcBatch &b, uint64_t cd)
{
	LcCigJob j;
	lc_cig_decode(b, cd, j);
	if (j.bad) { lc_cig_push_job(b, LC_CIG_FALLBACK, cd); return; }
	if (j.lq == j.rlen && j.w2 == 0) {
		// gap-free (B/bwa.c:141-149).  Its score only feeds the retry test of mem_reg2aln, and a retry of a gap-free candidate
		// (w2 stays 0) repeats the same alignment, so only the mismatch count is needed.
		int n_mm = 0;
		for (int k = 0; k < j.lq; ++k) n_mm += lc_cig_q(j, k) != lc_cig_t(b, j, k);
		const uint32_t cig = (uint32_t)j.lq << 4;
		lc_cig_emit(b, j, &cig, 1, n_mm);
		return;
	}
	const int w = lc_cig_band_w(b, j), cells = 2 * w + 1;
	int cls = LC_CIG_FALLBACK;
	// (a small batch has warps to spare: the cooperative fill of the warp-per-candidate body is the short way through a band that one
	// thread would walk cell by cell)
	if (!b.small && lc_cig_fits16(b, j)) for (int c = LC_CIG_NCLASS - 1; c >= 0; --c) if (cells <= lc_cig_class_cells(c)) cls = c;
	lc_cig_push_job(b, cls, cd);
}

LC_HD void lc_st16(uint32_t *p, uint32_t a, uint32_t b_, uint32_t c, uint32_t d)
{
#ifdef __CUDA_ARCH__
	*(uint4 *)p = make_uint4(a, b_, c, d);
#else
	p[0] = a; p[1] = b_; p[2] = c; p[3] = d;
#endif
}
LC_HD uint32_t lc_funnel_r(uint32_t lo, uint32_t hi, int s)
{
#ifdef __CUDA_ARCH__
	return __funnelshift_r(lo, hi, s);
#else
	return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31));
#endif
}

// eight consecutive bases 8k .. 8k+7 of the oriented query, one nibble each (first base in the low nibble); 0 outside the query
LC_HD uint32_t lc_cig_qword(const LcCigJob &j, int k)
{
	uint32_t v = 0;
	const int x0 = k << 3;
	if (x0 + 8 <= 0 || x0 >= j.lq) return 0;
#pragma unroll
	for (int t = 0; t < 8; ++t) {
		const int x = x0 + t;
		if (x >= 0 && x < j.lq) v |= (uint32_t)(lc_cig_q(j, x) & 15) << (t << 2);
	}
	return v;
}

// four substitution scores at once: byte k of the result = byte ((sel >> 4k) & 7) of the pool {lut_lo, lut_hi} (one PRMT on the device)
LC_HD uint32_t lc_cig_lut4(uint32_t lut_lo, uint32_t lut_hi, uint32_t sel)
{
#ifdef __CUDA_ARCH__
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lut_lo), "r"(lut_hi), "r"(sel));
	return r;
#else
	const uint64_t pool = (uint64_t)lut_hi << 32 | lut_lo;
	uint32_t r = 0;
	for (int k = 0; k < 4; ++k) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
	return r;
#endif
}

LC_HD int lc_bitrev3(int c) { return ((c & 1) << 2) | (c & 2) | ((c >> 2) & 1); }
// bits [lo, hi) of a bitmap, the 32 of them that start at `base`
LC_HD uint32_t lc_range_word(int lo, int hi, int base)
{
	const int l = lo - base, h = hi - base;
	const uint32_t mh = h <= 0 ? 0u : (h >= 32 ? 0xffffffffu : (1u << h) - 1u);
	const uint32_t ml = l <= 0 ? 0u : (l >= 32 ? 0xffffffffu : (1u << l) - 1u);
	return mh & ~ml;
}

// ---- one thread per job: ksw_global2 (B/ksw.c:504-606) in diagonal coordinates, NB = band cells of the job's class (multiple of 8).
// Slot d of row i is column j0 + d (j0 = i - w); pair p holds slots 2p (low half) and 2p + 1.  A slot is LIVE in a row when its column
// is inside the query and d <= 2w; the reference computes no other cell.  One bitmap per row, expanded to half-word masks by a multiply
// and a sign-replicating PRMT, keeps the others out: their M is forced to "minus infinity" and their H to H(i,-1).  F runs left to right
// through the pairs (one VIADDMNMX and one PRMT per pair on the dependent chain); E and H are elementwise.  The four direction bits of a cell come from "maximum != operand" tests, one nibble per cell, eight
// cells per traceback word in bit-reversed nibble order (the order the packed halves fall into).
template <int NB> LC_HD void lc_cig_band(const LcBatch &b, uint64_t cd)
{
	constexpr int G = NB / 8, GS = (G + 3) & ~3, NP = NB / 2;      // GS: words per traceback row, padded to whole 16-byte stores
	LcCigJob j;
	lc_cig_decode(b, cd, j);
	const LcScore &sc = b.sc;
	const int w = lc_cig_band_w(b, j), w2x = 2 * w;
	const int lq = j.lq, rlen = j.rlen;
	// substitution scores by (query code ^ target code): 0 = match, 1..3 = mismatch, 4..7 = ambiguous query base (-1, B/bwa.c:109-118)
	const uint32_t lut_lo = ((uint32_t)sc.a & 0xffu) | ((uint32_t)(-sc.b) & 0xffu) * 0x01010100u, lut_hi = 0xffffffffu;
	const uint32_t n_oe_del = lc_p2_dup(-(sc.o_del + sc.e_del)), n_e_del = lc_p2_dup(-sc.e_del), n_oe_ins = lc_p2_dup(-(sc.o_ins + sc.e_ins)), n_e_ins = lc_p2_dup(-sc.e_ins);
	const uint32_t n_e12 = ((uint32_t)(-sc.e_ins) & 0xffffu) | ((uint32_t)(-2 * sc.e_ins) & 0xffffu) << 16;
	const uint32_t NEG2 = lc_p2_dup(LC_CIG_NEG16), ONE2 = 0x00010001u;
	// the job's slice of the traceback pool: GS words per row (every carve of this kernel is a multiple of 16 bytes)
	const unsigned long long zo = lc_carve(b, b.z_cur, b.z_cap, (unsigned long long)rlen * GS * 4, LC_OVF_Z);
	if (zo == (unsigned long long)-1) return;
	uint32_t *z = (uint32_t *)(b.zpool + zo);
	uint32_t H[NP], E[NP + 1];
	// row -1 (B/ksw.c:520-523): slot d holds H(-1, d - w - 1), the diagonal of cell (0, d - w)
#pragma unroll
	for (int p = 0; p < NP; ++p) {
		uint32_t v = 0;
#pragma unroll
		for (int hf = 0; hf < 2; ++hf) {
			const int d = 2 * p + hf;
			const int x = d == w ? 0 : (d > w && d <= w2x ? -(sc.o_ins + sc.e_ins * (d - w)) : LC_CIG_NEG16);
			v |= ((uint32_t)x & 0xffffu) << (16 * hf);
		}
		H[p] = v; E[p] = NEG2;
	}
	E[NP] = NEG2;
	// the query window: R[g] = oriented bases 8(kb+g) .. +7
	uint32_t R[G + 1];
	int kb = (-w) >> 3;
#pragma unroll
	for (int g = 0; g <= G; ++g) R[g] = lc_cig_qword(j, kb + g);
	// the target: forward-strand position of row 0 and its step, 16 bases per cached word of .pac
	const int64_t l_pac = b.ix.l_pac;
	const int64_t f0 = j.rev ? (l_pac << 1) - j.re : j.rb;            // rev: row i is ref[re-1-i] = complement of forward base 2*l_pac - re + i
	const uint32_t comp = j.rev ? 3u : 0u;
	const uint32_t *pacw = (const uint32_t *)b.ix.pac;
	uint32_t pw = lc_ldg(pacw + (f0 >> 4));
	for (int i = 0; i < rlen; ++i) {
		const int j0 = i - w;                                         // column of slot 0
		if (i > 0 && (j0 & 7) == 0) {                                 // the window crossed a word boundary
			++kb;
#pragma unroll
			for (int g = 0; g < G; ++g) R[g] = R[g + 1];
			R[G] = lc_cig_qword(j, kb + G);
		}
		const int64_t fp = f0 + i;
		if (i > 0 && (fp & 15) == 0) pw = lc_ldg(pacw + (fp >> 4));
		const uint32_t tb = ((pw >> ((((int)(fp >> 2) & 3) << 3) + ((~(int)fp & 3) << 1))) & 3u) ^ comp;
		const uint32_t tbm = tb * 0x11111111u;
		const int ph4 = (j0 & 7) << 2;
		// H(i, -1), the left boundary while the band touches column 0: every slot that is not live takes it (only the slot of column -1
		// is ever read with it: slots further left are rewritten row after row until they are that slot, slots right of the live
		// range never feed a live one)
		const uint32_t HV = lc_p2_dup(-(sc.o_del + sc.e_del * (i + 1)));
		// live slots of the row: [s_lo, s_hi)
		const int s_lo = j0 < 0 ? -j0 : 0, s_hi = lc_min(w2x + 1, lq - j0);
		uint32_t vb[(NB + 31) / 32];
#pragma unroll
		for (int k = 0; k < (NB + 31) / 32; ++k) vb[k] = lc_range_word(s_lo, s_hi, 32 * k);
		uint32_t fb = NEG2;                                           // F entering the next slot, in both halves
		uint32_t zrow[4];
#pragma unroll
		for (int g = 0; g < G; ++g) {
			const uint32_t qword = lc_funnel_r(R[g], R[g + 1], ph4);
			const uint32_t x = qword ^ tbm;                          // nibble c: 0 = match, 1..3 = mismatch, 4..7 = ambiguous query base
			const uint32_t sa = lc_cig_lut4(lut_lo, lut_hi, x), sb = lc_cig_lut4(lut_lo, lut_hi, x >> 16);
			// half-word masks of the eight slots: bit 2k of the byte to the sign of byte k of ye, bit 2k + 1 likewise of yo
			const uint32_t b8 = (vb[g >> 2] >> (8 * (g & 3))) & 0xffu;
			const uint32_t ye = (b8 & 0x55u) * 0x02082080u, yo = ((b8 >> 1) & 0x55u) * 0x02082080u;
			uint32_t rr[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				const int p = 4 * g + k;
				const uint32_t S = lc_prmt(k < 2 ? sa : sb, 0u, (k & 1) ? 0xb3a2u : 0x9180u);       // two scores, sign-extended to 16 bits
				const uint32_t mk = lc_prmt(ye, yo, 0xcc88u + 0x1111u * (uint32_t)k);      // halves = sign of byte k of ye / of yo
				// a slot that is not live computes with M = "minus infinity": whatever it then derives stays in that category and
				// cannot reach a live slot as a finite value (F flows through it, E is only read by the slot on its left, one row on)
				const uint32_t Mr = lc_p2_add(H[p], S);
				const uint32_t M = (Mr & mk) | (NEG2 & ~mk);
				const uint32_t eo = lc_prmt(E[p], E[p + 1], 0x5432u);                                // E(i, column) of the two slots
				const uint32_t mx = lc_p2_max(M, eo);
				const uint32_t A = lc_p2_add(M, n_oe_ins);
				// F(j+1) = max(F(j) - e, A(j)): for the pair {f1, f2} = max({f0 - e, f0 - 2e}, {a0, max(a0 - e, a1)}); only one VIADDMNMX
				// and the PRMT that spreads f2 are on the chain from pair to pair
				const uint32_t pp = lc_p2_addmax(lc_prmt(NEG2, A, 0x5410u), n_e_ins, A);
				const uint32_t fn = lc_p2_addmax(fb, n_e12, pp);                                     // F leaving the two slots
				const uint32_t fe = lc_prmt(fb, fn, 0x5410u);                                        // F entering them
				fb = lc_prmt(fn, fn, 0x3232u);
				const uint32_t hh = lc_p2_max(mx, fe);
				const uint32_t t1 = lc_p2_add(M, n_oe_del);
				const uint32_t en = lc_p2_addmax(eo, n_e_del, t1);
				H[p] = (hh & mk) | (HV & ~mk);
				E[p] = en;
				// direction bits (B/ksw.c:560-575): 1 = E beats M, 2 = F beats both, 4 = E extends, 8 = F extends
				const uint32_t c1 = lc_p2_minu(mx ^ M, ONE2), c2 = lc_p2_minu(hh ^ mx, ONE2), c4 = lc_p2_minu(en ^ t1, ONE2), c8 = lc_p2_minu(fn ^ A, ONE2);
				rr[k] = c1 + 2u * c2 + 4u * c4 + 8u * c8;
			}
			// nibble of slot c of the group at bit 4 * lc_bitrev3(c)
			zrow[g & 3] = (rr[0] | rr[1] << 8) | (rr[2] | rr[3] << 8) << 4;
			if ((g & 3) == 3 || g == G - 1) {
#pragma unroll
				for (int q = (g & 3) + 1; q < 4; ++q) zrow[q] = 0;
				lc_st16(z + (size_t)i * GS + (g & ~3), zrow[0], zrow[1], zrow[2], zrow[3]);
			}
		}
	}
	// H[qlen] of the reference = H(rlen-1, lq-1): slot lq - rlen + w
	int score = LC_NEG_INF;
	{
		const int dq = lq - rlen + w;
#pragma unroll
		for (int p = 0; p < NP; ++p) if (p == (dq >> 1)) score = (int)(int16_t)((dq & 1) ? H[p] >> 16 : H[p] & 0xffffu);
	}
	if (b.cnt) {
		unsigned long long cells = 0;
		for (int i = 0; i < rlen; ++i) { const int beg = i > w ? i - w : 0, end = i + w + 1 < lq ? i + w + 1 : lq; cells += (unsigned long long)(end > beg ? end - beg : 0); }
		lc_atomic_add(&b.cnt->glb_calls, 1ull); lc_atomic_add(&b.cnt->glb_cells, cells);
	}
	// mem_reg2aln retries with a doubled band when the global score falls short of the local one (B/bwamem.c:1111-1119): rare, warp path
	if (j.w2 != b.opt.w << 2 && score < j.truesc - b.opt.a) { lc_cig_push_job(b, LC_CIG_FALLBACK, cd); return; }
	// traceback (B/ksw.c:583-604): ops come out back to front and merge as they are pushed
	uint32_t ops[LC_CIG_MAXOPS];
	int n = 0, n_mm = 0, state = 0, i = rlen - 1, k = (i + w + 1 < lq ? i + w + 1 : lq) - 1;
	bool ovf = false;
	while (i >= 0 && k >= 0) {
		int d = k - i + w;
		d = d < 0 ? 0 : (d >= NB ? NB - 1 : d);
		const uint32_t nib = z[(size_t)i * GS + (d >> 3)] >> (lc_bitrev3(d & 7) << 2) & 15u;
		state = state == 0 ? ((nib & 2u) ? 2 : (int)(nib & 1u)) : state == 1 ? (int)(nib >> 2 & 1u) : (int)(nib >> 3 & 1u) * 2;
		int op;
		if (state == 0) { op = 0; n_mm += lc_cig_q(j, k) != lc_cig_t(b, j, i); --i; --k; }
		else if (state == 1) { op = 2; --i; }
		else { op = 1; --k; }
		if (n && (int)(ops[n - 1] & 0xf) == op) ops[n - 1] += 1u << 4;
		else if (n == LC_CIG_MAXOPS) { ovf = true; break; }
		else ops[n++] = 1u << 4 | (uint32_t)op;
	}
	if (!ovf && i >= 0) { if (n && (ops[n - 1] & 0xf) == 2u) ops[n - 1] += (uint32_t)(i + 1) << 4; else if (n == LC_CIG_MAXOPS) ovf = true; else ops[n++] = (uint32_t)(i + 1) << 4 | 2u; }
	if (!ovf && k >= 0) { if (n && (ops[n - 1] & 0xf) == 1u) ops[n - 1] += (uint32_t)(k + 1) << 4; else if (n == LC_CIG_MAXOPS) ovf = true; else ops[n++] = (uint32_t)(k + 1) << 4 | 1u; }
	if (ovf) { lc_cig_push_job(b, LC_CIG_FALLBACK, cd); return; }
	for (int a = 0; a < n >> 1; ++a) { const uint32_t t = ops[a]; ops[a] = ops[n - 1 - a]; ops[n - 1 - a] = t; }
	lc_cig_emit(b, j, ops, n, n_mm);
}
