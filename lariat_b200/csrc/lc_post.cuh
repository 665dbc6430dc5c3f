// lc_post.cuh — stage 6b: what inference.GetAlignments does with every candidate after mem_reg2aln (I/lariat.go:1573-1650), on the device.
//
// The Go loop, per candidate: fetch the reference span of the hit with GetSeq (G/gobwa.go:50-80 -> bns_fetch_seq: one more cgo call and
// one malloc/free per candidate), walk the CIGAR — backwards on the reverse strand, against the reverse-complemented span — count M /
// I / D / S, and record where read and reference differ inside M runs.  Everything it reads is already in HBM when stage 6 ends
// (region, CIGAR, read, packed reference), so one thread per candidate repeats the walk here and the result carries
// {pos, aend, matches, mismatches, indels, soft clips, mismatch positions}: the shim has no per-candidate call left to make.
// The walk follows the Go code statement by statement, including what it does at the edges (a span clamped at a contig end leaves
// zero bytes in the Go slice, which never equal a read base; positions past the span are skipped).  LC_HD: tests/hostsim runs it too.
#pragma once
#include "lc_mem.cuh"

// one candidate: b.fin[slot] (region), b.aln[slot] (mem_reg2aln's outcome), read r.  write = false: count only (n_mm).
LC_HD void lc_alnx_walk(const LcBatch &b, int r, size_t slot, lc_alnx_t &x, int32_t *mm_ref, int32_t *mm_read)
{
	const lc_alnreg_t &reg = b.fin[slot];
	const lc_aln_t &al = b.aln[slot];
	const int64_t l_pac = b.ix.l_pac;
	const int l_read = (int)(b.off[r + 1] - b.off[r]);
	const uint8_t *read = b.seq + b.off[r];
	const uint8_t *raw = b.raw ? b.raw + b.off[r] : (const uint8_t *)0;
	const int64_t coff = b.ix.anns[reg.rid].offset;
	// InterpretAlign (G/gobwa.go:352-363): Offset / Alignment_end, contig-relative
	const int64_t pos0 = reg.rb < l_pac ? reg.rb - coff : l_pac * 2 - 1 - reg.rb - coff;
	const int64_t aend0 = reg.re < l_pac ? reg.re - coff : l_pac * 2 - 1 - reg.re - coff;
	const bool rev = al.is_rev != 0;
	const int64_t ref_start = rev ? aend0 + 1 : pos0, ref_end = rev ? pos0 + 1 : aend0;       // I/lariat.go:1579-1584
	// GetSeq(chrom, refStart, refEnd, reversed) (G/gobwa.go:50-80): the slice has refEnd - refStart bytes; bns_fetch_seq clamps the
	// fetched interval to the contig holding its midpoint, the rest of the slice stays zero
	const int64_t L = ref_end - ref_start;
	int64_t fb = ref_start + coff, fe = ref_end + coff;
	int64_t clen = 0;
	if (L > 0) { lc_fetch_clamp(b.ix, fb, (ref_start + coff + ref_end + coff) >> 1, fe); clen = fe > fb ? fe - fb : 0; }
	int matches = 0, indels = 0, indel_length = 0, soft_n = 0, soft_len = 0, n_mm = 0;
	int64_t ref_o = 0; int read_o = 0;
	const uint32_t *cig = b.cigar + al.cigar_off;
	const int nc = al.n_cigar;
	for (int t = 0; t < nc; ++t) {                                     // I/lariat.go:1590-1632 (cigarStart / cigarIncrement)
		const uint32_t c = cig[rev ? nc - 1 - t : t];
		const int op = (int)(c & 0xf), len = (int)(c >> 4);
		if (op == 0) {
			matches += len;
			for (int m = 0; m < len; ++m) {
				const int64_t k = ref_o + m;
				if (k >= L) continue;                                  // :1599
				if (read_o + m >= l_read) continue;                    // (the Go code panics here; a CIGAR never describes more than the read)
				// refSeq[k]: forward = base(fb + k); reversed = complement of base(fb + clen - 1 - k); zero past the clamped part
				int rc = -1;
				if (k < clen) rc = rev ? 3 - lc_ref_base(b.ix, fb + clen - 1 - k) : lc_ref_base(b.ix, fb + k);
				// :1605 compares the bytes of the two Go slices: 'A' 'C' 'G' 'T' (twoBitToSeq / twoBitToSeqComp) against the read as given
				const bool differ = rc < 0 || (raw ? raw[read_o + m] != (uint8_t)"ACGT"[rc] : (read[read_o + m] > 3 || read[read_o + m] != rc));
				if (differ) {
					if (mm_ref) {
						mm_ref[n_mm] = (int32_t)(rev ? ref_end - k : k + ref_start);      // :1606-1610
						mm_read[n_mm] = read_o + m;
					}
					++n_mm;
				}
			}
			ref_o += len; read_o += len;
		} else if (op == 1) { ++indels; indel_length += len; read_o += len; }
		else if (op == 2) { ++indels; indel_length += len; ref_o += len; }
		else if (op == 3) { ++soft_n; soft_len += len; read_o += len; }
	}
	int mismatches = al.NM - indel_length;                              // :1633-1637
	matches -= mismatches;
	if (mismatches < 0) mismatches = 0;
	x.pos = pos0; x.aend = aend0;
	if (pos0 != -1 && rev) { x.pos = aend0 + 1; x.aend = pos0 + 1; }   // :1645-1650
	x.matches = matches; x.mismatches = mismatches; x.indels = indels; x.indel_length = indel_length;
	x.soft_clipped = soft_n; x.soft_clipped_length = soft_len; x.n_mm = n_mm;
}

// one thread per candidate: count, carve the two position lists, fill them
LC_HD void lc_stage_alnx(const LcBatch &b, uint64_t cd)
{
	if (*b.overflow & (LC_OVF_CIGAR | LC_OVF_Z)) return;      // the CIGAR stage is about to be rerun with larger pools: its outputs are not all there
	const int r = (int)(cd >> 32);
	const size_t slot = (size_t)b.r_fin_off[r] + (uint32_t)cd;
	lc_alnx_t x;
	lc_alnx_walk(b, r, slot, x, (int32_t *)0, (int32_t *)0);
	x.mm_off = 0;
	if (x.n_mm > 0) {
		const unsigned long long o = lc_carve(b, b.mm_cur, b.mm_cap, (unsigned long long)x.n_mm, LC_OVF_MM);
		if (o == (unsigned long long)-1) return;
		x.mm_off = (uint32_t)o;
		lc_alnx_t y;
		lc_alnx_walk(b, r, slot, y, b.mm_ref + o, b.mm_read + o);
	}
	b.alnx[slot] = x;
}
