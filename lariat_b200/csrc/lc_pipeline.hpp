// lc_pipeline.hpp — host-side orchestration of one batch: pool sizing, stage sequence, overflow/regrow, result packing.
//
// Templated on a Backend that owns memory and launches.  The product instantiates it with the CUDA backend
// (lc_api.cu: cudaMalloc + the __global__ kernels of lc_kernels.cu).  tests/hostsim instantiates it with a
// serial loop backend to check the orchestration logic on a machine without a GPU; that backend is test code
// and is never compiled into libcuda_lariat.so.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "lc_mem.cuh"
#include "lc_seed.cuh"
#include "lc_cigar.cuh"
#include "lc_post.cuh"

struct LcHostIndex {
	uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0;
	std::vector<uint32_t> bwt;          // 64-byte occ blocks, as in the .bwt file after its 40-byte header
	std::vector<uint64_t> sa;
	int sa_intv = 0;
	std::vector<uint8_t> pac;
	int64_t l_pac = 0;
	std::vector<LcAnn> anns;
	std::vector<std::string> names;
};

static inline bool lc_slurp(const std::string &fn, std::vector<uint8_t> &out)
{
	FILE *f = fopen(fn.c_str(), "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	out.resize((size_t)n);
	bool ok = n == 0 || fread(out.data(), 1, (size_t)n, f) == (size_t)n;
	fclose(f);
	return ok;
}

// Reads <prefix>.bwt/.sa/.ann/.pac(/.alt) — the files `bwa index` writes and bwa_idx_load reads
// (B/bwt.c:421-462, B/bntseq.c:98-206, B/bwa.c:262-289).
static inline bool lc_load_index(const char *prefix, LcHostIndex &ix, std::string &err)
{
	std::vector<uint8_t> raw;
	std::string p(prefix);
	if (!lc_slurp(p + ".bwt", raw) || raw.size() < 40 + 64) { err = "cannot read " + p + ".bwt"; return false; }
	memcpy(&ix.primary, raw.data(), 8);
	memcpy(&ix.L2[1], raw.data() + 8, 32);
	ix.L2[0] = 0;
	ix.seq_len = ix.L2[4];
	size_t nw = (raw.size() - 40) / 4;
	ix.bwt.assign(((nw + 15) / 16) * 16 + 32, 0);   // padded to whole 64-byte blocks (+ slack for the one-past-the-end device block)
	memcpy(ix.bwt.data(), raw.data() + 40, nw * 4);
	if (!lc_slurp(p + ".sa", raw) || raw.size() < 56) { err = "cannot read " + p + ".sa"; return false; }
	uint64_t h[7];
	memcpy(h, raw.data(), 56);
	if (h[0] != ix.primary || h[6] != ix.seq_len) { err = "SA-BWT inconsistency in " + p; return false; }
	ix.sa_intv = (int)h[5];
	if (ix.sa_intv <= 0 || (ix.sa_intv & (ix.sa_intv - 1))) { err = "SA interval is not a power of two"; return false; }
	size_t n_sa = (size_t)((ix.seq_len + ix.sa_intv) / ix.sa_intv);
	if (raw.size() < 56 + (n_sa - 1) * 8) { err = "truncated " + p + ".sa"; return false; }
	ix.sa.resize(n_sa);
	ix.sa[0] = (uint64_t)-1;
	memcpy(ix.sa.data() + 1, raw.data() + 56, (n_sa - 1) * 8);
	FILE *f = fopen((p + ".ann").c_str(), "r");
	if (!f) { err = "cannot read " + p + ".ann"; return false; }
	long long lp; int n_seqs; unsigned seed;
	if (fscanf(f, "%lld%d%u", &lp, &n_seqs, &seed) != 3) { fclose(f); err = "bad .ann header"; return false; }
	ix.l_pac = lp;
	ix.anns.resize(n_seqs);
	ix.names.resize(n_seqs);
	for (int i = 0; i < n_seqs; ++i) {
		unsigned gi; char name[8192], rest[8192];
		long long off; int len, n_ambs;
		if (fscanf(f, "%u%8191s", &gi, name) != 2 || !fgets(rest, sizeof rest, f) || fscanf(f, "%lld%d%d", &off, &len, &n_ambs) != 3) {
			fclose(f); err = "bad .ann record"; return false;
		}
		ix.names[i] = name;
		ix.anns[i].offset = off; ix.anns[i].len = len; ix.anns[i].is_alt = 0;
	}
	fclose(f);
	if ((f = fopen((p + ".alt").c_str(), "r")) != 0) {      // ALT contig list (B/bntseq.c:179-204)
		char line[8192];
		while (fgets(line, sizeof line, f)) {
			if (line[0] == '@') continue;
			line[strcspn(line, "\t\r\n")] = 0;
			for (int i = 0; i < n_seqs; ++i) if (ix.names[i] == line) ix.anns[i].is_alt = 1;
		}
		fclose(f);
	}
	if (!lc_slurp(p + ".pac", raw)) { err = "cannot read " + p + ".pac"; return false; }
	ix.pac.assign((size_t)(ix.l_pac / 4 + 2) + 16, 0);
	memcpy(ix.pac.data(), raw.data(), std::min(raw.size(), (size_t)(ix.l_pac / 4 + 1)));
	if (ix.seq_len != (uint64_t)ix.l_pac * 2) { err = "index does not hold forward+reverse text (seq_len != 2*l_pac)"; return false; }
	return true;
}

// number of 32-byte device blocks for an index of seq_len symbols (one spare so that row seq_len itself always has a block)
static inline uint64_t lc_occ_blocks(uint64_t seq_len) { return (seq_len + 63) / 64 + 1; }
// everything of LcIndex that does not point into device memory, incl. the superblock bases of the device occ layout
static inline bool lc_index_scalars(const LcHostIndex &h, LcIndex &ix, std::string &err)
{
	memset(&ix, 0, sizeof(ix));
	ix.primary = h.primary; memcpy(ix.L2, h.L2, sizeof(ix.L2)); ix.seq_len = h.seq_len;
	ix.sa_mask = (uint64_t)h.sa_intv - 1; ix.sa_shift = __builtin_ctz((unsigned)h.sa_intv);
	ix.l_pac = h.l_pac; ix.n_seqs = (int)h.anns.size();
	const uint64_t nb = lc_occ_blocks(h.seq_len);
	if (((nb - 1) >> LC_SB_SHIFT) >= LC_MAX_SB) { err = "reference too large: more than 2^34 BWT symbols"; return false; }
	for (uint64_t s = 0; (s << LC_SB_SHIFT) < nb; ++s) lc_bwa_counts_at(h.bwt.data(), s << LC_SB_SHIFT, ix.sb[s]);
	return true;
}
// tests/hostsim only: the device layout built with host loops (the product builds it with kernels, lc_kernels.cu)
struct LcHostDevIndex { std::vector<uint4> occ; std::vector<uint32_t> sa32; std::vector<uint64_t> sa64; };
static inline bool lc_host_dev_index(const LcHostIndex &h, LcHostDevIndex &d, LcIndex &ix, bool dense, std::string &err, bool wide_sa = false)
{
	if (!lc_index_scalars(h, ix, err)) return false;
	const uint64_t nb = lc_occ_blocks(h.seq_len);
	d.occ.resize(2 * nb);
	for (uint64_t j = 0; j < nb; ++j) lc_occ_from_bwa(h.bwt.data(), j, ix.sb[j >> LC_SB_SHIFT], &d.occ[2 * j]);
	ix.occ = d.occ.data(); ix.sa = h.sa.data(); ix.pac = h.pac.data(); ix.anns = h.anns.data();
	if (dense) {
		const uint64_t n_samples = h.sa.size();
		if (h.seq_len < 0xffffffffull && !wide_sa) { d.sa32.assign(h.seq_len + 1, 0); for (uint64_t j = 0; j < n_samples; ++j) lc_sa_fill_sample(ix, j, d.sa32.data()); ix.sa32 = d.sa32.data(); }
		else { d.sa64.assign(h.seq_len + 1, 0); for (uint64_t j = 0; j < n_samples; ++j) lc_sa_fill_sample(ix, j, d.sa64.data()); ix.sa64 = d.sa64.data(); }
	}
	return true;
}

static inline void lc_default_opt(lc_opt_t *o)    // mem_opt_init (B/bwamem.c:48-84)
{
	memset(o, 0, sizeof(*o));
	o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
	o->w = 100; o->T = 30; o->zdrop = 100; o->pen_unpaired = 17; o->pen_clip5 = o->pen_clip3 = 5;
	o->max_mem_intv = 20; o->min_seed_len = 19; o->split_width = 10; o->max_occ = 500; o->max_chain_gap = 10000;
	o->max_ins = 10000; o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->XA_drop_ratio = 0.80f; o->split_factor = 1.5f;
	o->chunk_size = 10000000; o->n_threads = 1; o->max_XA_hits = 5; o->max_XA_hits_alt = 200; o->max_matesw = 50;
	o->mask_level_redun = 0.95f; o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
	o->mapQ_coef_len = 50; o->mapQ_coef_fac = 3;    // (int)log(50)
	for (int i = 0, k = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) o->mat[k++] = (i == 4 || j == 4) ? -1 : (i == j ? o->a : -o->b);
}

// ASCII base or raw code -> 2-bit code; restates nst_nt4_table (B/bntseq.c:47-64) — '-' (5 there) folds to 4 like every consumer does
LC_HD uint8_t lc_nt4(uint8_t c, bool is_ascii)
{
	if (!is_ascii) return c > 4 ? 4 : c;
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 4; }
}

struct LcPackArgs {            // stage 7: dense result arrays
	const int64_t *out_off;    // [n_reads+1] exclusive scan of r_n_fin
	lc_alnreg_t *out_reg; lc_aln_t *out_aln;
	// a batch that is one chunk of a larger call (lc_api.cu): its candidates / CIGAR words start at these positions of the call's arrays
	int64_t *off_final;        // [n_reads+1] out_off + reg_base
	int64_t reg_base; uint32_t cig_base;
	lc_alnx_t *out_alnx; uint32_t mm_base;      // stage 6b outputs (null when off)
};

// r in [0, n_reads]: entry n_reads only carries the total
// The regions (out_reg, off_final) are final after the pair stage, the alignments (out_aln, out_alnx) after the CIGAR stage: the
// pipeline packs them in two launches (null = not in this launch) so that the regions' D2H runs under the CIGAR kernels.
LC_HD void lc_stage_pack(const LcBatch &b, const LcPackArgs &pa, int r)
{
	const int64_t dst = pa.out_off[r];
	if (pa.off_final) pa.off_final[r] = dst + pa.reg_base;
	if (r >= b.n_reads) return;
	const int n = b.r_n_fin[r];
	const int64_t src = b.r_fin_off[r];
	for (int i = 0; i < n; ++i) {
		if (pa.out_reg) pa.out_reg[dst + i] = b.fin[src + i];
		if (pa.out_aln) {
			lc_aln_t a = b.aln[src + i];
			a.cigar_off += pa.cig_base;
			pa.out_aln[dst + i] = a;
		}
		if (pa.out_alnx) { lc_alnx_t x = b.alnx[src + i]; x.mm_off += pa.mm_base; pa.out_alnx[dst + i] = x; }
	}
}

struct LcHostResult {
	std::vector<int64_t> reg_off;
	std::vector<lc_alnreg_t> regs;
	std::vector<lc_aln_t> aln;
	std::vector<uint32_t> cigars;
	std::vector<lc_alnx_t> alnx; std::vector<int32_t> mm_ref, mm_read;      // stage 6b (empty when off)
	float ms[10] = {0};        // h2d, seed, sa, chain, extend, (dedup: folded into pair), pair, cigar, d2h, total
	int64_t launches = 0;
	LcCounters counters;
};

// A growable device buffer
template <class BE> struct LcBuf {
	void *p = nullptr; size_t cap = 0;
	bool ensure(BE &be, size_t bytes, bool keep = false)
	{
		if (bytes <= cap) return true;
		size_t want = bytes + bytes / 4 + 256;
		void *np = be.dmalloc(want);
		if (!np) return false;
		if (keep && p && cap) be.d2d(np, p, cap);
		if (p) be.dfree(p);
		p = np; cap = want;
		return true;
	}
	void release(BE &be) { if (p) be.dfree(p); p = nullptr; cap = 0; }
};

static inline bool lc_verbose() { static const bool v = getenv("LC_VERBOSE") != nullptr; return v; }
template <class BE> class LcPipeline {
public:
	BE &be;
	LcBatch b;
	std::string err;
	int n_pairs = 0;
	bool counting = false;
	// tunables (bytes / elements per read); grown on overflow
	double intv_per_read = 24, seed_per_read = 48, cigar_per_cand = 6;
	size_t z_pool_bytes = 64u << 20;
	double z_per_cand = 1024;           // traceback pool of the register-band CIGAR kernels (4 bits per band cell)
	int memcap = 256;
	int lst_cap = 0, pk_cap = 64;       // seeding state machine: shared-memory interval list (0 = from the index size) / per-lane seed list; overflow -> second pass
	int ext_heavy_seeds = 32, ext_heavy_chains = 8;      // extension: reads above either go to the warp-per-read kernel
	bool force_coop = false;                             // tests (hostsim)
	bool prep_heavy = true;                              // stage 5a: pairs with very many regions get a warp
	int small_reads = 16384;                             // batches of at most this many reads run in the low-latency form (0 = never)
	int chain_heavy = 48;                                // chaining: reads with more seeds than this go to the warp-per-read kernel (0 = never)
	int lst_cap2 = 64, pk_cap2 = 256;     // the second and third pass' capacities; overflow of the third -> the one-thread-per-read redo kernel
	int lst_cap3 = 0, pk_cap3 = 2048;       // third pass: off by default (0); what outgrows the second goes to the warp-per-read kernel
	double pk_per_read = 20;            // pool of unsorted seeds between the state machine and lc_seed_finish
	bool force_wide = false;
	bool want_alnx = false;             // stage 6b: GetAlignments post-processing (lc_post.cuh)
	double mm_per_cand = 4;

	explicit LcPipeline(BE &be_) : be(be_) { memset(&b, 0, sizeof(b)); }
	~LcPipeline() { release(); }

	void set_index(const LcIndex &ix, const lc_opt_t &opt)
	{
		b.ix = ix; b.opt = opt;
		b.sc.a = opt.a; b.sc.b = opt.b; b.sc.o_del = opt.o_del; b.sc.e_del = opt.e_del; b.sc.o_ins = opt.o_ins; b.sc.e_ins = opt.e_ins;
		b.xc = lc_ext_const(b.sc);
		b.pes_low = -35; b.pes_high = 500; b.max_anchors = 50;     // G/gobwa.go:229-237,287,311
	}

	// ---- upload: bases are ASCII or codes (is_ascii), off[2*n_pairs+1]
	// shortest / longest read of off[0..R]: the caller's offsets are not trusted (lengths in int64).  Branch-free, so it vectorises.
	static void length_range(const int64_t *off, size_t R, int64_t &lo, int64_t &hi)
	{
		int64_t mn = 0, mx = 0;
		for (size_t r = 0; r < R; ++r) { const int64_t len = off[r + 1] - off[r]; mn = len < mn ? len : mn; mx = len > mx ? len : mx; }
		lo = mn; hi = mx;
	}
	// known_lmax > 0: the caller has validated the offsets (length_range) and passes the longest read
	int upload(int n_pairs_, const uint8_t *bases, const int64_t *off, bool is_ascii, int known_lmax = 0)
	{
		n_pairs = n_pairs_;
		const int R = 2 * n_pairs;
		const int64_t off0 = off[0], total = off[R] - off0;     // off0 != 0: the batch is a slice of a larger packed array
		int lmax = known_lmax;
		if (lmax <= 0) {                   // 0 <= len <= LC_MAX_READ_LEN for every read
			int64_t lo, hi;
			length_range(off, (size_t)R, lo, hi);
			if (lo < 0) { err = "read offsets must be non-decreasing"; return LC_EARG; }
			if (hi > LC_MAX_READ_LEN) { err = "read longer than LC_MAX_READ_LEN"; return LC_EARG; }
			lmax = std::max<int>(1, (int)hi);
		}
		be.ev_record(0);
		if (!d_seq.ensure(be, (size_t)total + 16) || !d_off.ensure(be, sizeof(int64_t) * (R + 1))) return nomem();
		if (total) be.h2d(d_seq.p, bases + off0, (size_t)total);
		b.raw = nullptr;
		if (want_alnx && is_ascii) {      // stage 6b compares the reads byte for byte, as the Go loop does: keep them beside the codes
			if (!d_raw.ensure(be, (size_t)total + 16)) return nomem();
			if (total) be.d2d(d_raw.p, d_seq.p, (size_t)total);
			b.raw = (const uint8_t *)d_raw.p;
		}
		be.h2d(d_off.p, off, sizeof(int64_t) * (R + 1));
		if (off0) be.k_rebase((int64_t *)d_off.p, R + 1, off0);
		be.k_convert((uint8_t *)d_seq.p, (unsigned long long)total, is_ascii);   // ASCII -> 2-bit codes on the device (G/gobwa.go:159-167)
		be.ev_record(1);
		b.n_reads = R; b.seq = (const uint8_t *)d_seq.p; b.off = (const int64_t *)d_off.p; b.lmax = lmax;
		return LC_OK;
	}

	// per-read arrays, control block (cursors, overflow word, counters), per-worker scratch
	int setup()
	{
		const int R = b.n_reads;
		launches = 0; launch_broken = false;
		if (!d_perread.ensure(be, (size_t)R * (8 * 6 + 4 * 8) + 256) || !d_nfin.ensure(be, 4 * ((size_t)R + 1))) return nomem();
		b.r_n_fin = (int32_t *)d_nfin.p;
		be.dzero(d_nfin.p, 4 * ((size_t)R + 1));
		{
			uint8_t *p = (uint8_t *)d_perread.p;
			b.r_intv_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_seed_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_chain_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_reg_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_fin_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_pk_off = (int64_t *)p; p += 8 * (size_t)R;
			b.r_n_intv = (int32_t *)p; p += 4 * (size_t)R;
			b.r_n_seed = (int32_t *)p; p += 4 * (size_t)R;
			b.r_n_chain = (int32_t *)p; p += 4 * (size_t)R;
			b.r_n_cseed = (int32_t *)p; p += 4 * (size_t)R;
			b.r_n_reg = (int32_t *)p; p += 4 * (size_t)R;
			b.r_n_pk = (int32_t *)p; p += 4 * (size_t)R;
			b.r_frac_rep = (float *)p;
		}
		if (!d_ctl.ensure(be, 4096)) return nomem();
		{
			unsigned long long *c = (unsigned long long *)d_ctl.p;
			b.intv_cur = c + 0; b.seed_cur = c + 1; b.scratch_cur = c + 2; b.chain_cur = c + 3; b.cseed_cur = c + 4;
			b.reg_cur = c + 5; b.fin_cur = c + 6; b.cand_cur = c + 7; b.cigar_cur = c + 8; b.z_cur = c + 9;
			b.reg_total = c + 10; b.pk_cur = c + 11; b.mm_cur = c + 12; b.sortk_cur = c + 13;
			b.prep_heavy_n = prep_heavy ? (unsigned int *)(c + 14) : nullptr;
			b.ext_unit_back = (unsigned int *)(c + 15);
			b.overflow = (uint32_t *)(c + 16);
			b.work_ctr = (unsigned int *)(c + 20);
			b.redo_cur = (unsigned int *)(c + 22);
			b.cig_n = (unsigned int *)(c + 24);
			b.resc_n = (unsigned int *)(c + 27);
			b.ext_heavy_n = (unsigned int *)(c + 28);
			b.chain_heavy_n = (unsigned int *)(c + 29);
			b.ext_unit_n = (unsigned int *)(c + 30); b.ext_hvu_n = (unsigned int *)(c + 31);      // adjacent: the backend reads both with one copy
			b.chain_heavy = chain_heavy;
			b.force_coop = force_coop ? 1 : 0;
			b.small = (small_reads > 0 && b.n_reads <= small_reads) ? 1 : 0;
			b.ext_heavy_seeds = ext_heavy_seeds; b.ext_heavy_chains = ext_heavy_chains;
			b.cnt = counting ? (LcCounters *)(c + 32) : nullptr;
		}
		be.dzero(d_ctl.p, 4096);
		b.n_workers = be.n_workers();
		b.sortk = nullptr; b.sortk_cap = 0; b.regtmp = nullptr;            // set by stage 5 (run_stages)
		b.memcap = memcap;
		{
			// a forward extension pushes one interval per change of size: ~log4(symbols) entries on unique sequence, more in repeats
			int want = lst_cap;
			if (want <= 0) { want = 2; for (uint64_t n = b.ix.seq_len; n > 1; n >>= 2) ++want; }
			b.lst_cap = std::min(want, b.lmax + 1); b.pk_cap = pk_cap;
		} b.force_wide = force_wide ? 1 : 0;
		if (!d_wpk.ensure(be, sizeof(LcPk) * (size_t)be.seed_lanes() * b.pk_cap) || !d_redo.ensure(be, 4 * ((size_t)R + 1))) return nomem();
		b.w_pk = (LcPk *)d_wpk.p; b.redo = (int32_t *)d_redo.p;
		b.chain_list = (int32_t *)d_redo.p;      // the seeding stage is done with its redo list when chaining starts
		if (!d_wl.ensure(be, 16 * ((size_t)R + 1))) return nomem();
		b.wl_in = (uint64_t *)d_wl.p; b.wl_out = b.wl_in + (R + 1);
		b.w_int_n = std::max(3 * (b.lmax + 18), 2 * (b.lmax + 2) + 2 * b.lmax + 8) + 8;
		b.w_u64_n = (b.pes_high - b.pes_low + b.lmax) / 2 + 8;
		b.w_z_n = 8192;
		if (!d_wintv.ensure(be, sizeof(LcIntv) * (size_t)be.n_redo_workers() * (3 * (size_t)(b.lmax + 1) + b.memcap)) ||
			!d_wint.ensure(be, 4 * (size_t)b.n_workers * b.w_int_n) || !d_wu64.ensure(be, 8 * (size_t)b.n_workers * b.w_u64_n) ||
			!d_wz.ensure(be, (size_t)b.n_workers * b.w_z_n)) return nomem();
		b.w_intv = (LcIntv *)d_wintv.p; b.w_int = (int32_t *)d_wint.p; b.w_u64 = (unsigned long long *)d_wu64.p; b.w_z = (uint8_t *)d_wz.p;
		return LC_OK;
	}
	int prepare_scratch() { return setup(); }

	// stage 1 with regrow + rerun on overflow
	int stage_seed()
	{
		const int R = b.n_reads;
		be.ev_record(2);
		for (int attempt = 0;; ++attempt) {
			b.intv_cap = (unsigned long long)(intv_per_read * R) + 1024;
			b.seed_cap = (unsigned long long)(seed_per_read * R) + 4096;
			b.pk_pool_cap = (unsigned long long)(pk_per_read * R) + 1024;
			if (!d_intv.ensure(be, sizeof(LcIntvOut) * b.intv_cap) || !d_seed.ensure(be, 8 * b.seed_cap) || !d_pkpool.ensure(be, sizeof(LcPk) * b.pk_pool_cap)) return nomem();
			b.intv = (LcIntvOut *)d_intv.p; b.seed_row = (uint64_t *)d_seed.p; b.pk_pool = (LcPk *)d_pkpool.p;
			LcBatch bf = b;                         // the batch as the fallback kernel sees it (its redo list)
			if (b.small) {
				// few reads (one barcode group per call, merged or not): a lane of the state machine would take the read's ~700 dependent
				// extensions one tick at a time; the warp-per-read kernel finishes the read in a fraction of that, and there are warps to spare
				be.k_iota(b.redo, R);
				be.dfill(b.r_n_pk, 0xff, 4 * (size_t)R);      // "redone": lc_seed_finish skips the read
				be.k_seed_redo(bf, R); ++launches;
				read_ctl();
				n_redo = n_redo_pass1 = 0;
			} else {
			be.k_seed(b, nullptr, R); ++launches;   // the state machine: unsorted seeds per read + the redo list
			read_ctl();
			n_redo = (int)(uint32_t)h_ctl[22];
			n_redo_pass1 = n_redo;
			// further passes: the reads whose lists outgrew the slots (repeats: many interval sizes, many seeds) run through the same
			// state machine with longer and longer lists — fewer resident warps each time, but only for the reads that need them.
			// The overflow lists alternate between two buffers; pass p counts its overflows in control word 23 (low / high half).
			const int caps[2][2] = { { lst_cap2, pk_cap2 }, { lst_cap3, pk_cap3 } };
			const int32_t *in_list = b.redo;
			for (int pass = 0; pass < 2 && n_redo && !(h_ctl_ovf & LC_OVF_PK); ++pass) {
				if (caps[pass][0] <= 0) continue;           // pass switched off
				LcBatch b2 = b;
				b2.lst_cap = std::min(caps[pass][0], b.lmax + 1); b2.pk_cap = caps[pass][1];
				LcBuf<BE> &out = pass == 0 ? d_redo2 : d_redo3;
				if (!out.ensure(be, 4 * ((size_t)n_redo + 1)) ||
					!d_wpk2.ensure(be, sizeof(LcPk) * (size_t)be.seed_pass_lanes(b2, n_redo) * b2.pk_cap)) return nomem();
				b2.w_pk = (LcPk *)d_wpk2.p; b2.redo = (int32_t *)out.p; b2.redo_cur = b.redo_cur + 2 + pass;
				be.k_seed(b2, in_list, n_redo); ++launches;
				read_ctl();
				if (lc_verbose()) fprintf(stderr, "[lc] seeding pass %d (lists of %d entries, %d seeds) ran %d of %d reads\n", pass + 2, b2.lst_cap, b2.pk_cap, n_redo, R);
				n_redo = (int)(uint32_t)(h_ctl[23] >> (32 * pass));
				bf.redo = b2.redo; in_list = b2.redo;
			}
			if (lc_verbose() && n_redo) fprintf(stderr, "[lc] seeding: %d reads left for the warp-per-read kernel\n", n_redo);
			if (!(h_ctl_ovf & LC_OVF_PK)) {
				be.k_seed_finish(b); ++launches;    // sort + occurrence schedule + positions
				if (n_redo) { be.k_seed_redo(bf, n_redo); ++launches; }
				read_ctl();
			}
			}
			if (launch_broken) { err = "a CUDA kernel launch failed in the seeding stage"; return LC_ECUDA; }
			if (!(h_ctl_ovf & (LC_OVF_INTV | LC_OVF_SEED | LC_OVF_MEMLIST | LC_OVF_PK))) break;
			if (attempt > 12) { err = "seeding pools keep overflowing"; return LC_ENOMEM; }
			if (h_ctl_ovf & LC_OVF_INTV) intv_per_read = std::max(intv_per_read * 2, (double)h_ctl[0] / R * 1.25);
			if (h_ctl_ovf & LC_OVF_SEED) seed_per_read = std::max(seed_per_read * 2, (double)h_ctl[1] / R * 1.25);
			if (h_ctl_ovf & LC_OVF_PK) pk_per_read = std::max(pk_per_read * 2, (double)h_ctl[11] / R * 1.25);
			if (h_ctl_ovf & LC_OVF_MEMLIST) {
				memcap *= 4; b.memcap = memcap;
				if (!d_wintv.ensure(be, sizeof(LcIntv) * (size_t)be.n_redo_workers() * (3 * (size_t)(b.lmax + 1) + b.memcap))) return nomem();
				b.w_intv = (LcIntv *)d_wintv.p;
			}
			be.dzero(d_ctl.p, 4096);
		}
		be.ev_record(3);
		return LC_OK;
	}
	int run_seed_only() { int rc = setup(); if (!rc) rc = stage_seed(); return rc; }
	unsigned long long seed_intv_used() const { return h_ctl[0]; }

	// mem_reg2aln of one region against read 0 of the uploaded batch (source-compatible single-read form)
	int run_single_cigar(const lc_alnreg_t &ar, LcHostResult &res)
	{
		int rc = setup();
		if (rc) return rc;
		b.fin_cap = 4;
		if (!d_fin.ensure(be, sizeof(lc_alnreg_t) * 4) || !d_aln.ensure(be, sizeof(lc_aln_t) * 4) || !d_cand.ensure(be, 32)) return nomem();
		b.fin = (lc_alnreg_t *)d_fin.p; b.aln = (lc_aln_t *)d_aln.p; b.cand = (uint64_t *)d_cand.p;
		be.dzero(d_perread.p, (size_t)b.n_reads * (8 * 5 + 4 * 7));
		be.dzero(d_cand.p, 32);
		be.h2d(d_fin.p, &ar, sizeof(ar));
		const unsigned long long need = (unsigned long long)(b.lmax + 2) * (unsigned long long)(ar.re > ar.rb ? ar.re - ar.rb : 1) + 4096;
		b.cigar_cap = 4 * (unsigned long long)b.lmax + 64; b.z_cap = need;
		if (!d_cigar.ensure(be, 4 * b.cigar_cap) || !d_z.ensure(be, b.z_cap)) return nomem();
		b.cigar = (uint32_t *)d_cigar.p; b.zpool = (uint8_t *)d_z.p; b.alnx = nullptr;
		be.k_cigar(b, b.cand, 1, nullptr); ++launches;
		read_ctl();
		if (h_ctl_ovf) { err = "overflow in single cigar"; return LC_ENOMEM; }
		res.aln.resize(1); res.cigars.resize((size_t)h_ctl[8]);
		be.d2h(res.aln.data(), d_aln.p, sizeof(lc_aln_t));
		if (h_ctl[8]) be.d2h(res.cigars.data(), d_cigar.p, 4 * (size_t)h_ctl[8]);
		be.sync();
		return LC_OK;
	}

	int run(int score_delta)
	{
		int rc = run_stages(score_delta);
		if (!rc) rc = run_pack(0, 0);
		return rc;
	}

	// stages 1-6; afterwards n_cand / n_cigar_words are known on the host
	int run_stages(int score_delta)
	{
		int rc = run_front(score_delta);
		if (!rc) rc = run_cigar();
		return rc;
	}

	// stages 1-5 (seeding ... mate rescue); afterwards the regions are final and n_cand is known on the host
	int run_front(int score_delta)
	{
		const int R = b.n_reads;
		b.score_delta = score_delta;
		int rc0 = setup();
		if (rc0) return rc0;
		rc0 = stage_seed();
		if (rc0) return rc0;
		const unsigned long long S = h_ctl[1];
		// ---- stage 2-4 pools, sized exactly from the seed count
		// chain scratch (<= 120 B/seed + 440 B/read) + extension sort keys (8 B/seed + 16 B/read) + unit planning (40 B/chain + 16 B/read)
		b.scratch_cap = 172ull * S + 500ull * R + 4096;
		b.chain_cap = S + 16; b.cseed_cap = S + 16; b.reg_cap = S + 16;
		{
			// a heavy unit has more than ext_heavy_seeds seeds or more than ext_heavy_chains chains (each at least one seed)
			unsigned long long m = 0;
			if (ext_heavy_seeds > 0) m = (unsigned long long)ext_heavy_seeds + 1;
			if (ext_heavy_chains > 0 && (m == 0 || (unsigned long long)ext_heavy_chains + 1 < m)) m = (unsigned long long)ext_heavy_chains + 1;
			b.ext_hvu_cap = m ? S / m + 16 : 16;
		}
		if (!d_scratch.ensure(be, b.scratch_cap) || !d_chain.ensure(be, sizeof(LcChainOut) * b.chain_cap) ||
			!d_cseed.ensure(be, sizeof(LcSeed) * b.cseed_cap) || !d_reg.ensure(be, sizeof(lc_alnreg_t) * b.reg_cap) ||
			!d_chainaux.ensure(be, sizeof(LcChainAux) * b.chain_cap) || !d_units.ensure(be, 4 * b.chain_cap) || !d_hvu.ensure(be, 16 * b.ext_hvu_cap)) return nomem();
		b.scratch = (uint8_t *)d_scratch.p; b.chain = (LcChainOut *)d_chain.p; b.cseed = (LcSeed *)d_cseed.p; b.reg = (lc_alnreg_t *)d_reg.p;
		b.chain_aux = (LcChainAux *)d_chainaux.p; b.ext_units = (uint32_t *)d_units.p; b.ext_units_cap = b.chain_cap;
		b.ext_hvu_in = (uint64_t *)d_hvu.p; b.ext_hvu_out = b.ext_hvu_in + b.ext_hvu_cap;
		if (S && !(b.ix.sa32 || b.ix.sa64)) { be.k_sa(b, S); ++launches; }     // with the dense suffix array the seeding kernels wrote positions already
		be.ev_record(4);
		be.k_chain(b); launches += chain_heavy > 0 ? 2 : 1;      // one thread per read, then a warp per read with very many seeds
		be.ev_record(5);
		launches += be.k_extend(b);             // lanes (one read each); reads with very many chains: plan, lanes (one unit each), a warp per large unit
		be.ev_record(6);
		read_ctl();
		if (launch_broken) { err = "a CUDA kernel launch failed in the chain/extend stages"; return LC_ECUDA; }
		if (h_ctl_ovf) { err = "internal pool overflow in chain/extend (flags " + std::to_string(h_ctl_ovf) + ")"; return LC_ENOMEM; }
		const unsigned long long NREG = h_ctl[10];
		// ---- stage 5: per-pair dedup + rescue. every rescue call adds at most one region: sum of caps <= 3 * regions
		b.fin_cap = 3 * NREG + 64;
		if (!d_fin.ensure(be, sizeof(lc_alnreg_t) * b.fin_cap) || !d_aln.ensure(be, sizeof(lc_aln_t) * b.fin_cap) ||
			!d_cand.ensure(be, 8 * b.fin_cap)) return nomem();
		b.fin = (lc_alnreg_t *)d_fin.p; b.aln = (lc_aln_t *)d_aln.p; b.cand = (uint64_t *)d_cand.p;
		b.sortk_cap = NREG + 64;
		b.sortk = d_sortk.ensure(be, sizeof(LcSortKey) * b.sortk_cap) ? (LcSortKey *)d_sortk.p : nullptr;      // optional: without it the sorts go through handles
		b.regtmp = (b.sortk && prep_heavy && d_regtmp.ensure(be, sizeof(lc_alnreg_t) * b.sortk_cap)) ? (lc_alnreg_t *)d_regtmp.p : nullptr;
		be.k_pair(b, n_pairs); launches += b.prep_heavy_n ? 3 : 2;
		be.ev_record(7);
		read_ctl();
		if (launch_broken) { err = "a CUDA kernel launch failed in the pair stage"; return LC_ECUDA; }
		if (h_ctl_ovf) { err = "internal pool overflow in pair stage"; return LC_ENOMEM; }
		n_cand = h_ctl[7];
		n_rescue_pairs = (unsigned long long)(uint32_t)h_ctl[27];
		return LC_OK;
	}

	// stage 6: CIGARs (lc_cigar.cuh: plan -> register-band kernels -> warp fallback; regrow + rerun on overflow).  Reads the regions, does not change them.
	int run_cigar()
	{
		const unsigned long long NC = n_cand;
		b.cig_jobs_cap = NC + 1;
		if (!d_cigjobs.ensure(be, 8 * (size_t)LC_CIG_NLIST * b.cig_jobs_cap)) return nomem();
		b.cig_jobs = (uint64_t *)d_cigjobs.p;
		b.alnx = nullptr; b.mm_ref = b.mm_read = nullptr; b.mm_cap = 0;
		if (want_alnx) {
			if (!d_alnx.ensure(be, sizeof(lc_alnx_t) * b.fin_cap)) return nomem();
			b.alnx = (lc_alnx_t *)d_alnx.p;
		}
		for (int attempt = 0;; ++attempt) {
			b.cigar_cap = (unsigned long long)(cigar_per_cand * NC) + 1024;
			b.z_cap = std::max<unsigned long long>(z_pool_bytes, (unsigned long long)(z_per_cand * NC));
			if (!d_cigar.ensure(be, 4 * b.cigar_cap) || !d_z.ensure(be, b.z_cap)) return nomem();
			b.cigar = (uint32_t *)d_cigar.p; b.zpool = (uint8_t *)d_z.p;
			if (want_alnx) {
				b.mm_cap = (unsigned long long)(mm_per_cand * NC) + 1024;
				if (!d_mmref.ensure(be, 4 * b.mm_cap) || !d_mmread.ensure(be, 4 * b.mm_cap)) return nomem();
				b.mm_ref = (int32_t *)d_mmref.p; b.mm_read = (int32_t *)d_mmread.p;
			}
			if (NC) {
				be.k_cigar_plan(b, NC); ++launches;
				if (!b.small) launches += be.k_cigar_band(b, NC);      // (small batches: every job is on the fallback list)
				be.k_cigar_fallback(b, NC); ++launches;
				if (want_alnx) { be.k_alnx(b, NC); ++launches; }      // reads the CIGARs: a no-op if the stage above overflowed
			}
			read_ctl();
			if (launch_broken) { err = "a CUDA kernel launch failed in the cigar stage"; return LC_ECUDA; }
			n_cig_fallback = (unsigned long long)(uint32_t)(h_ctl[24 + LC_CIG_FALLBACK / 2] >> (32 * (LC_CIG_FALLBACK & 1)));
			if (!h_ctl_ovf) break;
			if (h_ctl_ovf & ~(uint32_t)(LC_OVF_CIGAR | LC_OVF_Z | LC_OVF_MM)) { err = "internal pool overflow in the cigar stage (flags " + std::to_string(h_ctl_ovf) + ")"; return LC_ENOMEM; }
			if (attempt > 12) { err = "cigar pools keep overflowing"; return LC_ENOMEM; }
			if (h_ctl_ovf & LC_OVF_CIGAR) cigar_per_cand = std::max(cigar_per_cand * 2, (double)h_ctl[8] / std::max<double>(NC, 1) * 1.25);
			if (h_ctl_ovf & LC_OVF_Z) { z_pool_bytes = std::max<size_t>(z_pool_bytes * 2, (size_t)(h_ctl[9] * 1.25)); z_per_cand = std::max(z_per_cand, (double)h_ctl[9] / std::max<double>(NC, 1) * 1.25); }
			if (h_ctl_ovf & LC_OVF_MM) mm_per_cand = std::max(mm_per_cand * 2, (double)h_ctl[12] / std::max<double>(NC, 1) * 1.25);
			be.dzero_words(b.cigar_cur, 1); be.dzero_words(b.z_cur, 1); be.dzero_words(b.mm_cur, 1); be.dzero_words((unsigned long long *)b.overflow, 1);
			be.dzero(b.cig_n, 4 * LC_CIG_NLIST);
		}
		be.ev_record(8);
		n_cigar_words = h_ctl[8]; n_mm = want_alnx ? h_ctl[12] : 0;
		return LC_OK;
	}

	// ---- stage 7: pack into dense arrays (candidate indices offset by reg_base, CIGAR indices by cig_base)
	int run_pack(int64_t reg_base, uint32_t cig_base, uint32_t mm_base = 0)
	{
		int rc = run_pack_regs(reg_base);
		if (!rc) rc = run_pack_aln(cig_base, mm_base);
		return rc;
	}
	// the regions and the per-read offsets: any time after run_front()
	int run_pack_regs(int64_t reg_base)
	{
		const int R = b.n_reads;
		const unsigned long long NC = n_cand;
		if (!d_outoff.ensure(be, 8 * ((size_t)R + 1)) || !d_outoff2.ensure(be, 8 * ((size_t)R + 1)) || !d_outreg.ensure(be, sizeof(lc_alnreg_t) * (NC + 1))) return nomem();
		be.scan_exclusive(b.r_n_fin, (int64_t *)d_outoff.p, R + 1); ++launches;   // r_n_fin[R] == 0, so out[R] = total
		LcPackArgs pa; memset(&pa, 0, sizeof(pa));
		pa.out_off = (const int64_t *)d_outoff.p; pa.out_reg = (lc_alnreg_t *)d_outreg.p; pa.off_final = (int64_t *)d_outoff2.p; pa.reg_base = reg_base;
		be.k_pack(b, pa); ++launches;
		return LC_OK;
	}
	// the alignments (and stage 6b's records): after run_cigar()
	int run_pack_aln(uint32_t cig_base, uint32_t mm_base = 0)
	{
		const unsigned long long NC = n_cand;
		if (!d_outaln.ensure(be, sizeof(lc_aln_t) * (NC + 1))) return nomem();
		LcPackArgs pa; memset(&pa, 0, sizeof(pa));
		pa.out_off = (const int64_t *)d_outoff.p; pa.out_aln = (lc_aln_t *)d_outaln.p; pa.cig_base = cig_base; pa.mm_base = mm_base;
		if (b.alnx) { if (!d_outalnx.ensure(be, sizeof(lc_alnx_t) * (NC + 1))) return nomem(); pa.out_alnx = (lc_alnx_t *)d_outalnx.p; }
		be.k_pack(b, pa); ++launches;
		be.ev_record(9);
		return LC_OK;
	}

	// D2H of the packed arrays into caller-provided host memory (pinned in lc_api.cu); n_off = how many offsets to copy
	// (n_reads for an inner chunk, n_reads + 1 for the last).  Asynchronous: call fetch_done() after be.sync().
	void fetch_into(int64_t *h_off, size_t n_off, lc_alnreg_t *h_regs, lc_aln_t *h_aln, uint32_t *h_cig,
					lc_alnx_t *h_alnx = nullptr, int32_t *h_mmref = nullptr, int32_t *h_mmread = nullptr)
	{
		be.d2h(h_off, d_outoff2.p, 8 * n_off);
		if (n_cand) be.d2h(h_regs, d_outreg.p, sizeof(lc_alnreg_t) * n_cand);
		fetch_aln_into(h_aln, h_cig, h_alnx, h_mmref, h_mmread);
	}
	// The two halves of fetch_into for a caller that packs in two steps.  The regions' copy goes to the backend's copy stream, ordered
	// after run_pack_regs(): it runs under the CIGAR kernels.  be.copy_sync() before the host reads them.
	void fetch_regs_into_side(int64_t *h_off, size_t n_off, lc_alnreg_t *h_regs)
	{
		be.copy_fork();
		be.d2h_copy(h_off, d_outoff2.p, 8 * n_off);
		if (n_cand) be.d2h_copy(h_regs, d_outreg.p, sizeof(lc_alnreg_t) * n_cand);
	}
	void fetch_aln_into(lc_aln_t *h_aln, uint32_t *h_cig, lc_alnx_t *h_alnx = nullptr, int32_t *h_mmref = nullptr, int32_t *h_mmread = nullptr)
	{
		if (b.alnx && h_alnx) {
			if (n_cand) be.d2h(h_alnx, d_outalnx.p, sizeof(lc_alnx_t) * n_cand);
			if (n_mm) { be.d2h(h_mmref, d_mmref.p, 4 * n_mm); be.d2h(h_mmread, d_mmread.p, 4 * n_mm); }
		}
		if (n_cand) be.d2h(h_aln, d_outaln.p, sizeof(lc_aln_t) * n_cand);
		if (n_cigar_words) be.d2h(h_cig, d_cigar.p, 4 * n_cigar_words);
		if (counting) be.d2h(&h_counters, (uint8_t *)d_ctl.p + 32 * 8, sizeof(LcCounters));
		be.ev_record(10);
	}
	void fetch_done(float ms[10], LcCounters &cnt)
	{
		ms[0] = be.ev_ms(0, 1); ms[1] = be.ev_ms(2, 3); ms[2] = be.ev_ms(3, 4); ms[3] = be.ev_ms(4, 5);
		ms[4] = be.ev_ms(5, 6); ms[5] = 0; ms[6] = be.ev_ms(6, 7); ms[7] = be.ev_ms(7, 8);
		ms[8] = be.ev_ms(9, 10); ms[9] = be.ev_ms(2, 9);
		if (counting) cnt = h_counters; else memset(&cnt, 0, sizeof(cnt));
	}

	int fetch(LcHostResult &res)
	{
		const int R = b.n_reads;
		res.reg_off.resize(R + 1);
		res.regs.resize(n_cand); res.aln.resize(n_cand); res.cigars.resize(n_cigar_words);
		be.d2h(res.reg_off.data(), d_outoff2.p, 8 * ((size_t)R + 1));
		if (n_cand) { be.d2h(res.regs.data(), d_outreg.p, sizeof(lc_alnreg_t) * n_cand); be.d2h(res.aln.data(), d_outaln.p, sizeof(lc_aln_t) * n_cand); }
		if (n_cigar_words) be.d2h(res.cigars.data(), d_cigar.p, 4 * n_cigar_words);
		res.alnx.resize(b.alnx ? n_cand : 0); res.mm_ref.resize(b.alnx ? n_mm : 0); res.mm_read.resize(b.alnx ? n_mm : 0);
		if (b.alnx && n_cand) be.d2h(res.alnx.data(), d_outalnx.p, sizeof(lc_alnx_t) * n_cand);
		if (b.alnx && n_mm) { be.d2h(res.mm_ref.data(), d_mmref.p, 4 * n_mm); be.d2h(res.mm_read.data(), d_mmread.p, 4 * n_mm); }
		be.ev_record(10);
		be.sync();
		if ((unsigned long long)res.reg_off[R] != n_cand) { err = "result size mismatch"; return LC_ECUDA; }
		res.ms[0] = be.ev_ms(0, 1); res.ms[1] = be.ev_ms(2, 3); res.ms[2] = be.ev_ms(3, 4); res.ms[3] = be.ev_ms(4, 5);
		res.ms[4] = be.ev_ms(5, 6); res.ms[5] = 0; res.ms[6] = be.ev_ms(6, 7); res.ms[7] = be.ev_ms(7, 8);
		res.ms[8] = be.ev_ms(9, 10); res.ms[9] = be.ev_ms(2, 9);
		res.launches = launches;
		memset(&res.counters, 0, sizeof(res.counters));
		if (counting) be.d2h(&res.counters, (uint8_t *)d_ctl.p + 32 * 8, sizeof(LcCounters)), be.sync();
		return LC_OK;
	}

	void release()
	{
		LcBuf<BE> *all[] = { &d_seq, &d_off, &d_perread, &d_ctl, &d_wintv, &d_wint, &d_wu64, &d_wz, &d_intv, &d_seed, &d_scratch, &d_chain,
							 &d_cseed, &d_reg, &d_fin, &d_aln, &d_cand, &d_cigar, &d_z, &d_outoff, &d_outoff2, &d_outreg, &d_outaln, &d_nfin, &d_wpk, &d_redo, &d_pkpool, &d_cigjobs, &d_wpk2, &d_redo2, &d_redo3, &d_wl,
							 &d_alnx, &d_mmref, &d_mmread, &d_outalnx, &d_raw, &d_chainaux, &d_units, &d_hvu, &d_sortk, &d_regtmp };
		for (auto *x : all) x->release(be);
	}

	int64_t launches = 0;
	int n_redo = 0, n_redo_pass1 = 0;
	unsigned long long n_cig_fallback = 0, n_rescue_pairs = 0;
	unsigned long long n_cand = 0, n_cigar_words = 0, n_mm = 0;

private:
	LcBuf<BE> d_seq, d_off, d_perread, d_ctl, d_wintv, d_wint, d_wu64, d_wz, d_intv, d_seed, d_scratch, d_chain, d_cseed, d_reg, d_fin,
		d_aln, d_cand, d_cigar, d_z, d_outoff, d_outoff2, d_outreg, d_outaln, d_nfin, d_wpk, d_redo, d_pkpool, d_cigjobs, d_wpk2, d_redo2, d_redo3, d_wl,
		d_alnx, d_mmref, d_mmread, d_outalnx, d_raw, d_chainaux, d_units, d_hvu, d_sortk, d_regtmp;
	unsigned long long h_ctl[32];
	uint32_t h_ctl_ovf = 0;
	LcCounters h_counters;
	int nomem() { err = "device allocation failed"; return LC_ENOMEM; }
	void read_ctl()
	{
		be.d2h(h_ctl, d_ctl.p, sizeof(h_ctl));
		be.sync();
		h_ctl_ovf = (uint32_t)h_ctl[16];
		if (be.launch_failed()) {      // a kernel was not launched: its outputs are garbage — make every caller stop (they all test the overflow word)
			h_ctl_ovf = 0xffffffffu; launch_broken = true;
			for (auto &w : h_ctl) w = 0;
		}
	}
	bool launch_broken = false;
};

