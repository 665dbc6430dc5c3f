// hostsim.cpp — TEST-ONLY serial execution of the product's per-read stage bodies on the CPU.
//
// The CUDA kernels of libcuda_lariat.so are thin grids around the __host__ __device__ stage bodies in
// lariat_b200/csrc/lc_mem.cuh.  This harness compiles the same bodies with the host compiler and runs them
// one "thread" after another through the same pipeline orchestration (lc_pipeline.hpp), so the logic, the
// pool arithmetic and the result packing can be checked against the oracle on a machine without a GPU.
// It is NOT part of the product: it is built only by tests (tests/test_hostsim.py) into tests/_build/,
// never linked into or loaded by libcuda_lariat.so, and the product has no CPU execution path.
//
// usage: hostsim <index-prefix> <pairs.bin> <out.bin> [-d score_delta] [-w workers] [-n max_pairs]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../lariat_b200/csrc/lc_pipeline.hpp"
#include "../../lariat_b200/csrc/lc_extsm.cuh"

struct HostBackend {
	int workers = 3;
	int n_workers() const { return workers; }
	int n_redo_workers() const { return workers; }
	int seed_lanes() const { return 1; }
	void *dmalloc(size_t n) { return malloc(n ? n : 1); }
	void dfree(void *p) { free(p); }
	void dzero(void *p, size_t n) { memset(p, 0, n); }
	void dfill(void *p, int byte, size_t n) { memset(p, byte, n); }
	void k_iota(int32_t *p, int n) { for (int i = 0; i < n; ++i) p[i] = i; }
	void dzero_words(unsigned long long *p, size_t n) { memset(p, 0, 8 * n); }
	void h2d(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void d2h(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void d2d(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void sync() {}
	void copy_fork() {}
	void d2h_copy(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void copy_sync() {}
	void ev_record(int) {}
	float ev_ms(int, int) { return 0.f; }
	void k_convert(uint8_t *p, unsigned long long n, bool ascii) { for (unsigned long long i = 0; i < n; ++i) p[i] = lc_nt4(p[i], ascii); }
	// each stage: items are dealt to `workers` scratch slots round-robin, like the grid-stride loops of the kernels
	// the seeding state machine with a one-lane group (lc_seed.cuh), then the redo list through the one-thread-per-read body
	int seed_pass_lanes(const LcBatch &, int) const { return 1; }
	bool launch_failed() const { return false; }
	void k_seed(const LcBatch &b, const int32_t *list, int n_items)
	{
		std::vector<uint32_t> lst((size_t)LC_SEED_ENT_WORDS * b.lst_cap + 4), q((size_t)lc_seed_q_words(b.lmax) + 4);
		if (b.force_wide || b.ix.seq_len >= 0xffffffffull) lc_seed_run_scalar<uint64_t>(b, lst.data(), b.lst_cap, q.data(), b.w_pk, b.pk_cap, idle_ticks, list, n_items);
		else lc_seed_run_scalar<uint32_t>(b, lst.data(), b.lst_cap, q.data(), b.w_pk, b.pk_cap, idle_ticks, list, n_items);
	}
	int idle_ticks = 0;      // ticks a lane spends waiting for the warp's next housekeeping round (device kernels), simulated
	void k_seed_finish(const LcBatch &b)
	{
		LcCoopScalar co;
		for (int r = 0; r < b.n_reads; ++r) {
			if (b.force_wide || b.ix.seq_len >= 0xffffffffull) lc_seed_finish<uint64_t>(b, co, r); else lc_seed_finish<uint32_t>(b, co, r);
		}
	}
	void k_seed_redo(const LcBatch &b, int n_redo) { for (int t = 0; t < n_redo; ++t) lc_stage_seed(b, b.redo[t], t % workers); }
	void k_sa(const LcBatch &b, unsigned long long n) { for (unsigned long long g = 0; g < n; ++g) lc_stage_sa(b, g); }
	void k_chain(const LcBatch &b)
	{
		for (int r = 0; r < b.n_reads; ++r) lc_stage_chain(b, r);
		for (unsigned int t = 0; t < b.chain_heavy_n[0]; ++t) lc_stage_chain_heavy_t(b, b.chain_list[t], LcChainExec1());      // the listed reads' kernel
	}
	// the extension state machine with stride-1 rows (lc_extsm.cuh); -X runs the older one-call-per-read body instead
	bool old_extend = false;
	int k_extend(const LcBatch &b)
	{
		if (old_extend || b.small) { for (int r = 0; r < b.n_reads; ++r) lc_stage_extend(b, r, -1, r % workers); return 1; }
		std::vector<uint32_t> he((size_t)lc_ext_he_cols(b.lmax)), qw((size_t)lc_ext_q_words(b.lmax));
		const bool c8 = lc_ext_cells8_ok(b) && !wide_cells;
		lc_ext_run_scalar(b, he.data(), qw.data(), c8, idle_ticks);
		if (!*b.ext_heavy_n) return 1;
		for (unsigned int t = *b.ext_heavy_n; t-- > 0;) lc_ext_plan(b, (int)(uint32_t)b.wl_in[t]);      // reversed: order must not matter
		n_units += lc_ext_n_units(b); n_heavy_units += b.ext_hvu_n[0];
		lc_ext_run_scalar(b, he.data(), qw.data(), c8, idle_ticks, true);                                // the units' lane kernel
		for (unsigned int t = b.ext_hvu_n[0]; t-- > 0;) lc_stage_extend(b, -1, (long long)(uint32_t)b.ext_hvu_in[t], (int)(t % workers));   // the large units' kernel
		return 4;
	}
	unsigned long long n_units = 0, n_heavy_units = 0;
	bool wide_cells = false;     // -H: 16-bit H/E cells even when the batch would fit the byte packing
	void k_pair(const LcBatch &b, int n_pairs)
	{
		for (int p = n_pairs - 1; p >= 0; --p) if (!lc_pair_prepare_defer(b, p)) lc_stage_pair_prepare(b, p, p % workers);   // reversed: order must not matter
		if (b.prep_heavy_n) for (unsigned int t = 0; t < b.prep_heavy_n[0]; ++t) lc_stage_pair_prepare(b, (int)(uint32_t)b.wl_out[t], (int)(t % workers));   // the listed pairs' launch
		for (unsigned int t = 0; t < *b.resc_n; ++t) lc_stage_pair_rescue(b, (int)(uint32_t)b.wl_in[t], (int)(b.wl_in[t] >> 48) & 3, (int)(t % workers));
	}
	void k_cigar(const LcBatch &b, const uint64_t *list, unsigned long long n, const unsigned int *n_dev)
	{
		if (n_dev) n = *n_dev;
		for (unsigned long long c = 0; c < n; ++c) lc_stage_cigar(b, list[c], (int)(c % workers));
	}
	// the register-band CIGAR path (lc_cigar.cuh); -G sends every candidate through the warp-kernel body instead
	bool old_cigar = false;
	void k_cigar_plan(const LcBatch &b, unsigned long long n)
	{
		for (unsigned long long c = 0; c < n; ++c) { if (old_cigar) lc_cig_push_job(b, LC_CIG_FALLBACK, b.cand[c]); else lc_cig_plan(b, b.cand[c]); }
	}
	int k_cigar_band(const LcBatch &b, unsigned long long)
	{
		for (int cls = 0; cls < LC_CIG_NCLASS; ++cls) {
			const uint64_t *jobs = b.cig_jobs + (size_t)cls * b.cig_jobs_cap;
			for (unsigned int t = 0; t < b.cig_n[cls]; ++t) {
				if (cls == 0) lc_cig_band<24>(b, jobs[t]); else if (cls == 1) lc_cig_band<32>(b, jobs[t]);
				else if (cls == 2) lc_cig_band<48>(b, jobs[t]); else if (cls == 3) lc_cig_band<64>(b, jobs[t]); else lc_cig_band<80>(b, jobs[t]);
			}
			n_band[cls] += b.cig_n[cls];
		}
		return LC_CIG_NCLASS;
	}
	unsigned long long n_band[LC_CIG_NCLASS] = {};
	void k_cigar_fallback(const LcBatch &b, unsigned long long) { k_cigar(b, b.cig_jobs + (size_t)LC_CIG_FALLBACK * b.cig_jobs_cap, 0, b.cig_n + LC_CIG_FALLBACK); }
	void k_alnx(const LcBatch &b, unsigned long long n) { for (unsigned long long c = n; c-- > 0;) lc_stage_alnx(b, b.cand[c]); }
	void k_pack(const LcBatch &b, const LcPackArgs &pa) { for (int r = 0; r <= b.n_reads; ++r) lc_stage_pack(b, pa, r); }
	void k_rebase(int64_t *off, int n, int64_t off0) { for (int i = 0; i < n; ++i) off[i] -= off0; }
	void scan_exclusive(const int32_t *in, int64_t *out, int n) { int64_t s = 0; for (int i = 0; i < n; ++i) { out[i] = s; s += in[i]; } }
};

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: hostsim <idx> <pairs.bin> <out.bin> [-d delta] [-w workers] [-n max_pairs] [-c]\n"); return 2; }
	const char *xpath = 0;
	int chain_heavy = -1;
	int delta = 25, maxp = -1, count = 0, lst_cap = -1, pk_cap = -1, lst_cap2 = -1, pk_cap2 = -1, heavy_seeds = -1, dense = 1, wide = 0, chunk = 0, small_reads = 0, force_coop = 0, late_regs = 0;
	HostBackend be;
	for (int i = 4; i < argc; ++i) {
		if (!strcmp(argv[i], "-d") && i + 1 < argc) delta = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-w") && i + 1 < argc) be.workers = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-n") && i + 1 < argc) maxp = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-c")) count = 1;
		else if (!strcmp(argv[i], "-C") && i + 1 < argc) lst_cap = atoi(argv[++i]);     // shared-memory list capacity of the seeding state machine
		else if (!strcmp(argv[i], "-M") && i + 1 < argc) pk_cap = atoi(argv[++i]);      // per-group seed list capacity
		else if (!strcmp(argv[i], "-D") && i + 1 < argc) lst_cap2 = atoi(argv[++i]);    // the same two capacities for the second pass (overflow there -> lc_stage_seed)
		else if (!strcmp(argv[i], "-E") && i + 1 < argc) pk_cap2 = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-W")) wide = 1;                                      // 64-bit interval arithmetic (as for indexes >= 2^32 symbols)
		else if (!strcmp(argv[i], "-I") && i + 1 < argc) be.idle_ticks = atoi(argv[++i]);  // idle ticks before every housekeeping step of the state machines
		else if (!strcmp(argv[i], "-J") && i + 1 < argc) heavy_seeds = atoi(argv[++i]);   // extension: reads above this many chained seeds take the per-read body
		else if (!strcmp(argv[i], "-Q") && i + 1 < argc) chain_heavy = atoi(argv[++i]);   // chaining: reads above this many seeds take the listed-read body (0 = never)
		else if (!strcmp(argv[i], "-Z") && i + 1 < argc) small_reads = atoi(argv[++i]);   // batches up to this many reads take the low-latency (per-read) bodies
		else if (!strcmp(argv[i], "-P")) force_coop = 1;                                // stage 5a: the cooperative dedup body (one lane here)
		else if (!strcmp(argv[i], "-X")) be.old_extend = true;
		else if (!strcmp(argv[i], "-G")) be.old_cigar = true;
		else if (!strcmp(argv[i], "-H")) be.wide_cells = true;
		else if (!strcmp(argv[i], "-L")) late_regs = 1;                                 // with -K: every second chunk takes the "CIGAR stage first" order
		else if (!strcmp(argv[i], "-K") && i + 1 < argc) chunk = atoi(argv[++i]);       // run the batch as chunks of this many pairs (the way lc_align_packed splits large calls)
		else if (!strcmp(argv[i], "-x") && i + 1 < argc) xpath = argv[++i];             // stage 6b on: write the GetAlignments records (LRX1) here
		else if (!strcmp(argv[i], "-S")) dense = 0;                                     // sampled suffix array + LF walks instead of the dense one
	}
	LcHostIndex hix; std::string err;
	if (!lc_load_index(argv[1], hix, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
	FILE *f = fopen(argv[2], "rb");
	if (!f) { perror(argv[2]); return 1; }
	char magic[4]; int32_t n;
	if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "LRP1", 4) || fread(&n, 4, 1, f) != 1) return 1;
	std::vector<int32_t> len(2 * (size_t)n);
	if (fread(len.data(), 4, len.size(), f) != len.size()) return 1;
	if (maxp >= 0 && maxp < n) n = maxp;
	std::vector<int64_t> off(2 * (size_t)n + 1, 0);
	for (int i = 0; i < 2 * n; ++i) off[i + 1] = off[i] + len[i];
	std::vector<uint8_t> bases((size_t)off[2 * n] + 1);
	if (fread(bases.data(), 1, (size_t)off[2 * n], f) != (size_t)off[2 * n]) return 1;
	fclose(f);

	LcIndex ix;
	LcHostDevIndex dix;
	if (!lc_host_dev_index(hix, dix, ix, dense != 0, err, wide != 0))  /* -W: 64-bit intervals AND the 64-bit dense suffix array */ { fprintf(stderr, "%s\n", err.c_str()); return 1; }
	lc_opt_t opt; lc_default_opt(&opt);
	LcPipeline<HostBackend> pl(be);
	pl.counting = count;
	if (lst_cap > 0) pl.lst_cap = lst_cap;
	if (pk_cap > 0) pl.pk_cap = pk_cap;
	if (lst_cap2 > 0) { pl.lst_cap2 = lst_cap2; pl.lst_cap3 = lst_cap2 + 2; }      // (third pass: barely larger, so that the last fallback still gets reads)
	if (heavy_seeds >= 0) pl.ext_heavy_seeds = heavy_seeds;
	if (chain_heavy >= 0) pl.chain_heavy = chain_heavy;
	pl.small_reads = small_reads; pl.force_coop = force_coop != 0;
	if (pk_cap2 > 0) { pl.pk_cap2 = pk_cap2; pl.pk_cap3 = pk_cap2 + 2; }
	pl.force_wide = wide != 0;
	pl.want_alnx = xpath != 0;
	pl.set_index(ix, opt);
	LcHostResult res;
	int rc = 0;
	if (chunk > 0 && chunk < n) {
		// chunked form: slices of the packed arrays, results appended at the running candidate / CIGAR-word totals
		res.reg_off.assign(2 * (size_t)n + 1, 0);
		int64_t reg_base = 0; uint32_t cig_base = 0, mm_base = 0;
		std::vector<uint8_t> raw(bases);          // every chunk converts its own slice in place on its own copy
		for (int p0 = 0; p0 < n && !rc; p0 += chunk) {
			const int np = std::min(chunk, n - p0);
			rc = pl.upload(np, raw.data(), off.data() + 2 * (size_t)p0, true);
			// the orders of lc_api.cu's lane_worker: the regions are packed and fetched before the CIGAR stage runs, or — when the chunk
			// before is not through its pair stage yet (-L: every second chunk here) — after it
			const bool late = late_regs && ((p0 / chunk) & 1);
			if (!rc) rc = pl.run_front(delta);
			if (!rc && late) rc = pl.run_cigar();
			if (!rc) rc = pl.run_pack_regs(reg_base);
			if (rc) break;
			res.regs.resize((size_t)reg_base + pl.n_cand); res.aln.resize((size_t)reg_base + pl.n_cand);
			pl.fetch_regs_into_side(res.reg_off.data() + 2 * (size_t)p0, 2 * (size_t)np + 1, res.regs.data() + reg_base);
			if (!late) rc = pl.run_cigar();
			if (!rc) rc = pl.run_pack_aln(cig_base, mm_base);
			if (rc) break;
			res.cigars.resize((size_t)cig_base + pl.n_cigar_words);
			if (xpath) { res.alnx.resize((size_t)reg_base + pl.n_cand); res.mm_ref.resize((size_t)mm_base + pl.n_mm + 1); res.mm_read.resize((size_t)mm_base + pl.n_mm + 1); }
			pl.fetch_aln_into(res.aln.data() + reg_base, res.cigars.data() + cig_base,
						  xpath ? res.alnx.data() + reg_base : nullptr, xpath ? res.mm_ref.data() + mm_base : nullptr, xpath ? res.mm_read.data() + mm_base : nullptr);
			mm_base += (uint32_t)pl.n_mm;
			float ms[10]; LcCounters cn; pl.fetch_done(ms, cn);
			res.launches += pl.launches;
			reg_base += (int64_t)pl.n_cand; cig_base += (uint32_t)pl.n_cigar_words;
		}
	} else {
		rc = pl.upload(n, bases.data(), off.data(), true);
		if (!rc) rc = pl.run(delta);
		if (!rc) rc = pl.fetch(res);
	}
	if (rc) { fprintf(stderr, "pipeline error %d: %s\n", rc, pl.err.c_str()); return 1; }
	FILE *o = fopen(argv[3], "wb");
	fwrite("LRO1", 1, 4, o); fwrite(&n, 4, 1, o);
	for (int r = 0; r < 2 * n; ++r) {
		int32_t k = (int32_t)(res.reg_off[r + 1] - res.reg_off[r]);
		fwrite(&k, 4, 1, o);
		if (k) fwrite(&res.regs[res.reg_off[r]], sizeof(lc_alnreg_t), k, o);
		for (int i = 0; i < k; ++i) {
			const lc_aln_t &a = res.aln[res.reg_off[r] + i];
			int64_t pos = a.pos; int32_t h[4] = { a.rid, a.is_rev, a.NM, a.n_cigar };
			fwrite(&pos, 8, 1, o); fwrite(h, 4, 4, o);
			if (a.n_cigar) fwrite(&res.cigars[a.cigar_off], 4, a.n_cigar, o);
		}
	}
	fclose(o);
	if (xpath) {
		FILE *x = fopen(xpath, "wb");
		fwrite("LRX1", 1, 4, x); fwrite(&n, 4, 1, x);
		for (int r = 0; r < 2 * n; ++r) {
			int32_t k = (int32_t)(res.reg_off[r + 1] - res.reg_off[r]);
			fwrite(&k, 4, 1, x);
			for (int i = 0; i < k; ++i) {
				const lc_alnx_t &a = res.alnx[res.reg_off[r] + i];
				int32_t h[7] = { a.matches, a.mismatches, a.indels, a.indel_length, a.soft_clipped, a.soft_clipped_length, a.n_mm };
				fwrite(&a.pos, 8, 1, x); fwrite(&a.aend, 8, 1, x); fwrite(h, 4, 7, x);
				if (a.n_mm) { fwrite(&res.mm_ref[a.mm_off], 4, a.n_mm, x); fwrite(&res.mm_read[a.mm_off], 4, a.n_mm, x); }
			}
		}
		fclose(x);
	}
	if (count) {
		const LcCounters &c = res.counters;
		fprintf(stderr, "{\"occ_blocks\": %llu, \"extend_calls\": %llu, \"sa_calls\": %llu, \"lf_steps\": %llu, \"ext_calls\": %llu, \"ext_cells\": %llu, "
				"\"sw_calls\": %llu, \"sw_cells\": %llu, \"glb_calls\": %llu, \"glb_cells\": %llu}\n", c.occ_blocks, c.extend_calls, c.sa_calls, c.lf_steps,
				c.ext_calls, c.ext_cells, c.sw_calls, c.sw_cells, c.glb_calls, c.glb_cells);
	}
	fprintf(stderr, "hostsim: %d pairs, %llu candidates, %lld launches, %d reads redone; cigar jobs by band class %llu/%llu/%llu/%llu/%llu, fallback %llu; %llu pairs to the rescue kernel; %llu + %llu extension units\n", n,
			(unsigned long long)res.regs.size(), (long long)res.launches, pl.n_redo, be.n_band[0], be.n_band[1], be.n_band[2], be.n_band[3], be.n_band[4], pl.n_cig_fallback, pl.n_rescue_pairs, be.n_units, be.n_heavy_units);
#ifdef LC_STATS
	fprintf(stderr, "rescue: %llu SW calls (mean list %.1f, max %llu): %llu skipped re-sorts, %llu one-pass inserts, %llu inserts with a hazard, %llu full calls on unfixed lists\n",
			lc_stats.sw, lc_stats.sw ? (double)lc_stats.n_sum / lc_stats.sw : 0., lc_stats.nmax, lc_stats.skipped, lc_stats.onepass, lc_stats.hazard, lc_stats.full_nofix);
	fprintf(stderr, "heavy extension: %llu reads, %llu chains (max %llu), %llu seeds, %llu regions, %llu containment tests\n", lc_stats.hv_reads, lc_stats.hv_chains,
			lc_stats.hv_maxchains, lc_stats.hv_seeds, lc_stats.hv_regs, lc_stats.hv_scan);
#endif
	return 0;
}
