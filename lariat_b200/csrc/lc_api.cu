// lc_api.cu — the extern "C" boundary of libcuda_lariat.so (declared in include/cuda_lariat.h).
// Index load + upload to HBM, the batched alignment call, the split upload/run/fetch form used for timing,
// single-read source-compatible forms, and the kernel-level entry points of the microbenchmark config.
#include <cuda_runtime.h>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>
#include <stdlib.h>
#include "lc_backend.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
void lc_api_set_error(const std::string &msg) { g_err = msg; }      // for the other translation units of the library (lc_index_build.cu)

// A growable pinned host buffer (cudaHostAlloc): D2H into it runs at PCIe speed and overlaps with kernels of the other lane.
struct LcPinned {
	void *p = nullptr; size_t cap = 0;
	bool ensure(size_t bytes, size_t keep)
	{
		if (bytes <= cap) return true;
		size_t want = bytes + bytes / 4 + 4096;
		void *np = nullptr;
		if (cudaHostAlloc(&np, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
		if (p && keep) memcpy(np, p, keep < cap ? keep : cap);
		if (p) cudaFreeHost(p);
		p = np; cap = want;
		return true;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// A second stream + pool set: large calls are cut into chunks of barcode groups that alternate between two lanes, so that one
// chunk's H2D / D2H and kernel tails overlap the other's kernels.
struct LcLane {
	LcCudaBackend be;
	LcPipeline<LcCudaBackend> pl;
	explicit LcLane(int dev) : be(dev), pl(be) {}
};

struct lc_ctx {
	int device = 0;
	LcHostIndex hix;
	void *d_bwt = nullptr, *d_sa = nullptr, *d_pac = nullptr, *d_anns = nullptr, *d_sa_dense = nullptr;
	LcIndex ix;
	lc_opt_t opt;
	LcCudaBackend be;
	LcPipeline<LcCudaBackend> pl;
	LcHostResult hres;
	lc_result_t res;
	std::mutex mu;
	lc_contigs_t contigs;
	std::vector<const char *> name_ptrs;
	std::vector<int32_t> lens;
	std::vector<int64_t> offsets;
	std::vector<uint8_t> stage;
	std::vector<int64_t> stage_off;
	bool have_batch = false;
	int last_delta = 25;
	// results of the batched calls: pinned host arrays, valid until the next call on this ctx
	LcPinned h_off, h_regs, h_aln, h_cig;
	std::vector<std::unique_ptr<LcLane> > lanes;      // lanes 2..n of chunked calls (lane 1 is be/pl above): own stream, pools and host thread
	int max_lanes = 2;              // LC_LANES
	int chunk_pairs = 500000;       // LC_CHUNK_PAIRS; calls of more than 1.5 chunks run chunked over two lanes, 0 = never
	explicit lc_ctx(int dev) : device(dev), be(dev), pl(be) {}
};

struct DevTmp {
	std::vector<void *> ptrs;
	~DevTmp() { for (void *p : ptrs) cudaFree(p); }
	template <typename T> T *up(const T *h, size_t n) { void *p = nullptr; cudaMalloc(&p, sizeof(T) * (n ? n : 1)); ptrs.push_back(p); if (h && n) cudaMemcpy(p, h, sizeof(T) * n, cudaMemcpyHostToDevice); return (T *)p; }
	template <typename T> T *alloc(size_t n) { return up<T>(nullptr, n); }
};

static LcScore score_of(const lc_opt_t &o) { LcScore s; s.a = o.a; s.b = o.b; s.o_del = o.o_del; s.e_del = o.e_del; s.o_ins = o.o_ins; s.e_ins = o.e_ins; return s; }


__global__ void matesw_single_kernel(const LcBatch b, const lc_alnreg_t a, int l_ms, const uint8_t *ms, lc_alnreg_t *ma, int *n_ma, int *ret)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		LcSwScratch ws;
		ws.H = b.w_int; ws.E = b.w_int + (b.lmax + 18); ws.Hmax = b.w_int + 2 * (b.lmax + 18); ws.b = b.w_u64; ws.z_unused = 0;
		int n = *n_ma;
		*ret = lc_matesw(b, a, l_ms, ms, ma, n, ws);
		*n_ma = n;
	}
}


extern "C" {

const char *lc_last_error(void) { return g_err.c_str(); }
const char *lc_build_info(void) { return "libcuda_lariat 0.1.0 (sm_100a; CUDA kernels: smem, sa_lookup, chain, extend, matesw, global_cigar)"; }
void lc_opt_init(lc_opt_t *opt) { lc_default_opt(opt); }

lc_ctx *lc_init(const char *idx_prefix, int device, const lc_opt_t *opt)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); fail(LC_ENODEV, "no CUDA device: libcuda_lariat has no CPU path"); return nullptr; }
	if (device < 0 || device >= ndev) { fail(LC_ENODEV, "bad device ordinal"); return nullptr; }
	lc_opt_t o;
	if (opt) o = *opt; else lc_default_opt(&o);
	{   // the scoring matrix must be bwa_fill_scmat(a, b) (B/bwa.c:109-118): kernels compute it arithmetically
		lc_opt_t chk = o;
		for (int i = 0, k = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) chk.mat[k++] = (i == 4 || j == 4) ? -1 : (i == j ? o.a : -o.b);
		if (memcmp(chk.mat, o.mat, 25) != 0) { fail(LC_EARG, "opt->mat must equal bwa_fill_scmat(a, b)"); return nullptr; }
		if (o.e_del <= 0 || o.e_ins <= 0 || o.a <= 0 || o.b < 0 || o.min_seed_len < 1 || o.max_occ < 1) { fail(LC_EARG, "unsupported scoring options"); return nullptr; }
		// the extension / rescue / CIGAR kernels keep H and E in packed 16-bit halves (DPX s16x2): every score and every
		// score-minus-penalty of a read of up to LC_MAX_READ_LEN bases must stay inside int16, or results would wrap silently
		const long long pen = std::max<long long>(std::max<long long>((long long)o.o_del + o.e_del, (long long)o.o_ins + o.e_ins), std::max<long long>(o.b, std::max(o.pen_clip5, o.pen_clip3)));
		if (o.o_del < 0 || o.o_ins < 0 || o.pen_clip5 < 0 || o.pen_clip3 < 0 || pen > 4096 || (long long)LC_MAX_READ_LEN * o.a + pen > 32767) {
			fail(LC_EARG, "scoring options out of the supported range: LC_MAX_READ_LEN*a + max penalty must be <= 32767 and penalties <= 4096 (16-bit DP cells)"); return nullptr;
		}
	}
	lc_ctx *c = new lc_ctx(device);
	std::string err;
	if (!lc_load_index(idx_prefix, c->hix, err)) { fail(LC_EINDEX, err); delete c; return nullptr; }
	cudaSetDevice(device);
	LcHostIndex &h = c->hix;
	LcIndex &ix = c->ix;
	if (!lc_index_scalars(h, ix, err)) { fail(LC_EINDEX, err); delete c; return nullptr; }
	// upload the file payloads, then re-lay the BWT on the device as 32-byte blocks (lc_fmindex.cuh) and drop the file layout
	const uint64_t n_blocks = lc_occ_blocks(h.seq_len);
	size_t nb = h.bwt.size() * 4, ns = h.sa.size() * 8, np = h.pac.size(), na = h.anns.size() * sizeof(LcAnn);
	void *d_file = nullptr;
	if (cudaMalloc(&d_file, nb) != cudaSuccess || cudaMalloc(&c->d_bwt, 32 * (size_t)n_blocks) != cudaSuccess || cudaMalloc(&c->d_sa, ns) != cudaSuccess ||
		cudaMalloc(&c->d_pac, np) != cudaSuccess || cudaMalloc(&c->d_anns, na ? na : 1) != cudaSuccess) {
		cudaGetLastError(); if (d_file) cudaFree(d_file); fail(LC_ENOMEM, "cannot allocate the index in HBM"); lc_destroy(c); return nullptr;
	}
	cudaMemcpy(d_file, h.bwt.data(), nb, cudaMemcpyHostToDevice);
	cudaMemcpy(c->d_sa, h.sa.data(), ns, cudaMemcpyHostToDevice);
	cudaMemcpy(c->d_pac, h.pac.data(), np, cudaMemcpyHostToDevice);
	cudaMemcpy(c->d_anns, h.anns.data(), na, cudaMemcpyHostToDevice);
	ix.occ = (const uint4 *)c->d_bwt; ix.sa = (const uint64_t *)c->d_sa;
	ix.pac = (const uint8_t *)c->d_pac; ix.anns = (const LcAnn *)c->d_anns;
	c->be.k_occ_convert((const uint32_t *)d_file, (uint4 *)c->d_bwt, n_blocks, ix);
	{
		std::string e;
		if (c->be.check(e)) { cudaFree(d_file); fail(LC_ECUDA, e); lc_destroy(c); return nullptr; }
	}
	cudaFree(d_file);
	// dense suffix array: one entry per BWT row, built from the sampled one with one LF step per row.  Optional: if HBM is
	// short (or LC_DENSE_SA=0) the kernels fall back to the reference's sampled array + LF walks; results are identical.
	{
		const char *env = getenv("LC_DENSE_SA");
		const bool want = !(env && env[0] == '0');
		const char *fw = getenv("LC_FORCE_WIDE");      // tests: the 64-bit paths (interval arithmetic, sa64) on a small index
		const bool wide = h.seq_len >= 0xffffffffull || (fw && fw[0] == '1');
		const size_t bytes = (size_t)(h.seq_len + 1) * (wide ? 8 : 4);
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		if (want && bytes < free_b / 2 && cudaMalloc(&c->d_sa_dense, bytes) == cudaSuccess) {
			c->be.k_sa_fill(ix, (unsigned long long)h.sa.size(), c->d_sa_dense, wide);
			std::string e;
			if (c->be.check(e)) { fail(LC_ECUDA, e); lc_destroy(c); return nullptr; }
			if (wide) ix.sa64 = (const uint64_t *)c->d_sa_dense; else ix.sa32 = (const uint32_t *)c->d_sa_dense;
		} else cudaGetLastError();
	}
	c->opt = o;
	c->pl.set_index(ix, o);
	// tuning knobs of the seeding state machine (tests use tiny values to force the redo path)
	if (const char *e = getenv("LC_SEED_LST_CAP")) { int v = atoi(e); if (v >= 2 && v <= 2048) c->pl.lst_cap = v; }
	if (const char *e = getenv("LC_SEED_PK_CAP")) { int v = atoi(e); if (v >= 1 && v <= 4096) c->pl.pk_cap = v; }
	if (const char *e = getenv("LC_SEED_LST_CAP2")) { int v = atoi(e); if (v >= 2 && v <= 2048) c->pl.lst_cap2 = v; }
	if (const char *e = getenv("LC_SEED_PK_CAP2")) { int v = atoi(e); if (v >= 1 && v <= 4096) c->pl.pk_cap2 = v; }
	if (const char *e = getenv("LC_SEED_LST_CAP3")) { int v = atoi(e); if (v >= 2 && v <= 2048) c->pl.lst_cap3 = v; }
	if (const char *e = getenv("LC_SEED_PK_CAP3")) { int v = atoi(e); if (v >= 1 && v <= 4096) c->pl.pk_cap3 = v; }
	if (const char *e = getenv("LC_EXT_HEAVY_SEEDS")) { int v = atoi(e); if (v >= 0) c->pl.ext_heavy_seeds = v; }
	if (const char *e = getenv("LC_EXT_HEAVY_CHAINS")) { int v = atoi(e); if (v >= 0) c->pl.ext_heavy_chains = v; }
	if (const char *e = getenv("LC_FORCE_WIDE")) c->pl.force_wide = e[0] == '1';
	if (const char *e = getenv("LC_CHUNK_PAIRS")) { int v = atoi(e); if (v >= 0) c->chunk_pairs = v; }
	if (const char *e = getenv("LC_LANES")) { int v = atoi(e); if (v >= 1 && v <= 8) c->max_lanes = v; }
	// host copies of the big FM-index arrays are not needed any more (pac stays for lc_fetch_seq)
	std::vector<uint32_t>().swap(h.bwt);
	std::vector<uint64_t>().swap(h.sa);
	for (size_t i = 0; i < h.names.size(); ++i) { c->name_ptrs.push_back(h.names[i].c_str()); c->lens.push_back(h.anns[i].len); c->offsets.push_back(h.anns[i].offset); }
	c->contigs.n_seqs = (int32_t)h.names.size(); c->contigs.l_pac = h.l_pac;
	c->contigs.names = c->name_ptrs.data(); c->contigs.lens = c->lens.data(); c->contigs.offsets = c->offsets.data();
	memset(&c->res, 0, sizeof(c->res));
	return c;
}

void lc_destroy(lc_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	c->pl.release();
	c->lanes.clear();
	c->h_off.release(); c->h_regs.release(); c->h_aln.release(); c->h_cig.release();
	cudaFree(c->d_bwt); cudaFree(c->d_sa); cudaFree(c->d_pac); cudaFree(c->d_anns); cudaFree(c->d_sa_dense);
	delete c;
}

const lc_contigs_t *lc_contigs(const lc_ctx *c) { return c ? &c->contigs : nullptr; }
void lc_set_counting(lc_ctx *c, int on) { if (c) c->pl.counting = on != 0; }
int lc_set_chunk_pairs(lc_ctx *c, int n_pairs)
{
	if (!c || n_pairs < 0) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	c->chunk_pairs = n_pairs;
	return LC_OK;
}
int lc_timer_start(lc_ctx *c)
{
	if (!c) return fail(LC_EARG, "null ctx");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	c->be.ev_record(14);
	return LC_OK;
}
int lc_timer_stop(lc_ctx *c, float *ms)
{
	if (!c || !ms) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	c->be.ev_record(15);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	*ms = c->be.ev_ms(14, 15);
	return LC_OK;
}

static void fill_result(lc_ctx *c)
{
	LcHostResult &h = c->hres;
	lc_result_t &r = c->res;
	r.n_reads = (int32_t)h.reg_off.size() - 1;
	r.n_regs = (int64_t)h.regs.size();
	r.reg_off = h.reg_off.data(); r.regs = h.regs.data(); r.aln = h.aln.data(); r.cigars = h.cigars.data();
	r.n_cigar_words = (int64_t)h.cigars.size();
	r.ms_h2d = h.ms[0]; r.ms_seed = h.ms[1]; r.ms_sa = h.ms[2]; r.ms_chain = h.ms[3]; r.ms_extend = h.ms[4]; r.ms_dedup = h.ms[5];
	r.ms_pair = h.ms[6]; r.ms_cigar = h.ms[7]; r.ms_d2h = h.ms[8]; r.ms_total = h.ms[9];
	r.gpu_launches = h.launches;
	const LcCounters &k = h.counters;
	r.cnt_occ_blocks = k.occ_blocks; r.cnt_extend_calls = k.extend_calls; r.cnt_sa_calls = k.sa_calls; r.cnt_lf_steps = k.lf_steps;
	r.cnt_ext_calls = k.ext_calls; r.cnt_ext_cells = k.ext_cells; r.cnt_sw_calls = k.sw_calls; r.cnt_sw_cells = k.sw_cells;
	r.cnt_glb_calls = k.glb_calls; r.cnt_glb_cells = k.glb_cells;
}

static int do_upload(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, bool ascii)
{
	if (!c || n_pairs < 0 || !off) return fail(LC_EARG, "bad arguments");
	cudaSetDevice(c->device);
	int rc = c->pl.upload(n_pairs, bases, off, ascii);
	if (rc) return fail(rc, c->pl.err);
	c->have_batch = true;
	return LC_OK;
}
static int do_run(lc_ctx *c, int score_delta)
{
	if (!c || !c->have_batch) return fail(LC_EARG, "no batch uploaded");
	cudaSetDevice(c->device);
	int rc = c->pl.run(score_delta);
	if (rc) return fail(rc, c->pl.err);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	return LC_OK;
}
static int do_fetch(lc_ctx *c, lc_result_t **out)
{
	if (!c || !c->have_batch) return fail(LC_EARG, "no batch uploaded");
	cudaSetDevice(c->device);
	int rc = c->pl.fetch(c->hres);
	if (rc) return fail(rc, c->pl.err);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	fill_result(c);
	if (out) *out = &c->res;
	return LC_OK;
}

// ---- the batched call: upload -> stages -> pack -> D2H into the pinned result arrays; large calls chunked over two lanes -----
struct LcChunkShared {
	std::mutex mu;
	std::condition_variable cv;
	std::vector<int64_t> cum_reg;       // [n_chunks + 1] candidates before chunk k, valid once ready[k]
	std::vector<int64_t> cum_cig;
	std::vector<char> ready;
	std::shared_mutex res_mu;           // shared: a D2H into the pinned arrays is in flight; exclusive: the arrays are being regrown
	int rc = LC_OK;
	std::string err;
	float ms[10] = {0};
	int64_t launches = 0;
	LcCounters cnt;
	void fail_with(int code, const std::string &e) { std::lock_guard<std::mutex> g(mu); if (rc == LC_OK) { rc = code; err = e; } }
};

static void add_counters(LcCounters &a, const LcCounters &b)
{
	a.occ_blocks += b.occ_blocks; a.extend_calls += b.extend_calls; a.sa_calls += b.sa_calls; a.lf_steps += b.lf_steps;
	a.ext_calls += b.ext_calls; a.ext_cells += b.ext_cells; a.sw_calls += b.sw_calls; a.sw_cells += b.sw_cells;
	a.glb_calls += b.glb_calls; a.glb_cells += b.glb_cells;
}

// chunks first, first + stride, ... of the call on one lane (its own stream, pools and host thread)
static void lane_worker(lc_ctx *c, LcCudaBackend *be, LcPipeline<LcCudaBackend> *pl, LcChunkShared *sh, int first, int stride, int n_chunks, int n_pairs,
						const uint8_t *bases, const int64_t *off, int score_delta)
{
	cudaSetDevice(c->device);
	for (int k = first; k < n_chunks; k += stride) {
		const int p0 = (int)((int64_t)n_pairs * k / n_chunks), p1 = (int)((int64_t)n_pairs * (k + 1) / n_chunks);
		int rc = LC_OK;
		{ std::lock_guard<std::mutex> g(sh->mu); rc = sh->rc; }
		if (!rc) rc = pl->upload(p1 - p0, bases, off + 2 * (size_t)p0, true);
		if (!rc) rc = pl->run_stages(score_delta);
		if (rc) sh->fail_with(rc, pl->err);
		int64_t reg_base, cig_base;
		{   // chunk k's place in the call's arrays = totals of the chunks before it (known when their CIGAR stage is done)
			std::unique_lock<std::mutex> g(sh->mu);
			sh->cv.wait(g, [&] { return sh->ready[k] != 0; });
			reg_base = sh->cum_reg[k]; cig_base = sh->cum_cig[k];
			sh->cum_reg[k + 1] = reg_base + (rc ? 0 : (int64_t)pl->n_cand);
			sh->cum_cig[k + 1] = cig_base + (rc ? 0 : (int64_t)pl->n_cigar_words);
			sh->ready[k + 1] = 1;
			if (!rc && sh->cum_cig[k + 1] > 0xffffffffll) { rc = LC_ENOMEM; if (sh->rc == LC_OK) { sh->rc = rc; sh->err = "more than 2^32 CIGAR words in one call"; } }
		}
		sh->cv.notify_all();
		if (rc) continue;
		rc = pl->run_pack(reg_base, (uint32_t)cig_base);
		if (rc) { sh->fail_with(rc, pl->err); continue; }
		const size_t need_reg = (size_t)(reg_base + (int64_t)pl->n_cand), need_cig = (size_t)(cig_base + (int64_t)pl->n_cigar_words);
		if (need_reg * sizeof(lc_alnreg_t) > c->h_regs.cap || need_reg * sizeof(lc_aln_t) > c->h_aln.cap || need_cig * 4 > c->h_cig.cap) {
			// regrow (rare: capacities persist across calls): extrapolate from the chunks seen so far, wait out in-flight copies, keep what is there
			std::unique_lock<std::shared_mutex> g(sh->res_mu);
			const double f = 1.2 * n_chunks / (double)(k + 1);
			const size_t want_reg = (size_t)(need_reg * f) + 1024, want_cig = (size_t)(need_cig * f) + 1024;
			if (!c->h_regs.ensure(want_reg * sizeof(lc_alnreg_t), c->h_regs.cap) || !c->h_aln.ensure(want_reg * sizeof(lc_aln_t), c->h_aln.cap) ||
				!c->h_cig.ensure(want_cig * 4, c->h_cig.cap)) { sh->fail_with(LC_ENOMEM, "cannot allocate pinned host memory for the results"); continue; }
		}
		float ms[10]; LcCounters cn;
		{
			std::shared_lock<std::shared_mutex> g(sh->res_mu);
			const size_t n_off = 2 * (size_t)(p1 - p0) + (k == n_chunks - 1 ? 1 : 0);
			pl->fetch_into((int64_t *)c->h_off.p + 2 * (size_t)p0, n_off, (lc_alnreg_t *)c->h_regs.p + reg_base, (lc_aln_t *)c->h_aln.p + reg_base,
						   (uint32_t *)c->h_cig.p + cig_base);
			be->sync();
		}
		std::string e;
		if (be->check(e)) { sh->fail_with(LC_ECUDA, e); continue; }
		pl->fetch_done(ms, cn);
		std::lock_guard<std::mutex> g(sh->mu);
		for (int i = 0; i < 10; ++i) sh->ms[i] += ms[i];
		sh->launches += pl->launches;
		add_counters(sh->cnt, cn);
	}
}

static int do_align_chunks(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out, int n_chunks);

// The batch pools scale with the seeds of the batch (~250 bytes per seed between chaining and extension), and repeat-rich reads have
// hundreds of seeds each: a call that cannot get its pools is retried in more, smaller chunks (results do not depend on the
// chunking), and the smaller chunk size is kept for the calls that follow.
static int do_align(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out)
{
	if (!c || n_pairs < 0 || !off || (n_pairs && !bases && off[2 * (size_t)n_pairs] > off[0])) return fail(LC_EARG, "bad arguments");
	cudaSetDevice(c->device);
	int n_chunks = 1;
	if (c->chunk_pairs > 0 && (double)n_pairs > 1.5 * c->chunk_pairs) n_chunks = (n_pairs + c->chunk_pairs - 1) / c->chunk_pairs;
	for (int attempt = 0;; ++attempt) {
		const int rc = do_align_chunks(c, n_pairs, bases, off, score_delta, out, n_chunks);
		if (rc != LC_ENOMEM || attempt >= 8 || n_pairs / n_chunks < 2000) return rc;
		c->pl.release();
		for (auto &l : c->lanes) l->pl.release();
		cudaGetLastError();
		n_chunks = n_chunks < 2 ? 4 : 2 * n_chunks;
		c->chunk_pairs = n_pairs / n_chunks + 1;
	}
}

static int do_align_chunks(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out, int n_chunks)
{
	const int n_lanes = std::max(1, std::min(n_chunks, c->max_lanes));
	while ((int)c->lanes.size() < n_lanes - 1) {
		std::unique_ptr<LcLane> l(new LcLane(c->device));
		l->pl.set_index(c->ix, c->opt);
		l->pl.lst_cap = c->pl.lst_cap; l->pl.pk_cap = c->pl.pk_cap; l->pl.force_wide = c->pl.force_wide;
		l->pl.lst_cap2 = c->pl.lst_cap2; l->pl.pk_cap2 = c->pl.pk_cap2; l->pl.lst_cap3 = c->pl.lst_cap3; l->pl.pk_cap3 = c->pl.pk_cap3;
		l->pl.ext_heavy_seeds = c->pl.ext_heavy_seeds; l->pl.ext_heavy_chains = c->pl.ext_heavy_chains;
		c->lanes.push_back(std::move(l));
	}
	for (auto &l : c->lanes) l->pl.counting = c->pl.counting;
	if (!c->h_off.ensure(8 * (2 * (size_t)n_pairs + 1), 0)) return fail(LC_ENOMEM, "cannot allocate pinned host memory for the results");
	LcChunkShared sh;
	sh.cum_reg.assign(n_chunks + 1, 0); sh.cum_cig.assign(n_chunks + 1, 0); sh.ready.assign(n_chunks + 1, 0); sh.ready[0] = 1;
	memset(&sh.cnt, 0, sizeof(sh.cnt));
	c->be.ev_record(12);
	{
		std::vector<std::thread> th;
		for (int l = 1; l < n_lanes; ++l)
			th.emplace_back(lane_worker, c, &c->lanes[l - 1]->be, &c->lanes[l - 1]->pl, &sh, l, n_lanes, n_chunks, n_pairs, bases, off, score_delta);
		lane_worker(c, &c->be, &c->pl, &sh, 0, n_lanes, n_chunks, n_pairs, bases, off, score_delta);
		for (auto &t : th) t.join();
	}
	c->be.ev_record(13);
	c->be.sync();
	c->have_batch = n_lanes == 1;      // lc_batch_run / lc_batch_fetch continue from a single-lane call only
	if (sh.rc) return fail(sh.rc, sh.err);
	lc_result_t &r = c->res;
	memset(&r, 0, sizeof(r));
	r.n_reads = 2 * n_pairs; r.n_regs = sh.cum_reg[n_chunks]; r.n_cigar_words = sh.cum_cig[n_chunks];
	r.reg_off = (const int64_t *)c->h_off.p; r.regs = (const lc_alnreg_t *)c->h_regs.p; r.aln = (const lc_aln_t *)c->h_aln.p; r.cigars = (const uint32_t *)c->h_cig.p;
	if (r.reg_off[r.n_reads] != r.n_regs) return fail(LC_ECUDA, "result size mismatch");
	// stage times: summed over the chunks (chunks of the two lanes overlap); ms_total: the whole call on the device clock, copies included when chunked
	r.ms_h2d = sh.ms[0]; r.ms_seed = sh.ms[1]; r.ms_sa = sh.ms[2]; r.ms_chain = sh.ms[3]; r.ms_extend = sh.ms[4]; r.ms_dedup = sh.ms[5];
	r.ms_pair = sh.ms[6]; r.ms_cigar = sh.ms[7]; r.ms_d2h = sh.ms[8]; r.ms_total = n_chunks > 1 ? c->be.ev_ms(12, 13) : sh.ms[9];
	r.gpu_launches = sh.launches;
	const LcCounters &k = sh.cnt;
	r.cnt_occ_blocks = k.occ_blocks; r.cnt_extend_calls = k.extend_calls; r.cnt_sa_calls = k.sa_calls; r.cnt_lf_steps = k.lf_steps;
	r.cnt_ext_calls = k.ext_calls; r.cnt_ext_cells = k.ext_cells; r.cnt_sw_calls = k.sw_calls; r.cnt_sw_cells = k.sw_cells;
	r.cnt_glb_calls = k.glb_calls; r.cnt_glb_cells = k.glb_cells;
	if (out) *out = &c->res;
	return LC_OK;
}

int lc_batch_upload(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off)
{
	if (!c) return fail(LC_EARG, "null ctx");
	std::lock_guard<std::mutex> g(c->mu);
	return do_upload(c, n_pairs, bases, off, true);
}
int lc_batch_run(lc_ctx *c, int score_delta)
{
	if (!c) return fail(LC_EARG, "null ctx");
	std::lock_guard<std::mutex> g(c->mu);
	return do_run(c, score_delta);
}
int lc_batch_fetch(lc_ctx *c, lc_result_t **out)
{
	if (!c) return fail(LC_EARG, "null ctx");
	std::lock_guard<std::mutex> g(c->mu);
	return do_fetch(c, out);
}
void lc_free_result(lc_ctx *c, lc_result_t *res) { (void)c; (void)res; /* results live in the ctx and are recycled by the next call */ }

int lc_align_packed(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out)
{
	if (!c) return fail(LC_EARG, "null ctx");
	std::lock_guard<std::mutex> g(c->mu);
	return do_align(c, n_pairs, bases, off, score_delta, out);
}

int lc_align_pairs(lc_ctx *c, int n_pairs, const uint8_t *const *seq1, const int32_t *len1, const uint8_t *const *seq2, const int32_t *len2,
				   int score_delta, lc_result_t **out)
{
	if (!c || n_pairs < 0 || (n_pairs && (!seq1 || !seq2 || !len1 || !len2))) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	c->stage_off.assign(2 * (size_t)n_pairs + 1, 0);
	for (int p = 0; p < n_pairs; ++p) {
		if (len1[p] < 0 || len2[p] < 0) return fail(LC_EARG, "negative read length");
		c->stage_off[2 * p + 1] = c->stage_off[2 * p] + len1[p];
		c->stage_off[2 * p + 2] = c->stage_off[2 * p + 1] + len2[p];
	}
	c->stage.resize((size_t)c->stage_off[2 * (size_t)n_pairs] + 16);
	for (int p = 0; p < n_pairs; ++p) {
		if (len1[p]) memcpy(&c->stage[c->stage_off[2 * p]], seq1[p], len1[p]);
		if (len2[p]) memcpy(&c->stage[c->stage_off[2 * p + 1]], seq2[p], len2[p]);
	}
	return do_align(c, n_pairs, c->stage.data(), c->stage_off.data(), score_delta, out);
}

int64_t lc_fetch_seq(const lc_ctx *c, int rid, int64_t start, int64_t end, uint8_t *out)
{
	if (!c || rid < 0 || rid >= (int)c->hix.anns.size() || !out) return fail(LC_EARG, "bad arguments");
	// bns_fetch_seq clamps [beg,end) to the contig (B/bntseq.c:430-438); GetSeq passes contig-relative coordinates (G/gobwa.go:56-59)
	const LcAnn &a = c->hix.anns[rid];
	int64_t b = start + a.offset, e = end + a.offset;
	if (e < b) { int64_t t = b; b = e; e = t; }
	if (b < a.offset) b = a.offset;
	if (e > a.offset + a.len) e = a.offset + a.len;
	int64_t n = 0;
	for (int64_t p = b; p < e; ++p) out[n++] = (c->hix.pac[p >> 2] >> ((~p & 3) << 1)) & 3;
	return n;
}

// ---- single-read forms --------------------------------------------------------------------------------------
int lc_mem_align1_core(lc_ctx *c, int l_seq, const uint8_t *seq, lc_alnreg_v *out)
{
	if (!c || !out || l_seq < 0) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	// a pair whose second end is empty: no rescue happens in either direction (G/gobwa.go:287,311), so read 0's list
	// is exactly mem_align1_core's
	int64_t off[3] = { 0, l_seq, l_seq };
	std::vector<uint8_t> codes(seq, seq + l_seq);
	cudaSetDevice(c->device);
	int rc = c->pl.upload(1, codes.data(), off, false);
	if (rc) return fail(rc, c->pl.err);
	c->have_batch = true;
	rc = do_run(c, 25);
	lc_result_t *res = nullptr;
	if (!rc) rc = do_fetch(c, &res);
	if (rc) return rc;
	size_t n = (size_t)(res->reg_off[1] - res->reg_off[0]);
	out->n = out->m = n;
	out->a = (lc_alnreg_t *)malloc(sizeof(lc_alnreg_t) * (n ? n : 1));
	if (n) memcpy(out->a, res->regs + res->reg_off[0], sizeof(lc_alnreg_t) * n);
	return LC_OK;
}

int lc_mem_matesw(lc_ctx *c, const lc_pestat_t pes[4], const lc_alnreg_t *a, int l_ms, const uint8_t *ms, lc_alnreg_v *ma)
{
	if (!c || !pes || !a || !ms || !ma || l_ms <= 0 || l_ms > LC_MAX_READ_LEN) return fail(LC_EARG, "bad arguments");
	if (!(pes[0].failed && !pes[1].failed && pes[2].failed && pes[3].failed))
		return fail(LC_EARG, "only Lariat's insert model (FR enabled, FF/RF/RR failed: G/gobwa.go:229-237) is supported");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	// stage the mate as a one-pair batch so that the per-worker scratch is sized for it
	int64_t off[3] = { 0, l_ms, l_ms };
	int rc = c->pl.upload(1, ms, off, false);
	if (!rc) rc = c->pl.prepare_scratch();
	if (rc) return fail(rc, c->pl.err);
	LcBatch b = c->pl.b;
	b.pes_low = pes[1].low; b.pes_high = pes[1].high;
	b.w_u64_n = (b.pes_high - b.pes_low + b.lmax) / 2 + 8;
	void *d_ma = nullptr, *d_small = nullptr, *d_bl = nullptr;
	size_t n0 = ma->n;
	cudaMalloc(&d_ma, sizeof(lc_alnreg_t) * (n0 + 2)); cudaMalloc(&d_small, 64); cudaMalloc(&d_bl, 8 * (size_t)b.w_u64_n);
	b.w_u64 = (unsigned long long *)d_bl;
	if (n0) cudaMemcpy(d_ma, ma->a, sizeof(lc_alnreg_t) * n0, cudaMemcpyHostToDevice);
	int hn[2] = { (int)n0, 0 };
	cudaMemcpy(d_small, hn, 8, cudaMemcpyHostToDevice);
	matesw_single_kernel<<<1, 32, 0, c->be.stream>>>(b, *a, l_ms, b.seq, (lc_alnreg_t *)d_ma, (int *)d_small, (int *)d_small + 1);
	cudaStreamSynchronize(c->be.stream);
	cudaMemcpy(hn, d_small, 8, cudaMemcpyDeviceToHost);
	std::string e;
	if (c->be.check(e)) { cudaFree(d_ma); cudaFree(d_small); cudaFree(d_bl); return fail(LC_ECUDA, e); }
	ma->a = (lc_alnreg_t *)realloc(ma->a, sizeof(lc_alnreg_t) * (size_t)(hn[0] + 1));   // mem_matesw reallocs ma->a in place (B/bwamem_pair.c:165)
	ma->n = ma->m = (size_t)hn[0];
	if (hn[0]) cudaMemcpy(ma->a, d_ma, sizeof(lc_alnreg_t) * (size_t)hn[0], cudaMemcpyDeviceToHost);
	cudaFree(d_ma); cudaFree(d_small); cudaFree(d_bl);
	return hn[1];
}

int lc_mem_reg2aln(lc_ctx *c, int l_query, const uint8_t *query, const lc_alnreg_t *ar, lc_aln_t *aln, uint32_t **cigar)
{
	if (!c || !query || !ar || !aln || l_query <= 0) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	int64_t off[3] = { 0, l_query, l_query };
	int rc = c->pl.upload(1, query, off, false);
	if (!rc) rc = c->pl.run_single_cigar(*ar, c->hres);
	if (rc) return fail(rc, c->pl.err);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	*aln = c->hres.aln[0];
	if (cigar) {
		*cigar = (uint32_t *)malloc(4 * (size_t)(aln->n_cigar + 1));
		for (int i = 0; i < aln->n_cigar; ++i) (*cigar)[i] = c->hres.cigars[aln->cigar_off + i];
		aln->cigar_off = 0;
	}
	return LC_OK;
}

// ---- microbenchmark / kernel-level entry points ---------------------------------------------------------------
int lc_ksw_extend_batch(lc_ctx *c, int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
						const int32_t *h0, int w, int end_bonus, int zdrop, int32_t *out, float *ms, uint64_t *cells)
{
	if (!c || n < 0 || !q || !qoff || !t || !toff || !h0 || !out) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	int qmax = 1;
	for (int i = 0; i < n; ++i) { qmax = std::max<int>(qmax, (int)(qoff[i + 1] - qoff[i])); if (h0[i] <= 0) return fail(LC_EARG, "h0 must be positive (B/ksw.c:385)"); }
	DevTmp d;
	LcExtJobs j;
	j.n = n; j.q = d.up(q, (size_t)qoff[n] + 1); j.qoff = d.up(qoff, (size_t)n + 1); j.t = d.up(t, (size_t)toff[n] + 1); j.toff = d.up(toff, (size_t)n + 1);
	j.h0 = d.up(h0, n); j.w = w; j.end_bonus = end_bonus; j.zdrop = zdrop; j.sc = score_of(c->opt);
	j.out = d.alloc<int32_t>(6 * (size_t)n);
	j.qmax = qmax;
	j.ctr = d.alloc<unsigned int>(1);
	cudaMemset(j.ctr, 0, 4);
	LcCounters *cnt = d.alloc<LcCounters>(1);
	cudaMemset(cnt, 0, sizeof(LcCounters));
	j.cnt = cells ? cnt : nullptr;
	c->be.sync();
	c->be.ev_record(12);
	c->be.k_ksw_extend(j);
	c->be.ev_record(13);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	cudaMemcpy(out, j.out, sizeof(int32_t) * 6 * (size_t)n, cudaMemcpyDeviceToHost);
	if (ms) *ms = c->be.ev_ms(12, 13);
	if (cells) { LcCounters h; cudaMemcpy(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost); *cells = h.ext_cells; }
	return LC_OK;
}

int lc_ksw_align_batch(lc_ctx *c, int n, const uint8_t *q, const int64_t *qoff, const uint8_t *t, const int64_t *toff,
					   int32_t *out, float *ms, uint64_t *cells)
{
	if (!c || n < 0 || !q || !qoff || !t || !toff || !out) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	int qmax = 1, tmax = 1;
	for (int i = 0; i < n; ++i) { qmax = std::max<int>(qmax, (int)(qoff[i + 1] - qoff[i])); tmax = std::max<int>(tmax, (int)(toff[i + 1] - toff[i])); }
	DevTmp d;
	LcSwJobs j;
	j.n = n; j.q = d.up(q, (size_t)qoff[n] + 1); j.qoff = d.up(qoff, (size_t)n + 1); j.t = d.up(t, (size_t)toff[n] + 1); j.toff = d.up(toff, (size_t)n + 1);
	j.sc = score_of(c->opt); j.min_seed_len = c->opt.min_seed_len;
	j.out = d.alloc<int32_t>(7 * (size_t)n);
	j.qmax = qmax;
	j.ctr = d.alloc<unsigned int>(1);
	cudaMemset(j.ctr, 0, 4);
	j.w_u64_n = tmax / 2 + 8;
	j.w_u64 = d.alloc<unsigned long long>((size_t)c->be.n_workers() * j.w_u64_n);
	LcCounters *cnt = d.alloc<LcCounters>(1);
	cudaMemset(cnt, 0, sizeof(LcCounters));
	j.cnt = cells ? cnt : nullptr;
	c->be.sync();
	c->be.ev_record(12);
	c->be.k_ksw_align(j);
	c->be.ev_record(13);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	cudaMemcpy(out, j.out, sizeof(int32_t) * 7 * (size_t)n, cudaMemcpyDeviceToHost);
	if (ms) *ms = c->be.ev_ms(12, 13);
	if (cells) { LcCounters h; cudaMemcpy(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost); *cells = h.sw_cells; }
	return LC_OK;
}

int64_t lc_collect_intv_batch(lc_ctx *c, int n, const uint8_t *codes, const int64_t *off, int64_t *out_off, uint64_t *out, int64_t out_cap, float *ms)
{
	if (!c || n < 0 || !codes || !off || !out_off || !out) return fail(LC_EARG, "bad arguments");
	if (n & 1) return fail(LC_EARG, "n must be even (reads are staged as pairs)");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	int rc = c->pl.upload(n / 2, codes, off, false);
	if (!rc) rc = c->pl.run_seed_only();
	if (rc) return fail(rc, c->pl.err);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	// D2H of the per-read offsets/counts and the interval pool; x1 is not kept past seeding and is reported as 0
	std::vector<int64_t> ioff(n); std::vector<int32_t> icnt(n);
	cudaMemcpy(ioff.data(), c->pl.b.r_intv_off, 8 * (size_t)n, cudaMemcpyDeviceToHost);
	cudaMemcpy(icnt.data(), c->pl.b.r_n_intv, 4 * (size_t)n, cudaMemcpyDeviceToHost);
	int64_t total = 0;
	for (int i = 0; i < n; ++i) { out_off[i] = total; total += icnt[i]; }
	out_off[n] = total;
	if (total > out_cap) return fail(LC_EARG, "out_cap too small");
	std::vector<LcIntvOut> pool((size_t)c->pl.seed_intv_used() + 1);
	if (c->pl.seed_intv_used()) cudaMemcpy(pool.data(), c->pl.b.intv, sizeof(LcIntvOut) * c->pl.seed_intv_used(), cudaMemcpyDeviceToHost);
	for (int i = 0; i < n; ++i)
		for (int k = 0; k < icnt[i]; ++k) {
			const LcIntvOut &v = pool[(size_t)ioff[i] + k];
			uint64_t *o = out + 4 * (size_t)(out_off[i] + k);
			o[0] = v.x0; o[1] = 0; o[2] = v.x2; o[3] = (uint64_t)v.qbeg << 32 | v.qend;
		}
	if (ms) *ms = c->be.ev_ms(2, 3);
	return total;
}

int lc_sa_batch(lc_ctx *c, int64_t n, const uint64_t *rows, uint64_t *out, float *ms)
{
	if (!c || n < 0 || !rows || !out) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	cudaSetDevice(c->device);
	for (int64_t i = 0; i < n; ++i) if (rows[i] > c->ix.seq_len) return fail(LC_EARG, "BWT row out of range");
	DevTmp d;
	uint64_t *dr = d.up(rows, (size_t)n), *dout = d.alloc<uint64_t>((size_t)n);
	c->be.sync();
	c->be.ev_record(12);
	c->be.k_sa_batch(c->ix, dr, dout, n, nullptr);
	c->be.ev_record(13);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	cudaMemcpy(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost);
	if (ms) *ms = c->be.ev_ms(12, 13);
	return LC_OK;
}

} // extern "C"
