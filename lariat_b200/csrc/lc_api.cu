// lc_api.cu — the extern "C" boundary of libcuda_lariat.so (declared in include/cuda_lariat.h).
// Index load + upload to HBM, the batched alignment call, the split upload/run/fetch form used for timing,
// single-read source-compatible forms, and the kernel-level entry points of the microbenchmark config.
#include <cuda_runtime.h>
#include <chrono>
#include <condition_variable>
#include <deque>
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

// A stream + pool set of its own: large calls are cut into chunks of barcode groups that alternate between two lanes, so that one
// chunk's H2D / D2H and kernel tails overlap the other's kernels; the coalescer (below) runs merged small calls on lanes, too.
struct LcLane {
	LcCudaBackend be;
	LcPipeline<LcCudaBackend> pl;
	LcPinned in_bases, in_off, out_off, out_regs, out_aln, out_cig, out_alnx, out_mmr, out_mmd;      // pinned staging of a coalesced batch
	explicit LcLane(int dev) : be(dev), pl(be) {}
	~LcLane()
	{
		pl.release(); in_bases.release(); in_off.release(); out_off.release(); out_regs.release(); out_aln.release(); out_cig.release();
		out_alnx.release(); out_mmr.release(); out_mmd.release();
	}
};

// The result of one call.  It belongs to the CALLER from the moment the call returns until lc_free_result: nothing the library does
// afterwards (on this ctx, a clone, or another thread) touches it — what Lariat needs, whose ChainedHit pointers alias the result arrays
// until the barcode's arena.Free() while other goroutines keep aligning (I/lariat.go:348-350,473,546; G/gobwa.go:107-126).
// Large calls land directly in pinned arrays (D2H at PCIe speed, recycled through a free list); the per-call results of merged small
// calls are ordinary heap arrays cut out of the merged batch's staging.
#define LC_RESULT_MAGIC 0x4c435253u
struct LcResultBlock {
	lc_result_t res;                 // first member: the lc_result_t* handed to the caller IS the block
	uint32_t magic = LC_RESULT_MAGIC;
	bool pinned = false;
	LcPinned p_off, p_regs, p_aln, p_cig, p_alnx, p_mmr, p_mmd;
	std::vector<int64_t> v_off; std::vector<lc_alnreg_t> v_regs; std::vector<lc_aln_t> v_aln; std::vector<uint32_t> v_cig;
	std::vector<lc_alnx_t> v_alnx; std::vector<int32_t> v_mmr, v_mmd;
	LcResultBlock() { memset(&res, 0, sizeof(res)); }
	~LcResultBlock() { p_off.release(); p_regs.release(); p_aln.release(); p_cig.release(); p_alnx.release(); p_mmr.release(); p_mmd.release(); }
};

struct LcRequest {                   // one small call waiting in the coalescer
	const uint8_t *bases; const int64_t *off; int n_pairs, score_delta;
	enum { QUEUED, TAKEN, DONE } state = QUEUED;
	int rc = LC_OK; std::string err;
	LcResultBlock *blk = nullptr;
};

// Everything the contexts of one lc_init share: the index in HBM (one copy per device, however many host threads / cloned contexts
// use it), the contig table, the free list of pinned result blocks and the coalescer of small calls.
struct LcSharedIndex {
	int device = 0;
	LcHostIndex hix;
	void *d_bwt = nullptr, *d_sa = nullptr, *d_pac = nullptr, *d_anns = nullptr, *d_sa_dense = nullptr;
	LcIndex ix;
	lc_opt_t opt;
	lc_contigs_t contigs;
	std::vector<const char *> name_ptrs;
	std::vector<int32_t> lens;
	std::vector<int64_t> offsets;
	// tuning copied into every pipeline (lc_init reads the environment once)
	int lst_cap = 0, pk_cap = 64, lst_cap2 = 64, pk_cap2 = 256, lst_cap3 = 0, pk_cap3 = 2048, ext_heavy_seeds = 32, ext_heavy_chains = 8, chain_heavy = 48, small_reads = 16384;
	bool force_wide = false;
	bool want_alnx = false;       // lc_set_alnx: GetAlignments post-processing in every pipeline of the index
	// pinned result blocks given back by lc_free_result
	std::mutex blk_mu;
	std::vector<LcResultBlock *> free_blocks;
	// coalescer: calls of at most co_call_max pairs from any thread / context are merged into batches of at most co_batch_max pairs;
	// up to co_lanes_max batches are in flight (each on a lane of its own), led by whichever waiting caller finds a free lane
	std::mutex co_mu;
	std::condition_variable co_cv;
	std::deque<LcRequest *> co_q;
	std::vector<std::unique_ptr<LcLane> > co_lanes;
	std::vector<LcLane *> co_free;
	int co_call_max = 4096, co_batch_max = 65536, co_lanes_max = 3, co_linger_us = 0;
	uint64_t co_batches = 0, co_requests = 0;
	~LcSharedIndex()
	{
		cudaSetDevice(device);
		co_lanes.clear();
		for (auto *b : free_blocks) delete b;
		cudaFree(d_bwt); cudaFree(d_sa); cudaFree(d_pac); cudaFree(d_anns); cudaFree(d_sa_dense);
	}
	void tune(LcPipeline<LcCudaBackend> &pl) const
	{
		pl.lst_cap = lst_cap; pl.pk_cap = pk_cap; pl.lst_cap2 = lst_cap2; pl.pk_cap2 = pk_cap2; pl.lst_cap3 = lst_cap3; pl.pk_cap3 = pk_cap3;
		pl.ext_heavy_seeds = ext_heavy_seeds; pl.ext_heavy_chains = ext_heavy_chains; pl.force_wide = force_wide; pl.want_alnx = want_alnx;
		pl.chain_heavy = chain_heavy; pl.small_reads = small_reads;
	}
};

struct lc_ctx {
	int device = 0;
	std::shared_ptr<LcSharedIndex> sh;
	LcCudaBackend be;
	LcPipeline<LcCudaBackend> pl;
	LcHostResult hres;
	lc_result_t res;                // result of the split upload/run/fetch form and the single-read forms (valid until the next such call)
	std::mutex mu;
	std::vector<uint8_t> stage;
	std::vector<int64_t> stage_off;
	bool have_batch = false;
	std::vector<std::unique_ptr<LcLane> > lanes;      // lanes 2..n of chunked calls (lane 1 is be/pl above): own stream, pools and host thread
	int max_lanes = 2;              // LC_LANES
	int chunk_pairs = 500000;       // LC_CHUNK_PAIRS; calls of more than 1.5 chunks run chunked over two lanes, 0 = never
	int chunk_first_pct = 50;       // size of a call's first chunk, in % of an equal share (do_align_chunks; swept on C2: 100 / 60 / 50 / 40 / 30 -> 79.8 / 77.8 / 77.3 / 77.3 / 78.3 ms)
	explicit lc_ctx(int dev) : device(dev), be(dev), pl(be) { memset(&res, 0, sizeof(res)); }
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

static int env_int(const char *name, int lo, int hi, int dflt)
{
	const char *e = getenv(name);
	if (!e) return dflt;
	const int v = atoi(e);
	return v >= lo && v <= hi ? v : dflt;
}

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
		// substitution scores are looked up as bytes (PRMT tables of the extension and CIGAR kernels)
		if ((long long)o.a + o.b > 127) { fail(LC_EARG, "scoring options out of the supported range: a + b must be <= 127 (byte score tables)"); return nullptr; }
	}
	std::shared_ptr<LcSharedIndex> sh(new LcSharedIndex);
	sh->device = device;
	std::string err;
	if (!lc_load_index(idx_prefix, sh->hix, err)) { fail(LC_EINDEX, err); return nullptr; }
	cudaSetDevice(device);
	lc_ctx *c = new lc_ctx(device);
	c->sh = sh;
	LcHostIndex &h = sh->hix;
	LcIndex &ix = sh->ix;
	if (!lc_index_scalars(h, ix, err)) { fail(LC_EINDEX, err); delete c; return nullptr; }
	// upload the file payloads, then re-lay the BWT on the device as 32-byte blocks (lc_fmindex.cuh) and drop the file layout
	const uint64_t n_blocks = lc_occ_blocks(h.seq_len);
	size_t nb = h.bwt.size() * 4, ns = h.sa.size() * 8, np = h.pac.size(), na = h.anns.size() * sizeof(LcAnn);
	void *d_file = nullptr;
	if (cudaMalloc(&d_file, nb) != cudaSuccess || cudaMalloc(&sh->d_bwt, 32 * (size_t)n_blocks) != cudaSuccess || cudaMalloc(&sh->d_sa, ns) != cudaSuccess ||
		cudaMalloc(&sh->d_pac, np) != cudaSuccess || cudaMalloc(&sh->d_anns, na ? na : 1) != cudaSuccess) {
		cudaGetLastError(); if (d_file) cudaFree(d_file); fail(LC_ENOMEM, "cannot allocate the index in HBM"); delete c; return nullptr;
	}
	cudaMemcpy(d_file, h.bwt.data(), nb, cudaMemcpyHostToDevice);
	cudaMemcpy(sh->d_sa, h.sa.data(), ns, cudaMemcpyHostToDevice);
	cudaMemcpy(sh->d_pac, h.pac.data(), np, cudaMemcpyHostToDevice);
	cudaMemcpy(sh->d_anns, h.anns.data(), na, cudaMemcpyHostToDevice);
	ix.occ = (const uint4 *)sh->d_bwt; ix.sa = (const uint64_t *)sh->d_sa;
	ix.pac = (const uint8_t *)sh->d_pac; ix.anns = (const LcAnn *)sh->d_anns;
	c->be.k_occ_convert((const uint32_t *)d_file, (uint4 *)sh->d_bwt, n_blocks, ix);
	{
		std::string e;
		if (c->be.check(e)) { cudaFree(d_file); fail(LC_ECUDA, e); delete c; return nullptr; }
	}
	cudaFree(d_file);
	sh->force_wide = env_int("LC_FORCE_WIDE", 0, 1, 0) != 0;      // tests: the 64-bit paths (interval arithmetic, sa64) on a small index
	// dense suffix array: one entry per BWT row, built from the sampled one with one LF step per row.  Optional: if HBM is
	// short (or LC_DENSE_SA=0) the kernels fall back to the reference's sampled array + LF walks; results are identical.
	{
		const bool want = env_int("LC_DENSE_SA", 0, 1, 1) != 0;
		const bool wide = h.seq_len >= 0xffffffffull || sh->force_wide;
		const size_t bytes = (size_t)(h.seq_len + 1) * (wide ? 8 : 4);
		size_t free_b = 0, total_b = 0;
		cudaMemGetInfo(&free_b, &total_b);
		if (want && bytes < free_b / 2 && cudaMalloc(&sh->d_sa_dense, bytes) == cudaSuccess) {
			c->be.k_sa_fill(ix, (unsigned long long)h.sa.size(), sh->d_sa_dense, wide);
			std::string e;
			if (c->be.check(e)) { fail(LC_ECUDA, e); delete c; return nullptr; }
			if (wide) ix.sa64 = (const uint64_t *)sh->d_sa_dense; else ix.sa32 = (const uint32_t *)sh->d_sa_dense;
		} else cudaGetLastError();
	}
	sh->opt = o;
	// tuning knobs of the seeding state machine (tests use tiny values to force the redo path)
	sh->lst_cap = env_int("LC_SEED_LST_CAP", 2, 2048, sh->lst_cap); sh->pk_cap = env_int("LC_SEED_PK_CAP", 1, 4096, sh->pk_cap);
	sh->lst_cap2 = env_int("LC_SEED_LST_CAP2", 2, 2048, sh->lst_cap2); sh->pk_cap2 = env_int("LC_SEED_PK_CAP2", 1, 4096, sh->pk_cap2);
	sh->lst_cap3 = env_int("LC_SEED_LST_CAP3", 0, 2048, sh->lst_cap3); sh->pk_cap3 = env_int("LC_SEED_PK_CAP3", 1, 4096, sh->pk_cap3);
	sh->ext_heavy_seeds = env_int("LC_EXT_HEAVY_SEEDS", 0, 1 << 30, sh->ext_heavy_seeds);
	sh->ext_heavy_chains = env_int("LC_EXT_HEAVY_CHAINS", 0, 1 << 30, sh->ext_heavy_chains);
	sh->small_reads = env_int("LC_SMALL_READS", 0, 1 << 30, sh->small_reads);      // batches up to this many reads: warp-per-read seeding and extension (0 = never)
	sh->chain_heavy = env_int("LC_CHAIN_HEAVY", 0, 1 << 30, sh->chain_heavy);
	sh->co_call_max = env_int("LC_COALESCE_MAX", 0, 1 << 20, sh->co_call_max);          // 0: small calls are not merged
	sh->co_batch_max = env_int("LC_COALESCE_BATCH", 1, 1 << 22, sh->co_batch_max);
	sh->co_lanes_max = env_int("LC_COALESCE_LANES", 1, 16, sh->co_lanes_max);
	sh->co_linger_us = env_int("LC_COALESCE_LINGER_US", 0, 100000, sh->co_linger_us);
	c->pl.set_index(ix, o);
	sh->tune(c->pl);
	c->chunk_pairs = env_int("LC_CHUNK_PAIRS", 0, 1 << 30, c->chunk_pairs);
	c->max_lanes = env_int("LC_LANES", 1, 8, c->max_lanes);
	// host copies of the big FM-index arrays are not needed any more (pac stays for lc_fetch_seq)
	std::vector<uint32_t>().swap(h.bwt);
	std::vector<uint64_t>().swap(h.sa);
	for (size_t i = 0; i < h.names.size(); ++i) { sh->name_ptrs.push_back(h.names[i].c_str()); sh->lens.push_back(h.anns[i].len); sh->offsets.push_back(h.anns[i].offset); }
	sh->contigs.n_seqs = (int32_t)h.names.size(); sh->contigs.l_pac = h.l_pac;
	sh->contigs.names = sh->name_ptrs.data(); sh->contigs.lens = sh->lens.data(); sh->contigs.offsets = sh->offsets.data();
	return c;
}

// A second context on the same index: its own stream, pools and lock, NO second copy of the index in HBM.  For callers that want
// several large calls in flight at once (one clone per worker); small calls do not need it — they are merged across threads anyway.
lc_ctx *lc_clone(lc_ctx *src)
{
	if (!src) { fail(LC_EARG, "null ctx"); return nullptr; }
	cudaSetDevice(src->device);
	lc_ctx *c = new lc_ctx(src->device);
	c->sh = src->sh;
	c->pl.set_index(c->sh->ix, c->sh->opt);
	c->sh->tune(c->pl);
	c->pl.counting = src->pl.counting;
	c->chunk_pairs = src->chunk_pairs; c->max_lanes = src->max_lanes;
	return c;
}

void lc_destroy(lc_ctx *c)      // the index leaves HBM when its last context is destroyed; results already handed out stay valid
{
	if (!c) return;
	cudaSetDevice(c->device);
	c->pl.release();
	c->lanes.clear();
	delete c;
}

const lc_contigs_t *lc_contigs(const lc_ctx *c) { return c ? &c->sh->contigs : nullptr; }
void lc_set_counting(lc_ctx *c, int on) { if (c) c->pl.counting = on != 0; }
void lc_set_alnx(lc_ctx *c, int on)
{
	if (!c) return;
	std::lock_guard<std::mutex> g(c->mu);
	std::lock_guard<std::mutex> g2(c->sh->co_mu);      // no merged batch is being assembled while the switch flips
	c->sh->want_alnx = on != 0;
	c->pl.want_alnx = on != 0;
}
int lc_set_chunk_pairs(lc_ctx *c, int n_pairs)
{
	if (!c || n_pairs < 0) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->mu);
	c->chunk_pairs = n_pairs;
	return LC_OK;
}
int lc_set_coalescing(lc_ctx *c, int max_call_pairs, int max_batch_pairs, int lanes, int linger_us)
{
	if (!c || max_call_pairs < 0 || max_batch_pairs < 1 || lanes < 1 || lanes > 16 || linger_us < 0) return fail(LC_EARG, "bad arguments");
	std::lock_guard<std::mutex> g(c->sh->co_mu);
	c->sh->co_call_max = max_call_pairs; c->sh->co_batch_max = max_batch_pairs; c->sh->co_lanes_max = lanes; c->sh->co_linger_us = linger_us;
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
	r.alnx = h.alnx.empty() ? nullptr : h.alnx.data(); r.mm_ref_pos = h.mm_ref.data(); r.mm_read_pos = h.mm_read.data(); r.n_mm = (int64_t)h.mm_ref.size();
	if (!r.alnx) { r.mm_ref_pos = r.mm_read_pos = nullptr; r.n_mm = 0; }
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

// ---- result blocks ---------------------------------------------------------------------------------------------------------
static LcResultBlock *take_pinned_block(LcSharedIndex &S)
{
	{
		std::lock_guard<std::mutex> g(S.blk_mu);
		if (!S.free_blocks.empty()) { LcResultBlock *b = S.free_blocks.back(); S.free_blocks.pop_back(); return b; }
	}
	LcResultBlock *b = new LcResultBlock;
	b->pinned = true;
	return b;
}
static void give_block(LcSharedIndex &S, LcResultBlock *b)
{
	if (!b) return;
	if (!b->pinned) { delete b; return; }
	memset(&b->res, 0, sizeof(b->res));
	std::lock_guard<std::mutex> g(S.blk_mu);
	if (S.free_blocks.size() < 4) S.free_blocks.push_back(b); else { cudaSetDevice(S.device); delete b; }
}
static void set_counters(lc_result_t &r, const LcCounters &k)
{
	r.cnt_occ_blocks = k.occ_blocks; r.cnt_extend_calls = k.extend_calls; r.cnt_sa_calls = k.sa_calls; r.cnt_lf_steps = k.lf_steps;
	r.cnt_ext_calls = k.ext_calls; r.cnt_ext_cells = k.ext_cells; r.cnt_sw_calls = k.sw_calls; r.cnt_sw_cells = k.sw_cells;
	r.cnt_glb_calls = k.glb_calls; r.cnt_glb_cells = k.glb_cells;
}
static void set_stage_ms(lc_result_t &r, const float ms[10])
{
	r.ms_h2d = ms[0]; r.ms_seed = ms[1]; r.ms_sa = ms[2]; r.ms_chain = ms[3]; r.ms_extend = ms[4]; r.ms_dedup = ms[5];
	r.ms_pair = ms[6]; r.ms_cigar = ms[7]; r.ms_d2h = ms[8]; r.ms_total = ms[9];
}
// every read 0 <= len <= LC_MAX_READ_LEN, in int64 (the caller's offsets are not trusted)
static bool offsets_ok(int n_pairs, const int64_t *off, int *lmax = nullptr)
{
	int64_t lo, hi;
	LcPipeline<LcCudaBackend>::length_range(off, 2 * (size_t)n_pairs, lo, hi);
	if (lmax) *lmax = (int)std::max<int64_t>(1, std::min<int64_t>(hi, LC_MAX_READ_LEN));
	return lo >= 0 && hi <= LC_MAX_READ_LEN;
}

// ---- the batched call: upload -> stages -> pack -> D2H into the pinned result arrays; large calls chunked over two lanes -----
struct LcChunkShared {
	std::mutex mu;
	std::condition_variable cv;
	std::vector<int64_t> cum_reg;       // [n_chunks + 1] candidates before chunk k, valid once ready[k]
	std::vector<int64_t> cum_cig, cum_mm;
	std::vector<char> ready_reg, ready_cig;
	int up_next = 0;                    // the chunk whose upload turn it is
	std::vector<int> bounds;            // [n_chunks + 1] chunk k = pairs bounds[k] .. bounds[k + 1]
	bool late_regs = false;             // LC_LATE_REGS=1 (tests): every chunk takes the "CIGAR stage first, regions' copy after" order
	std::shared_mutex res_mu;           // shared: a D2H into the pinned arrays is in flight; exclusive: the arrays are being regrown
	int rc = LC_OK;
	std::string err;
	float ms[10] = {0};
	int64_t launches = 0;
	LcCounters cnt;
	LcResultBlock *blk = nullptr;
	void fail_with(int code, const std::string &e) { std::lock_guard<std::mutex> g(mu); if (rc == LC_OK) { rc = code; err = e; } }
};

static void add_counters(LcCounters &a, const LcCounters &b)
{
	a.occ_blocks += b.occ_blocks; a.extend_calls += b.extend_calls; a.sa_calls += b.sa_calls; a.lf_steps += b.lf_steps;
	a.ext_calls += b.ext_calls; a.ext_cells += b.ext_cells; a.sw_calls += b.sw_calls; a.sw_cells += b.sw_cells;
	a.glb_calls += b.glb_calls; a.glb_cells += b.glb_cells;
}

// chunks first, first + stride, ... of the call on one lane (its own stream, pools and host thread).
// Copies are kept off the critical path in three ways: (1) the chunks' uploads take turns in chunk order — two lanes uploading at once
// would share the link and both start computing late; (2) the regions are final after the pair stage, so they are packed and copied out
// on the lane's copy stream while its CIGAR kernels run; (3) one lane's remaining copies sit under the other lane's kernels.
static void lane_worker(lc_ctx *c, LcCudaBackend *be, LcPipeline<LcCudaBackend> *pl, LcChunkShared *sh, int first, int stride, int n_chunks, int n_pairs,
						const uint8_t *bases, const int64_t *off, int score_delta)
{
	cudaSetDevice(c->device);
	LcResultBlock &B = *sh->blk;
	for (int k = first; k < n_chunks; k += stride) {
		const int p0 = sh->bounds[k], p1 = sh->bounds[k + 1];
		int rc = LC_OK, lmax = 0;
		// the chunk's offsets are checked here, while an earlier chunk uploads
		const bool off_ok = offsets_ok(p1 - p0, off + 2 * (size_t)p0, &lmax);
		{   // the upload turn of chunk k
			std::unique_lock<std::mutex> g(sh->mu);
			sh->cv.wait(g, [&] { return sh->up_next == k; });
			rc = sh->rc;
		}
		if (!rc && !off_ok) { rc = LC_EARG; pl->err = "read offsets must be non-decreasing and no read longer than LC_MAX_READ_LEN"; }
		if (!rc) { rc = pl->upload(p1 - p0, bases, off + 2 * (size_t)p0, true, lmax); be->sync(); }
		{ std::lock_guard<std::mutex> g(sh->mu); sh->up_next = k + 1; }
		sh->cv.notify_all();
		if (!rc) rc = pl->run_front(score_delta);
		if (rc) sh->fail_with(rc, pl->err);
		int64_t reg_base;
		bool cigar_done = false;
		for (;;) {   // chunk k's place in the call's region arrays = candidates of the chunks before it (known when their pair stage is done)
			{
				std::unique_lock<std::mutex> g(sh->mu);
				if ((!sh->ready_reg[k] || sh->late_regs) && !cigar_done && !rc) { g.unlock(); rc = pl->run_cigar(); cigar_done = true; if (rc) sh->fail_with(rc, pl->err); continue; }   // (the chunk before is not that far yet: the CIGAR stage first, the regions' copy after it)
				sh->cv.wait(g, [&] { return sh->ready_reg[k] != 0; });
				reg_base = sh->cum_reg[k];
				sh->cum_reg[k + 1] = reg_base + (rc ? 0 : (int64_t)pl->n_cand);
				sh->ready_reg[k + 1] = 1;
			}
			break;
		}
		sh->cv.notify_all();
		const bool ax = pl->want_alnx;
		const size_t need_reg = (size_t)(reg_base + (rc ? 0 : (int64_t)pl->n_cand));
		bool regs_in_flight = false;
		std::shared_lock<std::shared_mutex> rg(sh->res_mu, std::defer_lock);      // held while a D2H into the pinned arrays is in flight
		if (!rc) rc = pl->run_pack_regs(reg_base);
		if (!rc) {
			if (need_reg * sizeof(lc_alnreg_t) > B.p_regs.cap || need_reg * sizeof(lc_aln_t) > B.p_aln.cap || (ax && need_reg * sizeof(lc_alnx_t) > B.p_alnx.cap)) {
				// regrow (rare: blocks are recycled with their capacities): extrapolate from the chunks seen so far, wait out in-flight copies, keep what is there
				std::unique_lock<std::shared_mutex> g(sh->res_mu);
				const size_t want_reg = (size_t)(need_reg * (1.2 * n_pairs / (double)std::max(1, sh->bounds[k + 1]))) + 1024;
				if (!B.p_regs.ensure(want_reg * sizeof(lc_alnreg_t), B.p_regs.cap) || !B.p_aln.ensure(want_reg * sizeof(lc_aln_t), B.p_aln.cap) ||
					(ax && !B.p_alnx.ensure(want_reg * sizeof(lc_alnx_t), B.p_alnx.cap))) rc = LC_ENOMEM;
			}
			if (rc) sh->fail_with(rc, "cannot allocate pinned host memory for the results");
			else {
				rg.lock();
				const size_t n_off = 2 * (size_t)(p1 - p0) + (k == n_chunks - 1 ? 1 : 0);
				pl->fetch_regs_into_side((int64_t *)B.p_off.p + 2 * (size_t)p0, n_off, (lc_alnreg_t *)B.p_regs.p + reg_base);
				regs_in_flight = true;
			}
		} else sh->fail_with(rc, pl->err);
		if (!rc && !cigar_done) { rc = pl->run_cigar(); if (rc) sh->fail_with(rc, pl->err); }
		if (regs_in_flight) { be->copy_sync(); rg.unlock(); }
		int64_t cig_base, mm_base;
		{   // ... and in the CIGAR / mismatch-position arrays (known when their CIGAR stage is done)
			std::unique_lock<std::mutex> g(sh->mu);
			sh->cv.wait(g, [&] { return sh->ready_cig[k] != 0; });
			cig_base = sh->cum_cig[k]; mm_base = sh->cum_mm[k];
			sh->cum_mm[k + 1] = mm_base + (rc ? 0 : (int64_t)pl->n_mm);
			sh->cum_cig[k + 1] = cig_base + (rc ? 0 : (int64_t)pl->n_cigar_words);
			sh->ready_cig[k + 1] = 1;
			if (!rc && sh->cum_cig[k + 1] > 0xffffffffll) { rc = LC_ENOMEM; if (sh->rc == LC_OK) { sh->rc = rc; sh->err = "more than 2^32 CIGAR words in one call"; } }
			if (!rc && sh->cum_mm[k + 1] > 0xffffffffll) { rc = LC_ENOMEM; if (sh->rc == LC_OK) { sh->rc = rc; sh->err = "more than 2^32 mismatch positions in one call"; } }
		}
		sh->cv.notify_all();
		if (rc) continue;
		rc = pl->run_pack_aln((uint32_t)cig_base, (uint32_t)mm_base);
		if (rc) { sh->fail_with(rc, pl->err); continue; }
		const size_t need_cig = (size_t)(cig_base + (int64_t)pl->n_cigar_words);
		const size_t need_mm = ax ? (size_t)(mm_base + (int64_t)pl->n_mm) : 0;
		if (need_cig * 4 > B.p_cig.cap || (ax && (need_mm * 4 > B.p_mmr.cap || need_mm * 4 > B.p_mmd.cap))) {
			std::unique_lock<std::shared_mutex> g(sh->res_mu);
			const double f = 1.2 * n_pairs / (double)std::max(1, sh->bounds[k + 1]);
			const size_t want_cig = (size_t)(need_cig * f) + 1024, want_mm = (size_t)(need_mm * f) + 1024;
			if (!B.p_cig.ensure(want_cig * 4, B.p_cig.cap) || (ax && (!B.p_mmr.ensure(want_mm * 4, B.p_mmr.cap) || !B.p_mmd.ensure(want_mm * 4, B.p_mmd.cap)))) {
				sh->fail_with(LC_ENOMEM, "cannot allocate pinned host memory for the results"); continue;
			}
		}
		float ms[10]; LcCounters cn;
		{
			std::shared_lock<std::shared_mutex> g(sh->res_mu);
			pl->fetch_aln_into((lc_aln_t *)B.p_aln.p + reg_base, (uint32_t *)B.p_cig.p + cig_base, ax ? (lc_alnx_t *)B.p_alnx.p + reg_base : nullptr,
							   ax ? (int32_t *)B.p_mmr.p + mm_base : nullptr, ax ? (int32_t *)B.p_mmd.p + mm_base : nullptr);
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
	LcSharedIndex &S = *c->sh;
	// The first chunk's upload has nothing to hide under, so the first chunk is made smaller than the others (first_pct % of an
	// equal share, LC_CHUNK_FIRST; the rest is split evenly).  (A small LAST chunk as well — LC_CHUNK_TAPER=1 — was measured and is
	// off: a small batch running alone at the end loses more to its kernels' tails than its copy-out costs.)
	std::vector<int> bounds;
	{
		const char *e = getenv("LC_CHUNK_TAPER");
		const char *f = getenv("LC_CHUNK_FIRST");
		const int first_pct = f ? std::max(5, std::min(100, atoi(f))) : c->chunk_first_pct;
		if (n_chunks >= 2 && n_pairs >= 2 * n_chunks && e && atoi(e) == 1) {
			const int U = n_chunks;
			bounds.push_back(0);
			for (int j = 1; j <= U; ++j) bounds.push_back((int)((int64_t)n_pairs * (2 * j - 1) / (2 * U)));
			bounds.push_back(n_pairs);
			n_chunks = U + 1;
		} else if (n_chunks >= 2 && n_pairs >= 100 * n_chunks && first_pct < 100) {
			const int first = (int)((int64_t)n_pairs * first_pct / (100LL * n_chunks));
			bounds.push_back(0);
			for (int k = 0; k < n_chunks; ++k) bounds.push_back(first + (int)((int64_t)(n_pairs - first) * k / (n_chunks - 1)));
		} else for (int k = 0; k <= n_chunks; ++k) bounds.push_back((int)((int64_t)n_pairs * k / n_chunks));
	}
	const int n_lanes = std::max(1, std::min(n_chunks, c->max_lanes));
	while ((int)c->lanes.size() < n_lanes - 1) {
		std::unique_ptr<LcLane> l(new LcLane(c->device));
		l->pl.set_index(S.ix, S.opt);
		S.tune(l->pl);
		c->lanes.push_back(std::move(l));
	}
	for (auto &l : c->lanes) { l->pl.counting = c->pl.counting; l->pl.want_alnx = c->pl.want_alnx; }
	LcResultBlock *blk = take_pinned_block(S);
	if (!blk->p_off.ensure(8 * (2 * (size_t)n_pairs + 1), 0)) { give_block(S, blk); return fail(LC_ENOMEM, "cannot allocate pinned host memory for the results"); }
	LcChunkShared sh;
	sh.blk = blk;
	sh.bounds = bounds;
	{ const char *e = getenv("LC_LATE_REGS"); sh.late_regs = e && atoi(e) == 1; }
	sh.cum_reg.assign(n_chunks + 1, 0); sh.cum_cig.assign(n_chunks + 1, 0); sh.cum_mm.assign(n_chunks + 1, 0); sh.ready_reg.assign(n_chunks + 1, 0); sh.ready_cig.assign(n_chunks + 1, 0); sh.ready_reg[0] = sh.ready_cig[0] = 1;
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
	if (sh.rc) { give_block(S, blk); return fail(sh.rc, sh.err); }
	lc_result_t &r = blk->res;
	memset(&r, 0, sizeof(r));
	r.n_reads = 2 * n_pairs; r.n_regs = sh.cum_reg[n_chunks]; r.n_cigar_words = sh.cum_cig[n_chunks];
	r.reg_off = (const int64_t *)blk->p_off.p; r.regs = (const lc_alnreg_t *)blk->p_regs.p; r.aln = (const lc_aln_t *)blk->p_aln.p; r.cigars = (const uint32_t *)blk->p_cig.p;
	if (r.reg_off[r.n_reads] != r.n_regs) { give_block(S, blk); return fail(LC_ECUDA, "result size mismatch"); }
	if (c->pl.want_alnx) { r.alnx = (const lc_alnx_t *)blk->p_alnx.p; r.mm_ref_pos = (const int32_t *)blk->p_mmr.p; r.mm_read_pos = (const int32_t *)blk->p_mmd.p; r.n_mm = sh.cum_mm[n_chunks]; }
	// stage times: summed over the chunks (chunks of the two lanes overlap); ms_total: the whole call on the device clock, copies included when chunked
	set_stage_ms(r, sh.ms);
	if (n_chunks > 1) r.ms_total = c->be.ev_ms(12, 13);
	r.gpu_launches = sh.launches;
	set_counters(r, sh.cnt);
	if (out) *out = &blk->res; else give_block(S, blk);
	return LC_OK;
}

// ---- small calls: merged across threads ----------------------------------------------------------------------------------------
// Lariat's loop hands BWA one barcode group (~100 pairs) at a time from each of its worker goroutines (I/lariat.go:348-376).  One such
// call cannot fill a GPU, and the pipeline's latency (a read's seeding alone is ~700 dependent FM-index steps) does not shrink with the
// batch: so calls of up to co_call_max pairs, from ANY thread and ANY context of the index, are queued, and whichever waiting caller
// finds a free lane takes everything queued (up to co_batch_max pairs), runs it as ONE batch and cuts the results apart again.
// Results are identical to separate calls (reads are independent; test_coalesced_calls_equal_separate_calls).
static void run_coalesced(LcSharedIndex &S, LcLane *lane, std::vector<LcRequest *> &batch, int score_delta)
{
	cudaSetDevice(S.device);
	int tot = 0; int64_t nb = 0;
	for (auto *r : batch) { tot += r->n_pairs; nb += r->off[2 * (size_t)r->n_pairs] - r->off[0]; }
	auto fail_all = [&](int rc, const std::string &e) { for (auto *r : batch) { r->rc = rc; r->err = e; } };
	if (!lane->in_bases.ensure((size_t)nb + 64, 0) || !lane->in_off.ensure(8 * (2 * (size_t)tot + 1), 0)) return fail_all(LC_ENOMEM, "cannot allocate pinned staging for a merged batch");
	uint8_t *hb = (uint8_t *)lane->in_bases.p; int64_t *ho = (int64_t *)lane->in_off.p;
	{
		int64_t o = 0; size_t rr = 0;
		for (auto *r : batch) {
			const int64_t o0 = r->off[0], n = r->off[2 * (size_t)r->n_pairs] - o0;
			if (n) memcpy(hb + o, r->bases + o0, (size_t)n);
			for (size_t k = 0; k < 2 * (size_t)r->n_pairs; ++k) ho[rr++] = o + (r->off[k] - o0);
			o += n;
		}
		ho[rr] = o;
	}
	LcPipeline<LcCudaBackend> &pl = lane->pl;
	pl.want_alnx = S.want_alnx;
	int rc = pl.upload(tot, hb, ho, true);
	if (!rc) rc = pl.run_stages(score_delta);
	if (!rc) rc = pl.run_pack(0, 0);
	if (rc) return fail_all(rc, pl.err);
	const size_t NC = (size_t)pl.n_cand, NW = (size_t)pl.n_cigar_words, NM = (size_t)pl.n_mm;
	const bool ax = pl.want_alnx;
	if (!lane->out_off.ensure(8 * (2 * (size_t)tot + 1), 0) || !lane->out_regs.ensure(sizeof(lc_alnreg_t) * (NC + 1), 0) ||
		!lane->out_aln.ensure(sizeof(lc_aln_t) * (NC + 1), 0) || !lane->out_cig.ensure(4 * (NW + 1), 0) ||
		(ax && (!lane->out_alnx.ensure(sizeof(lc_alnx_t) * (NC + 1), 0) || !lane->out_mmr.ensure(4 * (NM + 1), 0) || !lane->out_mmd.ensure(4 * (NM + 1), 0))))
		return fail_all(LC_ENOMEM, "cannot allocate pinned staging for a merged batch");
	pl.fetch_into((int64_t *)lane->out_off.p, 2 * (size_t)tot + 1, (lc_alnreg_t *)lane->out_regs.p, (lc_aln_t *)lane->out_aln.p, (uint32_t *)lane->out_cig.p,
				  ax ? (lc_alnx_t *)lane->out_alnx.p : nullptr, ax ? (int32_t *)lane->out_mmr.p : nullptr, ax ? (int32_t *)lane->out_mmd.p : nullptr);
	lane->be.sync();
	std::string e;
	if (lane->be.check(e)) return fail_all(LC_ECUDA, e);
	float ms[10]; LcCounters cn;
	pl.fetch_done(ms, cn);
	const int64_t *roff = (const int64_t *)lane->out_off.p;
	const lc_alnreg_t *regs = (const lc_alnreg_t *)lane->out_regs.p;
	const lc_aln_t *aln = (const lc_aln_t *)lane->out_aln.p;
	const uint32_t *cig = (const uint32_t *)lane->out_cig.p;
	const lc_alnx_t *alnx = (const lc_alnx_t *)lane->out_alnx.p;
	const int32_t *mmr = (const int32_t *)lane->out_mmr.p, *mmd = (const int32_t *)lane->out_mmd.p;
	if ((size_t)roff[2 * (size_t)tot] != NC) return fail_all(LC_ECUDA, "result size mismatch");
	size_t r0 = 0;
	for (auto *rq : batch) {      // cut the merged result apart: each request gets arrays of its own
		const size_t R = 2 * (size_t)rq->n_pairs;
		const int64_t c0 = roff[r0], c1 = roff[r0 + R];
		LcResultBlock *b = new LcResultBlock;
		b->v_off.resize(R + 1);
		for (size_t k = 0; k <= R; ++k) b->v_off[k] = roff[r0 + k] - c0;
		b->v_regs.assign(regs + c0, regs + c1);
		b->v_aln.assign(aln + c0, aln + c1);
		size_t nw = 0;
		for (auto &a : b->v_aln) nw += (size_t)a.n_cigar;
		b->v_cig.resize(nw + 1);
		nw = 0;
		for (auto &a : b->v_aln) {      // CIGARs are carved from the batch's pool in no particular order: gather them per candidate
			if (a.n_cigar) memcpy(&b->v_cig[nw], cig + a.cigar_off, 4 * (size_t)a.n_cigar);
			a.cigar_off = (uint32_t)nw; nw += (size_t)a.n_cigar;
		}
		lc_result_t &r = b->res;
		if (ax) {      // likewise the mismatch position lists
			b->v_alnx.assign(alnx + c0, alnx + c1);
			size_t nm = 0;
			for (auto &x : b->v_alnx) nm += (size_t)x.n_mm;
			b->v_mmr.resize(nm + 1); b->v_mmd.resize(nm + 1);
			nm = 0;
			for (auto &x : b->v_alnx) {
				if (x.n_mm) { memcpy(&b->v_mmr[nm], mmr + x.mm_off, 4 * (size_t)x.n_mm); memcpy(&b->v_mmd[nm], mmd + x.mm_off, 4 * (size_t)x.n_mm); }
				x.mm_off = (uint32_t)nm; nm += (size_t)x.n_mm;
			}
			r.alnx = b->v_alnx.data(); r.mm_ref_pos = b->v_mmr.data(); r.mm_read_pos = b->v_mmd.data(); r.n_mm = (int64_t)nm;
		}
		r.n_reads = (int32_t)R; r.n_regs = c1 - c0; r.n_cigar_words = (int64_t)nw;
		r.reg_off = b->v_off.data(); r.regs = b->v_regs.data(); r.aln = b->v_aln.data(); r.cigars = b->v_cig.data();
		set_stage_ms(r, ms);      // times and launches are those of the merged batch the call rode in
		r.gpu_launches = pl.launches;
		rq->blk = b;
		r0 += R;
	}
}

static int coalesced_align(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out)
{
	LcSharedIndex &S = *c->sh;
	LcRequest rq;
	rq.bases = bases; rq.off = off; rq.n_pairs = n_pairs; rq.score_delta = score_delta;
	std::unique_lock<std::mutex> g(S.co_mu);
	S.co_q.push_back(&rq);
	++S.co_requests;
	for (;;) {
		if (rq.state == LcRequest::DONE) break;
		const bool lane_free = !S.co_free.empty() || (int)S.co_lanes.size() < S.co_lanes_max;
		if (rq.state == LcRequest::QUEUED && lane_free) {
			// lead one batch: everything queued with the head's score_delta, up to co_batch_max pairs
			LcLane *lane = nullptr;
			if (!S.co_free.empty()) { lane = S.co_free.back(); S.co_free.pop_back(); }
			else {
				S.co_lanes.emplace_back(nullptr);      // reserve the slot under the lock, build the lane outside it
				const size_t slot = S.co_lanes.size() - 1;
				g.unlock();
				cudaSetDevice(S.device);
				std::unique_ptr<LcLane> l(new LcLane(S.device));
				l->pl.set_index(S.ix, S.opt);
				S.tune(l->pl);
				g.lock();
				lane = l.get();
				S.co_lanes[slot] = std::move(l);
			}
			if (S.co_linger_us > 0 && (int)S.co_q.size() < 2) {      // optional: give other callers a moment to join
				g.unlock(); std::this_thread::sleep_for(std::chrono::microseconds(S.co_linger_us)); g.lock();
			}
			std::vector<LcRequest *> batch;
			int tot = 0;
			if (!S.co_q.empty()) {
				const int delta = S.co_q.front()->score_delta;
				for (auto it = S.co_q.begin(); it != S.co_q.end();) {
					if ((*it)->score_delta == delta && (batch.empty() || tot + (*it)->n_pairs <= S.co_batch_max)) {
						batch.push_back(*it); tot += (*it)->n_pairs; (*it)->state = LcRequest::TAKEN; it = S.co_q.erase(it);
					} else ++it;
				}
				++S.co_batches;
				g.unlock();
				run_coalesced(S, lane, batch, delta);
				g.lock();
				for (auto *r : batch) r->state = LcRequest::DONE;
			}
			S.co_free.push_back(lane);
			S.co_cv.notify_all();
			continue;
		}
		S.co_cv.wait(g);
	}
	g.unlock();
	if (rq.rc) { if (rq.blk) delete rq.blk; return fail(rq.rc, rq.err); }
	if (out) *out = &rq.blk->res; else delete rq.blk;
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
void lc_free_result(lc_ctx *c, lc_result_t *res)
{
	if (!c || !res || res == &c->res) return;      // the split / single-read forms hand out the context's own block: nothing to free
	LcResultBlock *b = (LcResultBlock *)res;
	if (b->magic != LC_RESULT_MAGIC) return;
	give_block(*c->sh, b);
}

int lc_align_packed(lc_ctx *c, int n_pairs, const uint8_t *bases, const int64_t *off, int score_delta, lc_result_t **out)
{
	if (!c || n_pairs < 0 || !off || (n_pairs && !bases && off[2 * (size_t)n_pairs] > off[0])) return fail(LC_EARG, "bad arguments");
	// small calls are merged across threads — unless the work counters are on (they are per call) or the caller asked for chunks this small
	if (n_pairs > 0 && n_pairs <= c->sh->co_call_max && !c->pl.counting && !(c->chunk_pairs > 0 && (double)n_pairs > 1.5 * c->chunk_pairs)) {
		if (!offsets_ok(n_pairs, off)) return fail(LC_EARG, "read offsets must be non-decreasing and no read longer than LC_MAX_READ_LEN");
		return coalesced_align(c, n_pairs, bases, off, score_delta, out);
	}
	// (larger calls: every chunk's offsets are checked by its lane before anything of the chunk is read)
	std::lock_guard<std::mutex> g(c->mu);
	return do_align(c, n_pairs, bases, off, score_delta, out);
}

int lc_align_pairs(lc_ctx *c, int n_pairs, const uint8_t *const *seq1, const int32_t *len1, const uint8_t *const *seq2, const int32_t *len2,
				   int score_delta, lc_result_t **out)
{
	if (!c || n_pairs < 0 || (n_pairs && (!seq1 || !seq2 || !len1 || !len2))) return fail(LC_EARG, "bad arguments");
	// the packed copy is the calling thread's own (the call may be merged with other threads' calls)
	std::vector<int64_t> off(2 * (size_t)n_pairs + 1, 0);
	for (int p = 0; p < n_pairs; ++p) {
		if (len1[p] < 0 || len2[p] < 0) return fail(LC_EARG, "negative read length");
		off[2 * p + 1] = off[2 * p] + len1[p];
		off[2 * p + 2] = off[2 * p + 1] + len2[p];
	}
	std::vector<uint8_t> stage((size_t)off[2 * (size_t)n_pairs] + 16);
	for (int p = 0; p < n_pairs; ++p) {
		if (len1[p]) memcpy(&stage[off[2 * p]], seq1[p], len1[p]);
		if (len2[p]) memcpy(&stage[off[2 * p + 1]], seq2[p], len2[p]);
	}
	return lc_align_packed(c, n_pairs, stage.data(), off.data(), score_delta, out);
}

int64_t lc_fetch_seq(const lc_ctx *c, int rid, int64_t start, int64_t end, uint8_t *out)
{
	if (!c || rid < 0 || rid >= (int)c->sh->hix.anns.size() || !out) return fail(LC_EARG, "bad arguments");
	// bns_fetch_seq clamps [beg,end) to the contig (B/bntseq.c:430-438); GetSeq passes contig-relative coordinates (G/gobwa.go:56-59)
	const LcAnn &a = c->sh->hix.anns[rid];
	int64_t b = start + a.offset, e = end + a.offset;
	if (e < b) { int64_t t = b; b = e; e = t; }
	if (b < a.offset) b = a.offset;
	if (e > a.offset + a.len) e = a.offset + a.len;
	int64_t n = 0;
	for (int64_t p = b; p < e; ++p) out[n++] = (c->sh->hix.pac[p >> 2] >> ((~p & 3) << 1)) & 3;
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
	j.h0 = d.up(h0, n); j.w = w; j.end_bonus = end_bonus; j.zdrop = zdrop; j.sc = score_of(c->sh->opt);
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
	j.sc = score_of(c->sh->opt); j.min_seed_len = c->sh->opt.min_seed_len;
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
	for (int64_t i = 0; i < n; ++i) if (rows[i] > c->sh->ix.seq_len) return fail(LC_EARG, "BWT row out of range");
	DevTmp d;
	uint64_t *dr = d.up(rows, (size_t)n), *dout = d.alloc<uint64_t>((size_t)n);
	c->be.sync();
	c->be.ev_record(12);
	c->be.k_sa_batch(c->sh->ix, dr, dout, n, nullptr);
	c->be.ev_record(13);
	std::string e;
	if (c->be.check(e)) return fail(LC_ECUDA, e);
	cudaMemcpy(out, dout, 8 * (size_t)n, cudaMemcpyDeviceToHost);
	if (ms) *ms = c->be.ev_ms(12, 13);
	return LC_OK;
}

} // extern "C"
