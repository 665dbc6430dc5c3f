// lc_backend.cuh — CUDA backend of LcPipeline: device memory, one stream, CUDA events, kernel launches.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "lc_pipeline.hpp"

#define LC_N_EVENTS 16

struct LcExtJobs {      // batch of independent ksw_extend2 jobs (microbenchmark / parity entry point)
	int n; const uint8_t *q; const int64_t *qoff; const uint8_t *t; const int64_t *toff; const int32_t *h0;
	int w, end_bonus, zdrop; LcScore sc; int32_t *out; int qmax; unsigned int *ctr; LcCounters *cnt;
};
struct LcSwJobs {       // batch of independent ksw_align2 jobs with mem_matesw's flags
	int n; const uint8_t *q; const int64_t *qoff; const uint8_t *t; const int64_t *toff;
	LcScore sc; int min_seed_len; int32_t *out; int qmax; unsigned int *ctr; unsigned long long *w_u64; int w_u64_n; LcCounters *cnt;
};

struct LcCudaBackend {
	int device, n_sm = 0, workers = 0;
	cudaStream_t stream = 0;
	cudaEvent_t ev[LC_N_EVENTS];
	cudaStream_t aux[LC_CIG_NCLASS] = {};      // side streams: the band classes of the CIGAR stage run concurrently (fork/join on events)
	cudaEvent_t ev_fork = 0, ev_join[LC_CIG_NCLASS] = {};
	cudaStream_t cpy = 0; cudaEvent_t ev_cpy = 0;      // copy stream: the regions' D2H runs under the CIGAR kernels (LcPipeline::fetch_regs_into_side)
	void *scan_tmp = nullptr; size_t scan_tmp_bytes = 0;
	void *seed_glists = nullptr; size_t seed_glists_bytes = 0;      // interval lists of the later seeding passes (smem_kernel<.., GL>)
	// housekeeping gates of the state-machine kernels (lc_kernels.cu): serve waiting lanes every (mask + 1) ticks or when n lanes wait
	int seed_gate_mask = 3, seed_gate_n = 12, ext_gate_mask = 15, ext_row_mask = 1, ext_row_n = 33;
	int seed_gl_ctas = 8, seed_redo_ctas = 4, chain_ctas = 7, chain_heavy_ctas = 12;      // resident CTAs per SM of the latency-bound kernels (tuning knobs)

	explicit LcCudaBackend(int device_);
	~LcCudaBackend();
	int n_workers() const { return workers; }
	int n_redo_workers() const { return n_sm * 128; }     // smem_redo_kernel: one CTA of 128 threads per SM
	void *dmalloc(size_t n);
	void dfree(void *p);
	void dzero(void *p, size_t n);
	void dfill(void *p, int byte, size_t n);
	void k_iota(int32_t *p, int n);
	void dzero_words(unsigned long long *p, size_t n);
	void h2d(void *d, const void *s, size_t n);
	void d2h(void *d, const void *s, size_t n);
	void d2d(void *d, const void *s, size_t n);
	void sync();
	void copy_fork();                                  // what is enqueued on the copy stream from here on waits for the main stream's work so far
	void d2h_copy(void *d, const void *s, size_t n);
	void copy_sync();
	void ev_record(int i);
	float ev_ms(int a, int b);
	int grid_for(unsigned long long items, int tpb, int ctas_per_sm) const;
	void k_convert(uint8_t *p, unsigned long long n, bool ascii);
	void k_seed(const LcBatch &b, const int32_t *list, int n_items);      // list: read ids of a second pass, null = reads 0..n_items-1
	int seed_grid(const LcBatch &b, int n_items, bool listed);
	int seed_pass_lanes(const LcBatch &b, int n_items);                   // threads k_seed would launch for this pass (sizes the per-lane seed lists)
	int seed_lanes();
	void k_seed_finish(const LcBatch &b);
	void k_seed_redo(const LcBatch &b, int n_redo);
	void k_occ_convert(const uint32_t *bwa, uint4 *occ, unsigned long long n_blocks, const LcIndex &ix);
	void k_sa_fill(const LcIndex &ix, unsigned long long n_samples, void *out, bool wide);
	void k_sa(const LcBatch &b, unsigned long long n);
	void k_chain(const LcBatch &b);
	int k_extend(const LcBatch &b);                 // returns the number of kernels launched
	void k_pair(const LcBatch &b, int n_pairs);
	const uint64_t *sort_worklist(const LcBatch &b, const unsigned int *n_dev, unsigned int &n);
	void k_cigar(const LcBatch &b, const uint64_t *list, unsigned long long n, const unsigned int *n_dev);
	void k_cigar_plan(const LcBatch &b, unsigned long long n);
	int k_cigar_band(const LcBatch &b, unsigned long long n_max);
	void k_cigar_fallback(const LcBatch &b, unsigned long long n_max);
	void k_alnx(const LcBatch &b, unsigned long long n);
	void k_pack(const LcBatch &b, const LcPackArgs &pa);
	void k_rebase(int64_t *off, int n, int64_t off0);
	void scan_exclusive(const int32_t *in, int64_t *out, int n);
	void k_ksw_extend(const LcExtJobs &j);
	void k_ksw_align(const LcSwJobs &j);
	void k_sa_batch(const LcIndex &ix, const uint64_t *rows, uint64_t *out, long long n, LcCounters *cnt);
	int check(std::string &err);
	// first launch error seen since the last check(): a failed launch leaves a stage's outputs uninitialised, so the pipeline
	// asks after every host round trip instead of carrying on with garbage counts
	cudaError_t launch_err = cudaSuccess;
	void note_launch() { const cudaError_t e = cudaGetLastError(); if (e != cudaSuccess && launch_err == cudaSuccess) launch_err = e; }
	bool launch_failed() { note_launch(); return launch_err != cudaSuccess; }
};
