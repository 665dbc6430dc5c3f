// lc_kernels.cu — the CUDA kernels of libcuda_lariat.so (sm_100a) and the backend that launches them.
//
// One kernel per pipeline stage (SURVEY.md §2 "kernels the new framework must write"):
//   convert_kernel      ASCII -> 2-bit codes                                   (G/gobwa.go:159-167)
//   smem_kernel         mem_collect_intv: SMEM / re-seed / LAST-like seeding   (B/bwamem.c:114-162, B/bwt.c:262-379)
//   sa_lookup_kernel    bwt_sa for every scheduled occurrence                  (B/bwt.c:53-59,86-129; B/bwamem.c:283)
//   chain_kernel        mem_chain tree insert/merge + mem_chain_flt            (B/bwamem.c:190-385)
//   extend_kernel       mem_chain2aln -> ksw_extend2, one lane per read        (B/bwamem.c:632-786, B/ksw.c:380-479)
//   extend_heavy_kernel the same for reads with very many chains, one warp per read
//   matesw_kernel       mem_sort_dedup_patch + the two rescue phases + mem_matesw/ksw_align2
//                                                                              (B/bwamem.c:437-489, G/gobwa.go:264-324, B/bwamem_pair.c:111-180)
//   cigar_plan_kernel / cigar_band_kernel<NB> / global_cigar_kernel
//                       mem_reg2aln -> bwa_gen_cigar2 -> ksw_global2           (B/bwamem.c:1086-1156, B/bwa.c:121-207, B/ksw.c:504-606)
//   pack_kernel         dense result arrays in barcode/read order
// All are integer/byte kernels; grids are sized in multiples of the SM count and stride over the batch.
#include <cuda_runtime.h>
#include <mutex>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include "lc_backend.cuh"
#include "lc_sw_warp.cuh"
#include "lc_seed.cuh"
#include "lc_extsm.cuh"
#include "lc_cigar.cuh"

#define LC_TPB 128
#define LC_SEED_TPB 64
#define LC_SEED_GL_MIN 24      // list capacity above which a pass over listed reads keeps its lists in global memory
#define LC_EXT_TPB 64

__global__ void __launch_bounds__(256) convert_kernel(uint8_t *p, unsigned long long n, int is_ascii)
{
	// 16 bytes per thread per step
	unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 16, stride = (unsigned long long)gridDim.x * blockDim.x * 16;
	for (; i < n; i += stride) {
		if (i + 16 <= n && (((size_t)(p + i)) & 15) == 0) {
			uint4 v = *(const uint4 *)(p + i);
			unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				unsigned o = 0;
#pragma unroll
				for (int s = 0; s < 32; s += 8) o |= (unsigned)lc_nt4((uint8_t)(w[k] >> s), is_ascii) << s;
				w[k] = o;
			}
			*(uint4 *)(p + i) = make_uint4(w[0], w[1], w[2], w[3]);
		} else {
			for (unsigned long long j = i; j < n && j < i + 16; ++j) p[j] = lc_nt4(p[j], is_ascii);
		}
	}
}

// ---- index preparation (lc_init): the .bwt payload re-laid as 32-byte blocks, and the dense suffix array ------------
struct LcSbTable { uint64_t sb[LC_MAX_SB][4]; };
__global__ void __launch_bounds__(256) occ_convert_kernel(const uint32_t *bwa, uint4 *occ, unsigned long long n_blocks, const LcSbTable t)
{
	unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (; j < n_blocks; j += stride) {
		uint4 o[2];
		lc_occ_from_bwa(bwa, j, t.sb[j >> LC_SB_SHIFT], o);
		occ[2 * j] = o[0]; occ[2 * j + 1] = o[1];
	}
}
// every sampled row walks the text backwards and labels the rows it passes until it meets the next sampled row:
// one LF step per BWT row in total (the reference's bwt_sa pays ~31 steps per LOOKUP, B/bwt.c:86-96)
template <typename T>
__global__ void __launch_bounds__(256) sa_fill_kernel(const LcIndex ix, unsigned long long n_samples, T *out)
{
	unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (; j < n_samples; j += stride) lc_sa_fill_sample(ix, j, out);
}

// 8 bases of read `rr` per lane -> one word of 4-bit codes at dst[t * dst_stride] (three aligned words, funnel-shifted by the
// source misalignment, nibbles compacted).  Whole warp; the caller synchronises.
__device__ __forceinline__ void lc_warp_pack_read(const LcBatch &b, int rr, int lane, uint32_t *dst, int dst_stride)
{
	const int64_t o = b.off[rr];
	const int len = (int)(b.off[rr + 1] - o);
	const uint8_t *src = b.seq + o;
	const int mis = (int)((size_t)src & 3), nw = (len + 7) >> 3;
	const uint32_t *w = (const uint32_t *)(src - mis);
	for (int t = lane; t < nw; t += 32) {
		const uint32_t w0 = __ldg(w + 2 * t), w1 = __ldg(w + 2 * t + 1), w2 = __ldg(w + 2 * t + 2);
		uint32_t lo = __funnelshift_r(w0, w1, mis * 8), hi = __funnelshift_r(w1, w2, mis * 8);
		lo &= 0x0f0f0f0fu; lo = (lo | lo >> 4) & 0x00ff00ffu; lo = (lo | lo >> 8) & 0xffffu;
		hi &= 0x0f0f0f0fu; hi = (hi | hi >> 4) & 0x00ff00ffu; hi = (hi | hi >> 8) & 0xffffu;
		dst[(size_t)t * dst_stride] = lo | hi << 16;
	}
}

// ---- seeding ---------------------------------------------------------------------------------------------------
// One lane per read, one bwt_extend per tick (lc_seed.cuh).  Reads are dealt through an atomic counter; a lane that needs a
// new read gets it copied into its slot by the whole warp (word-wise, coalesced).
// Housekeeping (resolve() and the read fetch) is divergent code that runs with a handful of live lanes; running it after every
// tick — some lane of 32 nearly always wants it — made it a third of the kernel's instructions (profiles/r01m).  Lanes that
// need it now wait (their tick() is a no-op on a stale interval) until `gate_n` lanes are waiting or the tick number is a
// multiple of gate_mask + 1, and are then served together.
// GL (the later passes: reads in repeats whose lists need up to a read length of entries): the lists live in global memory — the same
// [word][lane] layout per warp, so lanes at the same depth touch one 128-byte line — and only the reads stay in shared memory.  With
// 150-entry lists in shared memory two warps fit an SM and the pass crawls at the latency of its dependent FM-index loads
// (profiles/r2f: 59 ms for the third pass of the 3.1 Gb config); a list entry is written once and read once, sequentially, and is
// not what a tick waits for.
template <typename U, bool COUNT, bool GL>
__global__ void __launch_bounds__(LC_SEED_TPB) smem_kernel(const LcBatch b, const unsigned gate_mask, const int gate_n, const int32_t *list, const int n_items, uint32_t *glists)
{   // list: read ids of a second pass (the reads whose lists outgrew the first pass' slots), null: reads 0 .. n_items-1
	extern __shared__ uint32_t lc_seed_smem[];
	const int lane = threadIdx.x & 31;
	uint32_t *wbase, *qbase;
	if (GL) {
		wbase = glists + ((size_t)blockIdx.x * (LC_SEED_TPB / 32) + (threadIdx.x >> 5)) * (size_t)(LC_SEED_ENT_WORDS * b.lst_cap * 32);
		qbase = lc_seed_smem + (size_t)(threadIdx.x >> 5) * lc_seed_q_words(b.lmax) * 32;
	} else {
		wbase = lc_seed_smem + (size_t)(threadIdx.x >> 5) * lc_seed_warp_words(b.lmax, b.lst_cap);      // this warp's lists, then its reads
		qbase = wbase + LC_SEED_ENT_WORDS * b.lst_cap * 32;
	}
	LcPk *memv = b.w_pk + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * b.pk_cap;
	LcSeedLane<U, COUNT, 32> sd(b, wbase + lane, b.lst_cap, qbase + lane, memv, b.pk_cap);
	for (unsigned it = 0;; ++it) {
		const unsigned pend = __ballot_sync(LC_FULL, sd.phase < LC_SP_FWD);               // waiting for resolve() or for a read
		if (pend && ((it & gate_mask) == 0 || __popc(pend) >= gate_n)) {                  // warp-uniform
			if (__any_sync(LC_FULL, sd.phase < LC_SP_FWD && sd.phase != LC_SP_FETCH)) sd.resolve();
			unsigned need = __ballot_sync(LC_FULL, sd.phase == LC_SP_FETCH);
			while (need) {
				int base = 0;
				if (lane == 0) base = (int)atomicAdd(b.work_ctr + 3, (unsigned)__popc(need));
				base = __shfl_sync(LC_FULL, base, 0);
				int myr = -1;
				if (need >> lane & 1) { myr = base + __popc(need & ((1u << lane) - 1)); myr = myr >= n_items ? -1 : (list ? list[myr] : myr); }
				for (unsigned m = need; m; m &= m - 1) {
					const int L = __ffs(m) - 1;
					const int rr = __shfl_sync(LC_FULL, myr, L);
					if (rr < 0) continue;
					const int64_t o = b.off[rr];
					const int len = (int)(b.off[rr + 1] - o);
					if (len < b.opt.min_seed_len) continue;
					lc_warp_pack_read(b, rr, lane, qbase + L, 32);
				}
				__syncwarp();
				if (need >> lane & 1) {
					if (myr < 0) sd.phase = LC_SP_DONE;
					else {
						const int64_t o = b.off[myr];
						const int len = (int)(b.off[myr + 1] - o);
						if (len < b.opt.min_seed_len) { b.r_n_pk[myr] = 0; b.r_pk_off[myr] = 0; }      // B/bwamem.c:260; stays in FETCH
						else sd.begin_read(myr, len);
					}
				}
				__syncwarp();
				if (__any_sync(LC_FULL, sd.phase < LC_SP_FWD && sd.phase != LC_SP_FETCH)) sd.resolve();     // the new reads' first steps
				need = __ballot_sync(LC_FULL, sd.phase == LC_SP_FETCH);
			}
		}
		if (!__ballot_sync(LC_FULL, sd.phase != LC_SP_DONE)) break;
		sd.tick();
	}
}

// 4 lanes per read: sort the seeds, lay out the occurrence schedule, fetch positions (lc_seed_finish)
struct LcCoopQuad {
	int lane, gbase;
	unsigned gmask;
	__device__ __forceinline__ bool leader() const { return (lane & 3) == 0; }
	__device__ __forceinline__ int sub() const { return lane & 3; }
	__device__ __forceinline__ int width() const { return 4; }
	__device__ __forceinline__ void sync() const { __syncwarp(gmask); }
	__device__ __forceinline__ unsigned long long bcast(unsigned long long v) const { return __shfl_sync(gmask, v, gbase); }
};
template <typename U>
__global__ void __launch_bounds__(256) seed_finish_kernel(const LcBatch b)
{
	LcCoopQuad co;
	co.lane = threadIdx.x & 31; co.gbase = co.lane & ~3; co.gmask = 0xfu << co.gbase;
	const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, ng = (gridDim.x * blockDim.x) >> 2;
	for (int r = g; r < b.n_reads; r += ng) lc_seed_finish<U>(b, co, r);
}

// reads the state machine could not hold in its slots — in practice the reads inside high-copy repeats, whose interval lists run to a
// hundred entries: one WARP per read (lc_smem1_t: the backward phase extends the whole list in parallel), lists in global memory
struct LcSeedExecWarp {
	int ln;
	__device__ __forceinline__ int lane_id() const { return ln; }
	__device__ __forceinline__ int n_lanes() const { return 32; }
	__device__ __forceinline__ bool leader() const { return ln == 0; }
	__device__ __forceinline__ void sync() const { __syncwarp(); }
	__device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(LC_FULL, p); }
	__device__ __forceinline__ uint64_t shfl(uint64_t v, int src) const { return __shfl_sync(LC_FULL, (unsigned long long)v, src); }
};
__global__ void __launch_bounds__(LC_TPB) smem_redo_kernel(const LcBatch b, int n_redo)
{
	LcSeedExecWarp ex;
	ex.ln = threadIdx.x & 31;
	const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	for (;;) {
		unsigned int t = 0;
		if (ex.ln == 0) t = atomicAdd(b.work_ctr + 3, 1u);
		t = __shfl_sync(LC_FULL, t, 0);
		if (t >= (unsigned int)n_redo) break;
		lc_stage_seed_t(b, b.redo[t], gwarp, ex);
		__syncwarp();
	}
}

__global__ void __launch_bounds__(256) sa_lookup_kernel(const LcBatch b, unsigned long long n)
{
	unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (; g < n; g += stride) lc_stage_sa(b, g);
}

__global__ void __launch_bounds__(LC_TPB) chain_kernel(const LcBatch b)
{
	const int worker = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	for (int r = worker; r < b.n_reads; r += stride) lc_stage_chain(b, r);
}

// Stage 3b: the reads chain_kernel listed (very many seeds: repeats), one warp each (LcChainJob::run, lc_mem.cuh)
struct LcChainExecWarp {
	int ln;
	__device__ __forceinline__ int lane() const { return ln; }
	__device__ __forceinline__ int width() const { return 32; }
	__device__ __forceinline__ bool leader() const { return ln == 0; }
	__device__ __forceinline__ void sync() const { __syncwarp(); }
	__device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(LC_FULL, p); }
	__device__ __forceinline__ int bcast(int v) const { return __shfl_sync(LC_FULL, v, 0); }
};
__global__ void __launch_bounds__(LC_TPB) chain_heavy_kernel(const LcBatch b)
{
	LcChainExecWarp ex;
	ex.ln = threadIdx.x & 31;
	const unsigned int n = b.chain_heavy_n[0];      // written by chain_kernel, which finished before this launch started
	for (;;) {
		unsigned int t = 0;
		if (ex.ln == 0) t = atomicAdd(b.chain_heavy_n + 1, 1u);
		t = __shfl_sync(LC_FULL, t, 0);
		if (t >= n) break;
		lc_stage_chain_heavy_t(b, b.chain_list[t], ex);
		__syncwarp();
	}
}

// Stage 4: one lane per read, mem_chain2aln + ksw_extend2 as a state machine (lc_extsm.cuh).  Per warp the shared memory holds
// the 32 H/E rows ([column][lane], packed by CELLS) and the 32 reads (4 bits per base, [word][lane], one spare word in front).
template <class CELLS> __host__ __device__ inline int lc_ext_warp_words(int lmax) { return (CELLS::words(lc_ext_he_cols(lmax)) + lc_ext_q_words(lmax)) * 32; }
#define LC_EXT_MTAB_WORDS 64      // lc_ext_fill_masks table (36 words) in front of the per-warp areas
// Gates (as in smem_kernel): the read fetch and resolve() run on ticks that are multiples of hk_mask + 1, the end-of-row work on
// multiples of row_mask + 1; lanes that need either idle until then, so that the divergent code runs with several live lanes.
// UNIT: the work items are the units lc_ext_plan made of the reads with very many chains (b.ext_units) instead of reads.
template <class CELLS, bool UNIT>
__global__ void __launch_bounds__(LC_EXT_TPB, 8) extend_kernel(const LcBatch b, const unsigned hk_mask, const unsigned row_mask)
{
	extern __shared__ uint32_t lc_ext_smem[];
	const int lane = threadIdx.x & 31;
	uint32_t *wbase = lc_ext_smem + LC_EXT_MTAB_WORDS + (size_t)(threadIdx.x >> 5) * lc_ext_warp_words<CELLS>(b.lmax);
	uint32_t *qf = wbase + (CELLS::words(lc_ext_he_cols(b.lmax)) + 1) * 32;         // word 0 of the read copies; word -1 is the spare
	if (threadIdx.x == 0) lc_ext_fill_masks(lc_ext_smem);
	__syncthreads();
	LcExtLane<32, CELLS, UNIT> ex(b, wbase + lane, qf + lane, lc_ext_smem);
	const int n_items = UNIT ? lc_ext_n_units(b) : b.n_reads;
	unsigned int *ctr = UNIT ? b.ext_unit_n + 1 : b.work_ctr + 0;
	const unsigned rmask = row_mask & 255u, row_n = row_mask >> 8;
	for (unsigned it = 0;; ++it) {
		if ((it & hk_mask) == 0) {                               // warp-uniform
			const unsigned need = __ballot_sync(LC_FULL, ex.phase == LC_XP_FETCH);
			if (need) {
				int base = 0;
				if (lane == 0) base = (int)atomicAdd(ctr, (unsigned)__popc(need));
				base = __shfl_sync(LC_FULL, base, 0);
				int myi = -1, myr = -1;
				if (need >> lane & 1) { myi = base + __popc(need & ((1u << lane) - 1)); if (myi >= n_items) myi = -1; }
				const long long myu = (UNIT && myi >= 0) ? (long long)lc_ext_unit_at(b, myi) : -1;
				if (myi >= 0) myr = UNIT ? b.chain_aux[myu].r : myi;
				for (unsigned m = need; m; m &= m - 1) {
					const int L = __ffs(m) - 1;
					const int rr = __shfl_sync(LC_FULL, myr, L);
					if (rr >= 0 && b.r_n_chain[rr] > 0) lc_warp_pack_read(b, rr, lane, qf + L, 32);
				}
				__syncwarp();
				if (need >> lane & 1) {
					if (myi < 0) ex.phase = LC_XP_DONE;
					else if (UNIT) ex.begin_unit(myu);
					else ex.begin_read(myr);
				}
			}
			if (!__ballot_sync(LC_FULL, ex.phase != LC_XP_DONE)) break;
			if (__any_sync(LC_FULL, ex.phase < LC_XP_DP && ex.phase != LC_XP_FETCH)) ex.resolve();
		}
		// the end-of-row gate: every (mask + 1) ticks, or as soon as row_n lanes wait (row_n > 32: never) — warp-uniform either way
		ex.tick((it & rmask) == 0 || (row_n <= 32 && __popc(__ballot_sync(LC_FULL, ex.phase == LC_XP_ROWEND)) >= (int)row_n));
	}
}

// Stage 4b: the reads extend_kernel set aside (very many chains) are cut into units (lc_ext_plan, one thread per listed read) ...
__global__ void __launch_bounds__(LC_TPB) ext_plan_kernel(const LcBatch b, const unsigned int n)
{
	for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) lc_ext_plan(b, (int)(uint32_t)b.wl_in[t]);
}
// ... most of which go through extend_kernel<CELLS, true>; a unit that is itself very large (many overlapping chains: tandem arrays) gets
// a warp: cooperative ksw_extend2 rows (lc_sw_warp.cuh)
// list == null: the items are the reads 0 .. n-1, whole (small batches: every read gets a warp)
__global__ void __launch_bounds__(LC_TPB) extend_heavy_kernel(const LcBatch b, const uint64_t *list, const unsigned int n)
{
	extern __shared__ int lc_smem[];
	const int lane = threadIdx.x & 31, per = lc_warp_smem_bytes(b.lmax), rows = lc_warp_row_ints(b.lmax);
	LcExecWarp ex;
	ex.H = (int *)((char *)lc_smem + (threadIdx.x >> 5) * per); ex.E = ex.H + rows; ex.qs = (uint8_t *)(ex.E + rows); ex.lane = lane;
	ex.Hmax = nullptr; ex.bl = nullptr; ex.cig = nullptr; ex.zloc = nullptr;
	for (;;) {
		unsigned int t = 0;
		if (lane == 0) t = atomicAdd(b.ext_hvu_n + 1, 1u);
		t = __shfl_sync(LC_FULL, t, 0);
		if (t >= n) break;
		if (list) lc_stage_extend_t(b, -1, (long long)(uint32_t)list[t], ex);
		else lc_stage_extend_t(b, (int)t, -1, ex);
		__syncwarp();
	}
}

// Stage 5a: one thread per pair — dedup of both ends, final lists, and the rescue phases for as long as they are no-ops
__global__ void __launch_bounds__(LC_TPB) pair_prepare_kernel(const LcBatch b, int n_pairs)
{
	const int worker = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
	int *wi = b.w_int + (size_t)worker * b.w_int_n;
	for (int p = worker; p < n_pairs; p += stride) if (!lc_pair_prepare_defer(b, p)) lc_pair_prepare(b, p, wi, wi + (b.lmax + 18));
}
// ... pairs with very many regions (repeats) were listed instead: a warp each, the same body with the copies and list scans shared out
__global__ void __launch_bounds__(LC_TPB) pair_prepare_heavy_kernel(const LcBatch b)
{
	LcExecWarp ex;
	ex.H = ex.E = ex.Hmax = nullptr; ex.qs = nullptr; ex.bl = nullptr; ex.cig = nullptr; ex.zloc = nullptr;
	ex.lane = threadIdx.x & 31;
	const int worker = (blockIdx.x * blockDim.x + threadIdx.x) & ~31;      // the warp uses the scratch slot of its first thread
	int *wi = b.w_int + (size_t)worker * b.w_int_n;
	const unsigned int n = b.prep_heavy_n[0];
	for (;;) {
		unsigned int t = 0;
		if (ex.lane == 0) t = atomicAdd(b.prep_heavy_n + 1, 1u);
		t = __shfl_sync(LC_FULL, t, 0);
		if (t >= n) break;
		lc_pair_prepare_t(b, (int)(uint32_t)b.wl_out[t], wi, wi + (b.lmax + 18), ex);
		__syncwarp();
	}
}
// Stage 5b: one warp per pair that needs a Smith-Waterman — the two rescue phases with the warp-cooperative local SW
__global__ void __launch_bounds__(LC_TPB, 4) matesw_kernel(const LcBatch b, const uint64_t *list, const unsigned int n)
{
	extern __shared__ int lc_smem[];
	const int lane = threadIdx.x & 31, per = lc_warp_smem_bytes(b.lmax, 3), rows = lc_warp_row_ints(b.lmax);
	const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	LcExecWarp ex;
	ex.H = (int *)((char *)lc_smem + (threadIdx.x >> 5) * per); ex.E = ex.H + rows; ex.Hmax = ex.E + rows; ex.qs = (uint8_t *)(ex.Hmax + rows);
	ex.lane = lane; ex.bl = b.w_u64 + (size_t)gwarp * b.w_u64_n;
	for (;;) {
		unsigned int t = 0;
		if (lane == 0) t = atomicAdd(b.work_ctr + 1, 1u);
		t = __shfl_sync(LC_FULL, t, 0);
		if (t >= n) break;
		lc_pair_rescue_t(b, (int)(uint32_t)list[t], (int)(list[t] >> 48) & 3, ex);
		__syncwarp();
	}
}

// Stage 6 (lc_cigar.cuh): one thread per candidate finishes the gap-free ones and bins the rest by band width ...
__global__ void __launch_bounds__(256) cigar_plan_kernel(const LcBatch b, unsigned long long n)
{
	unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (; c < n; c += stride) lc_cig_plan(b, b.cand[c]);
}
// ... one thread per binned job runs the banded global alignment with the band in registers, traces back and writes the CIGAR ...
template <int NB>
__global__ void __launch_bounds__(128) cigar_band_kernel(const LcBatch b, int cls)
{
	const unsigned int n = b.cig_n[cls];
	const uint64_t *jobs = b.cig_jobs + (size_t)cls * b.cig_jobs_cap;
	for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) lc_cig_band<NB>(b, jobs[t]);
}
// ... and whatever those two hand back (wide bands, band-doubling retries, degenerate regions) gets a warp per candidate:
// gap-free candidates are a cooperative compare, the rest a cooperative banded NW + lane-0 traceback.
// n_dev != null: the list length is read from device memory (no host round trip between the kernels of the stage).
__global__ void __launch_bounds__(LC_TPB) global_cigar_kernel(const LcBatch b, const uint64_t *list, unsigned long long n_host, const unsigned int *n_dev)
{
	extern __shared__ int lc_smem[];
	const int lane = threadIdx.x & 31, per = lc_warp_smem_bytes(b.lmax), rows = lc_warp_row_ints(b.lmax);
	const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const unsigned long long n = n_dev ? (unsigned long long)*n_dev : n_host;
	LcExecWarp ex;
	ex.H = (int *)((char *)lc_smem + (threadIdx.x >> 5) * per); ex.E = ex.H + rows; ex.qs = (uint8_t *)(ex.E + rows); ex.lane = lane;
	ex.Hmax = nullptr; ex.bl = nullptr;
	ex.cig = (uint32_t *)(b.w_int + (size_t)gwarp * b.w_int_n); ex.zloc = b.w_z + (size_t)gwarp * b.w_z_n;
	for (;;) {
		unsigned int c = 0;
		if (lane == 0) c = atomicAdd(b.work_ctr + 2, 1u);
		c = __shfl_sync(LC_FULL, c, 0);
		if (c >= n) break;
		lc_stage_cigar_t(b, list[c], ex);
		__syncwarp();
	}
}

// Stage 6b (lc_post.cuh): one thread per candidate — the CIGAR walk of inference.GetAlignments (I/lariat.go:1573-1650)
__global__ void __launch_bounds__(256) alnx_kernel(const LcBatch b, unsigned long long n)
{
	unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (unsigned long long)gridDim.x * blockDim.x;
	for (; c < n; c += stride) lc_stage_alnx(b, b.cand[c]);
}

__global__ void __launch_bounds__(256) pack_kernel(const LcBatch b, const LcPackArgs pa)
{
	const int stride = gridDim.x * blockDim.x;
	for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= b.n_reads; r += stride) lc_stage_pack(b, pa, r);
}
__global__ void __launch_bounds__(256) rebase_kernel(int64_t *off, int n, int64_t off0)
{
	const int stride = gridDim.x * blockDim.x;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) off[i] -= off0;
}

// ---- microbenchmark kernels (BASELINE.json configs[4]) -------------------------------------------------------
// ksw_extend2 on independent jobs, one warp per job: the warp-cooperative routine (lc_extend2_warp) that extend_heavy_kernel and the
// low-latency route use — NOT the per-lane state machine of extend_kernel, which tests/hostsim and the end-to-end dumps cover
__global__ void __launch_bounds__(LC_TPB) ksw_extend_kernel(const LcExtJobs j)
{
	extern __shared__ int lc_smem[];
	const int lane = threadIdx.x & 31, per = lc_warp_smem_bytes(j.qmax), rows = lc_warp_row_ints(j.qmax);
	int *H = (int *)((char *)lc_smem + (threadIdx.x >> 5) * per), *E = H + rows;
	uint8_t *qs = (uint8_t *)(E + rows);
	for (;;) {
		int i = 0;
		if (lane == 0) i = (int)atomicAdd(j.ctr, 1u);
		i = __shfl_sync(LC_FULL, i, 0);
		if (i >= j.n) break;
		LcSeqFwd qf = { j.q + j.qoff[i] };
		LcSeqFwd tf = { j.t + j.toff[i] };
		LcExtResult r = lc_extend2_warp((int)(j.qoff[i + 1] - j.qoff[i]), qf, (int)(j.toff[i + 1] - j.toff[i]), tf, j.sc, j.w, j.end_bonus, j.zdrop,
										j.h0[i], H, E, qs, lane, j.cnt);
		if (lane == 0) {
			int32_t *o = j.out + 6 * (size_t)i;
			o[0] = r.score; o[1] = r.qle; o[2] = r.tle; o[3] = r.gtle; o[4] = r.gscore; o[5] = r.max_off;
		}
		__syncwarp();
	}
}

// ksw_align2 on independent jobs: the warp-cooperative routine matesw_kernel uses, one warp per job
__global__ void __launch_bounds__(LC_TPB) ksw_align_kernel(const LcSwJobs j)
{
	extern __shared__ int lc_smem[];
	const int lane = threadIdx.x & 31, per = lc_warp_smem_bytes(j.qmax, 3), rows = lc_warp_row_ints(j.qmax);
	const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	int *H = (int *)((char *)lc_smem + (threadIdx.x >> 5) * per), *E = H + rows, *Hmax = E + rows;
	uint8_t *qs = (uint8_t *)(Hmax + rows);
	unsigned long long *bl = j.w_u64 + (size_t)gwarp * j.w_u64_n;
	for (;;) {
		int i = 0;
		if (lane == 0) i = (int)atomicAdd(j.ctr, 1u);
		i = __shfl_sync(LC_FULL, i, 0);
		if (i >= j.n) break;
		const int ql = (int)(j.qoff[i + 1] - j.qoff[i]), tl = (int)(j.toff[i + 1] - j.toff[i]);
		LcSeqFwd qf = { j.q + j.qoff[i] };
		LcSeqFwd tf = { j.t + j.toff[i] };
		const int xtra = LC_XSUBO | LC_XSTART | (ql * j.sc.a < 250 ? LC_XBYTE : 0) | (j.min_seed_len * j.sc.a);
		LcSwResult r = lc_sw_align2_warp(ql, qf, tl, tf, j.sc, xtra, H, E, Hmax, qs, bl, lane, j.cnt);
		if (lane == 0) {
			int32_t *o = j.out + 7 * (size_t)i;
			o[0] = r.score; o[1] = r.te; o[2] = r.qe; o[3] = r.score2; o[4] = r.te2; o[5] = r.tb; o[6] = r.qb;
		}
		__syncwarp();
	}
}

__global__ void __launch_bounds__(256) sa_batch_kernel(const LcIndex ix, const uint64_t *rows, uint64_t *out, long long n, LcCounters *cnt)
{
	long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
	for (; g < n; g += stride) out[g] = lc_sa(ix, rows[g], cnt);
}

// ---- backend -------------------------------------------------------------------------------------------------
static std::once_flag lc_attr_once[64];
static void lc_set_kernel_attrs(int device);
struct LcCastI64 { __host__ __device__ int64_t operator()(const int32_t &x) const { return (int64_t)x; } };

LcCudaBackend::LcCudaBackend(int device_) : device(device_)
{
	cudaSetDevice(device);
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	n_sm = prop.multiProcessorCount;
	if (device >= 0 && device < 64) std::call_once(lc_attr_once[device], lc_set_kernel_attrs, device);
	cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
	for (int i = 0; i < LC_N_EVENTS; ++i) cudaEventCreate(&ev[i]);
	cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming);
	cudaStreamCreateWithFlags(&cpy, cudaStreamNonBlocking); cudaEventCreateWithFlags(&ev_cpy, cudaEventDisableTiming);
	for (int i = 0; i < LC_CIG_NCLASS; ++i) { cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking); cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming); }
	workers = n_sm * 4 * LC_TPB;     // 4 resident CTAs of 128 threads per SM
	auto env_int = [](const char *name, int &v, int lo, int hi) { if (const char *e = getenv(name)) { const int x = atoi(e); if (x >= lo && x <= hi) v = x; } };
	env_int("LC_SEED_GL_CTAS", seed_gl_ctas, 1, 32); env_int("LC_SEED_REDO_CTAS", seed_redo_ctas, 1, 16); env_int("LC_CHAIN_CTAS", chain_ctas, 1, 16); env_int("LC_CHAIN_HEAVY_CTAS", chain_heavy_ctas, 1, 16);
	env_int("LC_SEED_GATE_MASK", seed_gate_mask, 0, 255); env_int("LC_SEED_GATE_N", seed_gate_n, 1, 33);     // tuning knobs; mask must be 2^k - 1
	env_int("LC_EXT_GATE_MASK", ext_gate_mask, 0, 255); env_int("LC_EXT_ROW_MASK", ext_row_mask, 0, 255); env_int("LC_EXT_ROW_N", ext_row_n, 1, 33);
}
LcCudaBackend::~LcCudaBackend()
{
	cudaSetDevice(device);
	if (scan_tmp) cudaFree(scan_tmp);
	if (seed_glists) cudaFree(seed_glists);
	for (int i = 0; i < LC_N_EVENTS; ++i) cudaEventDestroy(ev[i]);
	for (int i = 0; i < LC_CIG_NCLASS; ++i) { cudaEventDestroy(ev_join[i]); cudaStreamDestroy(aux[i]); }
	cudaEventDestroy(ev_fork);
	cudaEventDestroy(ev_cpy); cudaStreamDestroy(cpy);
	cudaStreamDestroy(stream);
}
void *LcCudaBackend::dmalloc(size_t n) { void *p = 0; if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) { cudaGetLastError(); return 0; } return p; }
void LcCudaBackend::dfree(void *p) { cudaStreamSynchronize(stream); cudaFree(p); }
void LcCudaBackend::dzero(void *p, size_t n) { cudaMemsetAsync(p, 0, n, stream); }
void LcCudaBackend::dfill(void *p, int byte, size_t n) { cudaMemsetAsync(p, byte, n, stream); }
__global__ void iota_kernel(int32_t *p, int n) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = i; }
void LcCudaBackend::k_iota(int32_t *p, int n) { if (n > 0) iota_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(p, n); }
void LcCudaBackend::dzero_words(unsigned long long *p, size_t n) { cudaMemsetAsync(p, 0, 8 * n, stream); }
void LcCudaBackend::h2d(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream); }
void LcCudaBackend::d2h(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream); }
void LcCudaBackend::d2d(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, stream); }
void LcCudaBackend::sync() { cudaStreamSynchronize(stream); }
void LcCudaBackend::copy_fork() { cudaEventRecord(ev_cpy, stream); cudaStreamWaitEvent(cpy, ev_cpy, 0); }
void LcCudaBackend::d2h_copy(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, cpy); }
void LcCudaBackend::copy_sync() { cudaStreamSynchronize(cpy); }
void LcCudaBackend::ev_record(int i) { cudaEventRecord(ev[i], stream); }
float LcCudaBackend::ev_ms(int a, int b) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[a], ev[b]) != cudaSuccess) { cudaGetLastError(); return 0; } return ms; }
int LcCudaBackend::grid_for(unsigned long long items, int tpb, int ctas_per_sm) const
{
	unsigned long long need = (items + tpb - 1) / tpb;
	unsigned long long cap = (unsigned long long)n_sm * ctas_per_sm;
	if (need >= cap) return (int)cap;
	// round up to a multiple of the SM count when the batch is big enough, else just what is needed
	return (int)(need ? need : 1);
}
void LcCudaBackend::k_convert(uint8_t *p, unsigned long long n, bool ascii)
{
	if (!n) return;
	convert_kernel<<<grid_for((n + 15) / 16, 256, 8), 256, 0, stream>>>(p, n, ascii ? 1 : 0);
}
static bool lc_wide(const LcBatch &b) { return b.ix.seq_len >= 0xffffffffull || b.force_wide; }
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is state of the FUNCTION (per device), not of a stream or a launch, and lanes,
// cloned contexts and coalescer leaders launch the same kernels from many host threads with different sizes.  Round 1 serialised
// every "set the limit, then launch" pair under one process-wide lock; now the limit of every kernel with dynamic shared memory is
// raised ONCE per device to the device's opt-in maximum (it is only a permission: the carve-out of a launch follows its own size),
// and launches need no lock at all.
static void lc_set_kernel_attrs(int device)
{
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return; }
	const int mx = (int)prop.sharedMemPerBlockOptin;
	auto set = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, mx); };
	set(smem_kernel<uint32_t, true, false>); set(smem_kernel<uint32_t, false, false>); set(smem_kernel<uint64_t, true, false>); set(smem_kernel<uint64_t, false, false>);
	set(extend_kernel<LcCells8<32>, false>); set(extend_kernel<LcCells16<32>, false>); set(extend_kernel<LcCells8<32>, true>); set(extend_kernel<LcCells16<32>, true>);
	set(extend_heavy_kernel); set(matesw_kernel); set(global_cigar_kernel); set(ksw_extend_kernel); set(ksw_align_kernel);
	cudaGetLastError();
}
// a pass over listed reads with long lists keeps the lists in global memory (smem_kernel<.., GL>)
static bool lc_seed_gl(const LcBatch &b, bool listed) { return listed && b.lst_cap > LC_SEED_GL_MIN; }
template <bool GL> static int lc_seed_grid_t(const LcBatch &b, int n_items, int n_sm, int gl_ctas, int &smem)
{
	smem = GL ? (LC_SEED_TPB / 32) * lc_seed_q_words(b.lmax) * 32 * 4 : (LC_SEED_TPB / 32) * lc_seed_warp_words(b.lmax, b.lst_cap) * 4;
	int per_sm = 1;
	auto prep = [&](auto kern) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LC_SEED_TPB, smem); };
	if (lc_wide(b)) { if (b.cnt) prep(smem_kernel<uint64_t, true, GL>); else prep(smem_kernel<uint64_t, false, GL>); }
	else { if (b.cnt) prep(smem_kernel<uint32_t, true, GL>); else prep(smem_kernel<uint32_t, false, GL>); }
	if (per_sm < 1) per_sm = 1;
	if (GL && per_sm > gl_ctas) per_sm = gl_ctas;      // the per-lane seed and list scratch is sized by the resident lanes
	const long long need = ((long long)n_items + LC_SEED_TPB - 1) / LC_SEED_TPB;
	const long long cap = (long long)n_sm * per_sm;      // persistent: one wave of resident CTAs, reads dealt by an atomic counter
	return (int)(need < cap ? (need ? need : 1) : cap);
}
int LcCudaBackend::seed_grid(const LcBatch &b, int n_items, bool listed)
{
	int smem;
	return lc_seed_gl(b, listed) ? lc_seed_grid_t<true>(b, n_items, n_sm, seed_gl_ctas, smem) : lc_seed_grid_t<false>(b, n_items, n_sm, seed_gl_ctas, smem);
}
int LcCudaBackend::seed_pass_lanes(const LcBatch &b, int n_items) { return seed_grid(b, n_items, true) * LC_SEED_TPB; }
template <bool GL> static void lc_launch_seed(const LcBatch &b, const int32_t *list, int n_items, int n_sm, int gl_ctas, cudaStream_t stream, unsigned gm, int gn, uint32_t *glists)
{
	int smem;
	const int grid = lc_seed_grid_t<GL>(b, n_items, n_sm, gl_ctas, smem);
	if (lc_wide(b)) {
		if (b.cnt) smem_kernel<uint64_t, true, GL><<<grid, LC_SEED_TPB, smem, stream>>>(b, gm, gn, list, n_items, glists);
		else smem_kernel<uint64_t, false, GL><<<grid, LC_SEED_TPB, smem, stream>>>(b, gm, gn, list, n_items, glists);
	} else {
		if (b.cnt) smem_kernel<uint32_t, true, GL><<<grid, LC_SEED_TPB, smem, stream>>>(b, gm, gn, list, n_items, glists);
		else smem_kernel<uint32_t, false, GL><<<grid, LC_SEED_TPB, smem, stream>>>(b, gm, gn, list, n_items, glists);
	}
}
void LcCudaBackend::k_seed(const LcBatch &b, const int32_t *list, int n_items)
{
	cudaMemsetAsync(b.work_ctr + 3, 0, 4, stream);
	const unsigned gm = (unsigned)seed_gate_mask;
	const int gn = seed_gate_n;
	if (lc_seed_gl(b, list != nullptr)) {
		const size_t need = (size_t)seed_grid(b, n_items, true) * (LC_SEED_TPB / 32) * (size_t)(LC_SEED_ENT_WORDS * b.lst_cap * 32) * 4;
		if (need > seed_glists_bytes) {
			if (seed_glists) { cudaStreamSynchronize(stream); cudaFree(seed_glists); }
			seed_glists = nullptr; seed_glists_bytes = 0;
			if (cudaMalloc(&seed_glists, need) != cudaSuccess) { cudaGetLastError(); seed_glists = nullptr; }
			else seed_glists_bytes = need;
		}
		if (seed_glists) { lc_launch_seed<true>(b, list, n_items, n_sm, seed_gl_ctas, stream, gm, gn, (uint32_t *)seed_glists); return; }
		// no memory for the lists: the shared-memory form below is slower but complete (seed_pass_lanes over-estimates its lanes, which is harmless)
	}
	lc_launch_seed<false>(b, list, n_items, n_sm, seed_gl_ctas, stream, gm, gn, nullptr);
}
void LcCudaBackend::k_seed_finish(const LcBatch &b)
{
	const int grid = grid_for((unsigned long long)b.n_reads * 4, 256, 8);
	if (lc_wide(b)) seed_finish_kernel<uint64_t><<<grid, 256, 0, stream>>>(b);
	else seed_finish_kernel<uint32_t><<<grid, 256, 0, stream>>>(b);
}
int LcCudaBackend::seed_lanes() { return n_sm * 2048; }      // upper bound on resident lanes of smem_kernel
void LcCudaBackend::k_seed_redo(const LcBatch &b, int n_redo)
{
	cudaMemsetAsync(b.work_ctr + 3, 0, 4, stream);
	smem_redo_kernel<<<grid_for((unsigned long long)n_redo * 32, LC_TPB, seed_redo_ctas), LC_TPB, 0, stream>>>(b, n_redo);      // at most n_sm * 64 warps <= worker slots (n_redo_workers)
}
void LcCudaBackend::k_occ_convert(const uint32_t *bwa, uint4 *occ, unsigned long long n_blocks, const LcIndex &ix)
{
	LcSbTable t;
	memcpy(t.sb, ix.sb, sizeof(t.sb));
	occ_convert_kernel<<<grid_for(n_blocks, 256, 8), 256, 0, stream>>>(bwa, occ, n_blocks, t);
}
void LcCudaBackend::k_sa_fill(const LcIndex &ix, unsigned long long n_samples, void *out, bool wide)
{
	if (wide) sa_fill_kernel<uint64_t><<<grid_for(n_samples, 256, 8), 256, 0, stream>>>(ix, n_samples, (uint64_t *)out);
	else sa_fill_kernel<uint32_t><<<grid_for(n_samples, 256, 8), 256, 0, stream>>>(ix, n_samples, (uint32_t *)out);
}
void LcCudaBackend::k_sa(const LcBatch &b, unsigned long long n) { sa_lookup_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(b, n); }
void LcCudaBackend::k_chain(const LcBatch &b)
{
	chain_kernel<<<grid_for(b.n_reads, LC_TPB, chain_ctas), LC_TPB, 0, stream>>>(b);
	// the list length is only known on the device: a resident-size grid whose warps pull listed reads from a counter (no host round trip)
	if (b.chain_heavy > 0) chain_heavy_kernel<<<grid_for((unsigned long long)b.n_reads * 32, LC_TPB, chain_heavy_ctas), LC_TPB, 0, stream>>>(b);
}
template <class CELLS, bool UNIT> static void lc_launch_extend(const LcBatch &b, long long n_items, int n_sm, cudaStream_t stream, unsigned hk_mask, unsigned row_mask)
{
	const int smem = ((LC_EXT_TPB / 32) * lc_ext_warp_words<CELLS>(b.lmax) + LC_EXT_MTAB_WORDS) * 4;
	int per_sm = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, extend_kernel<CELLS, UNIT>, LC_EXT_TPB, smem);
	if (per_sm < 1) per_sm = 1;
	const long long need = (n_items + LC_EXT_TPB - 1) / LC_EXT_TPB, cap = (long long)n_sm * per_sm;
	extend_kernel<CELLS, UNIT><<<(int)(need < cap ? (need ? need : 1) : cap), LC_EXT_TPB, smem, stream>>>(b, hk_mask, row_mask);
}
int LcCudaBackend::k_extend(const LcBatch &b)
{
	// byte-packed H/E cells whenever no score of the batch can exceed 255 (half the shared memory per warp, twice the resident warps)
	if (b.small) {      // few reads: a warp per read is the short way through (a lane of the state machine walks its read's cells eight at a time)
		const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(b.lmax);
		extend_heavy_kernel<<<grid_for((unsigned long long)b.n_reads * 32, LC_TPB, 4), LC_TPB, smem, stream>>>(b, nullptr, (unsigned int)b.n_reads);
		return 1;
	}
	const bool cells8 = lc_ext_cells8_ok(b) && !getenv("LC_EXT_WIDE_CELLS");
	const unsigned hk = (unsigned)ext_gate_mask, rw = (unsigned)ext_row_mask | (unsigned)ext_row_n << 8;
	if (cells8) lc_launch_extend<LcCells8<32>, false>(b, b.n_reads, n_sm, stream, hk, rw);
	else lc_launch_extend<LcCells16<32>, false>(b, b.n_reads, n_sm, stream, hk, rw);
	int launches = 1;
	// the reads with very many chains: cut into units, the units through the same lane kernel, the few large units through a warp each
	unsigned int n = 0;
	cudaMemcpyAsync(&n, b.ext_heavy_n, 4, cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (!n) return launches;
	ext_plan_kernel<<<grid_for(n, LC_TPB, 8), LC_TPB, 0, stream>>>(b, n); ++launches;
	unsigned int cnt[4] = { 0, 0, 0, 0 };      // ext_unit_n[0..1], ext_hvu_n[0..1] are adjacent control words
	unsigned int n_back = 0;
	cudaMemcpyAsync(cnt, b.ext_unit_n, 16, cudaMemcpyDeviceToHost, stream);
	cudaMemcpyAsync(&n_back, b.ext_unit_back, 4, cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (cnt[0] + n_back) {
		if (cells8) lc_launch_extend<LcCells8<32>, true>(b, (long long)cnt[0] + n_back, n_sm, stream, hk, rw);
		else lc_launch_extend<LcCells16<32>, true>(b, (long long)cnt[0] + n_back, n_sm, stream, hk, rw);
		++launches;
	}
	unsigned int nh = cnt[2] < b.ext_hvu_cap ? cnt[2] : (unsigned int)b.ext_hvu_cap;
	if (nh) {
		const uint64_t *list = b.ext_hvu_in;      // heaviest first; the warps pull them from a counter
		if (nh > 1) {
			size_t need = 0;
			cub::DeviceRadixSort::SortKeysDescending(nullptr, need, b.ext_hvu_in, b.ext_hvu_out, (int)nh, 32, 48, stream);
			if (need > scan_tmp_bytes) { if (scan_tmp) cudaFree(scan_tmp); cudaMalloc(&scan_tmp, need + 256); scan_tmp_bytes = need + 256; }
			cub::DeviceRadixSort::SortKeysDescending(scan_tmp, need, b.ext_hvu_in, b.ext_hvu_out, (int)nh, 32, 48, stream);
			list = b.ext_hvu_out;
		}
		const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(b.lmax);
		extend_heavy_kernel<<<grid_for((unsigned long long)nh * 32, LC_TPB, 4), LC_TPB, smem, stream>>>(b, list, nh); ++launches;
	}
	return launches;
}
// Length of a device work list (one small D2H + stream sync), and the list sorted by weight, heaviest first (LPT order): the
// kernels that give a whole warp to an item finish when their slowest item does, so the slow ones must not start last.
const uint64_t *LcCudaBackend::sort_worklist(const LcBatch &b, const unsigned int *n_dev, unsigned int &n)
{
	n = 0;
	cudaMemcpyAsync(&n, n_dev, 4, cudaMemcpyDeviceToHost, stream);
	cudaStreamSynchronize(stream);
	if (n < 2) return b.wl_in;
	size_t need = 0;
	cub::DeviceRadixSort::SortKeysDescending(nullptr, need, b.wl_in, b.wl_out, (int)n, 32, 48, stream);
	if (need > scan_tmp_bytes) { if (scan_tmp) cudaFree(scan_tmp); cudaMalloc(&scan_tmp, need + 256); scan_tmp_bytes = need + 256; }
	cub::DeviceRadixSort::SortKeysDescending(scan_tmp, need, b.wl_in, b.wl_out, (int)n, 32, 48, stream);
	return b.wl_out;
}
void LcCudaBackend::k_pair(const LcBatch &b, int n_pairs)
{
	// at most n_workers threads: the patch scratch (w_int) is indexed by global thread id
	pair_prepare_kernel<<<grid_for((unsigned long long)n_pairs, LC_TPB, 4), LC_TPB, 0, stream>>>(b, n_pairs);
	// the listed pairs (count known on the device only: a resident-size grid, warps pull from a counter)
	if (b.prep_heavy_n) pair_prepare_heavy_kernel<<<grid_for((unsigned long long)n_pairs * 32, LC_TPB, 4), LC_TPB, 0, stream>>>(b);
	// at most n_workers/32 warps: the sub-optimal-row lists live in the per-worker scratch, indexed by global warp id.
	unsigned int n = 0;
	const uint64_t *list = sort_worklist(b, b.resc_n, n);      // pairs that need a Smith-Waterman, heaviest first
	if (n) {
			const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(b.lmax, 3);
		matesw_kernel<<<grid_for((unsigned long long)n * 32, LC_TPB, 4), LC_TPB, smem, stream>>>(b, list, n);
	}
}
void LcCudaBackend::k_cigar(const LcBatch &b, const uint64_t *list, unsigned long long n, const unsigned int *n_dev)
{
	const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(b.lmax);
	cudaMemsetAsync(b.work_ctr + 2, 0, 4, stream);      // the stage may be re-launched after a pool regrow
	global_cigar_kernel<<<grid_for(n * 32, LC_TPB, 4), LC_TPB, smem, stream>>>(b, list, n, n_dev);
}
void LcCudaBackend::k_cigar_plan(const LcBatch &b, unsigned long long n) { cigar_plan_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(b, n); }
int LcCudaBackend::k_cigar_band(const LcBatch &b, unsigned long long n_max)
{
	// list lengths are only known on the device: grids are sized for the worst case (every candidate in one list) and stride over the real length
	// The classes are independent (separate job lists, atomically carved output): they run side by side on four streams so
	// that the thin tail of one class (a few long jobs, one thread each) overlaps the bulk of the others.
	const int grid = grid_for(n_max, 128, 8);
	cudaEventRecord(ev_fork, stream);
	for (int c = 0; c < LC_CIG_NCLASS; ++c) cudaStreamWaitEvent(aux[c], ev_fork, 0);
	cigar_band_kernel<80><<<grid, 128, 0, aux[4]>>>(b, 4);
	cigar_band_kernel<64><<<grid, 128, 0, aux[3]>>>(b, 3);
	cigar_band_kernel<48><<<grid, 128, 0, aux[2]>>>(b, 2);
	cigar_band_kernel<32><<<grid, 128, 0, aux[1]>>>(b, 1);
	cigar_band_kernel<24><<<grid, 128, 0, aux[0]>>>(b, 0);
	for (int c = 0; c < LC_CIG_NCLASS; ++c) { cudaEventRecord(ev_join[c], aux[c]); cudaStreamWaitEvent(stream, ev_join[c], 0); }
	return LC_CIG_NCLASS;
}
void LcCudaBackend::k_cigar_fallback(const LcBatch &b, unsigned long long n_max)
{
	// sized for a small list (the usual case); the warps pull candidates from a counter, so a long list only takes longer
	k_cigar(b, b.cig_jobs + (size_t)LC_CIG_FALLBACK * b.cig_jobs_cap, n_max < 65536 ? n_max : 65536, b.cig_n + LC_CIG_FALLBACK);
}
void LcCudaBackend::k_alnx(const LcBatch &b, unsigned long long n) { alnx_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(b, n); }
void LcCudaBackend::k_pack(const LcBatch &b, const LcPackArgs &pa) { pack_kernel<<<grid_for((unsigned long long)b.n_reads + 1, 256, 8), 256, 0, stream>>>(b, pa); }
void LcCudaBackend::k_rebase(int64_t *off, int n, int64_t off0) { rebase_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(off, n, off0); }
void LcCudaBackend::scan_exclusive(const int32_t *in, int64_t *out, int n)
{
	cub::TransformInputIterator<int64_t, LcCastI64, const int32_t *> it(in, LcCastI64());
	size_t need = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, need, it, out, n, stream);
	if (need > scan_tmp_bytes) { if (scan_tmp) { cudaStreamSynchronize(stream); cudaFree(scan_tmp); } cudaMalloc(&scan_tmp, need + 256); scan_tmp_bytes = need + 256; }
	cub::DeviceScan::ExclusiveSum(scan_tmp, need, it, out, n, stream);
}
void LcCudaBackend::k_ksw_extend(const LcExtJobs &j)
{
	const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(j.qmax);
	ksw_extend_kernel<<<grid_for((unsigned long long)j.n * 32, LC_TPB, 8), LC_TPB, smem, stream>>>(j);
}
void LcCudaBackend::k_ksw_align(const LcSwJobs &j)
{
	const int smem = (LC_TPB / 32) * lc_warp_smem_bytes(j.qmax, 3);
	ksw_align_kernel<<<grid_for((unsigned long long)j.n * 32, LC_TPB, 4), LC_TPB, smem, stream>>>(j);
}
void LcCudaBackend::k_sa_batch(const LcIndex &ix, const uint64_t *rows, uint64_t *out, long long n, LcCounters *cnt)
{
	sa_batch_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(ix, rows, out, n, cnt);
}
int LcCudaBackend::check(std::string &err)
{
	cudaError_t e = cudaStreamSynchronize(stream);
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e == cudaSuccess && launch_err != cudaSuccess) { e = launch_err; launch_err = cudaSuccess; }
	if (e != cudaSuccess) { err = std::string("CUDA error: ") + cudaGetErrorString(e); return -1; }
	return 0;
}
