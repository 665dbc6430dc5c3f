// lc_index_build.cu — CUDA backend of the index builder (lc_index_build.hpp) and its C-ABI entry point lc_index_build
// (replaces `bwa index` = bwa_idx_build, B/bwtindex.c:251-316, for N-free references).
// Sorting is cub::DeviceRadixSort (stable LSD radix sort) and the scans are cub::DeviceScan — library code for the generic
// primitives; everything specific to the BWT construction is the element-wise kernels of lc_index_build.hpp.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <chrono>
#include <string>
#include <stdlib.h>
#include "lc_index_build.hpp"
#include "../../include/cuda_lariat.h"

void lc_api_set_error(const std::string &msg);      // lc_api.cu: the text lc_last_error() returns

template <class F> __global__ void __launch_bounds__(256) lc_ib_for_kernel(const F f, const size_t n)
{
	for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) f(i);
}
struct LcIbMaxOp { __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t &a, const uint32_t &b) const { return a > b ? a : b; } };

struct LcIbCudaBackend {
	int n_sm = 148;
	cudaStream_t stream = 0;
	void *tmp = nullptr; size_t tmp_bytes = 0;
	cudaError_t first_err = cudaSuccess;
	double sec_sort = 0;
	explicit LcIbCudaBackend(int device)
	{
		cudaSetDevice(device);
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) n_sm = prop.multiProcessorCount;
		cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
	}
	~LcIbCudaBackend() { if (tmp) cudaFree(tmp); cudaStreamDestroy(stream); }
	void note(cudaError_t e) { if (e != cudaSuccess && first_err == cudaSuccess) first_err = e; }
	void *dmalloc(size_t n) { void *p = nullptr; if (cudaMalloc(&p, n ? n : 8) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
	void dfree(void *p) { cudaStreamSynchronize(stream); cudaFree(p); }
	void dzero(void *p, size_t n) { note(cudaMemsetAsync(p, 0, n, stream)); }
	void h2d(void *d, const void *s, size_t n) { note(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream)); }
	void d2h(void *d, const void *s, size_t n) { note(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream)); }
	void sync() { note(cudaStreamSynchronize(stream)); }
	template <class F> void launch(size_t n, const F &f)
	{
		if (!n) return;
		const size_t need = (n + 255) / 256, cap = (size_t)n_sm * 16;       // grid = a multiple of the SM count once the work is large enough
		lc_ib_for_kernel<F><<<(unsigned)(need < cap ? need : cap), 256, 0, stream>>>(f, n);
		note(cudaGetLastError());
	}
	bool ensure_tmp(size_t bytes)
	{
		if (bytes <= tmp_bytes) return true;
		if (tmp) { cudaStreamSynchronize(stream); cudaFree(tmp); tmp = nullptr; tmp_bytes = 0; }
		if (cudaMalloc(&tmp, bytes + 1024) != cudaSuccess) { note(cudaGetLastError()); return false; }
		tmp_bytes = bytes + 1024;
		return true;
	}
	void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint64_t *vin, uint64_t *vout, size_t n, int end_bit)
	{
		size_t need = 0;
		note(cub::DeviceRadixSort::SortPairs(nullptr, need, kin, kout, vin, vout, (int)n, 0, end_bit, stream));
		if (!ensure_tmp(need)) return;
		note(cub::DeviceRadixSort::SortPairs(tmp, need, kin, kout, vin, vout, (int)n, 0, end_bit, stream));
	}
	void scan_max_u32(const uint32_t *in, uint32_t *out, size_t n)
	{
		size_t need = 0;
		note(cub::DeviceScan::InclusiveScan(nullptr, need, in, out, LcIbMaxOp(), (int)n, stream));
		if (!ensure_tmp(need)) return;
		note(cub::DeviceScan::InclusiveScan(tmp, need, in, out, LcIbMaxOp(), (int)n, stream));
	}
	void scan_sum_u32(const uint32_t *in, uint32_t *out, size_t n)
	{
		size_t need = 0;
		note(cub::DeviceScan::ExclusiveSum(nullptr, need, in, out, (int)n, stream));
		if (!ensure_tmp(need)) return;
		note(cub::DeviceScan::ExclusiveSum(tmp, need, in, out, (int)n, stream));
	}
	void scan_c4(const LcIbC4 *in, LcIbC4 *out, size_t n)
	{
		size_t need = 0;
		LcIbC4 zero; zero.c[0] = zero.c[1] = zero.c[2] = zero.c[3] = 0;
		note(cub::DeviceScan::ExclusiveScan(nullptr, need, in, out, LcIbC4Add(), zero, (int)n, stream));
		if (!ensure_tmp(need)) return;
		note(cub::DeviceScan::ExclusiveScan(tmp, need, in, out, LcIbC4Add(), zero, (int)n, stream));
	}
	bool ok(std::string &err)
	{
		cudaError_t e = cudaStreamSynchronize(stream);
		if (e == cudaSuccess) e = cudaGetLastError();
		if (e == cudaSuccess) e = first_err;
		if (e != cudaSuccess) { err = std::string("CUDA error: ") + cudaGetErrorString(e); return false; }
		return true;
	}
};

extern "C" int lc_index_build(const char *prefix, int n_contigs, const char *const *names, const uint8_t *const *codes, const int64_t *lens, int device)
{
	if (!prefix || !names || !codes || !lens || n_contigs <= 0) { lc_api_set_error("bad arguments"); return LC_EARG; }
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); lc_api_set_error("no CUDA device: libcuda_lariat has no CPU path"); return LC_ENODEV; }
	if (device < 0 || device >= ndev) { lc_api_set_error("bad device ordinal"); return LC_ENODEV; }
	uint64_t chunk = 0;
	if (const char *e = getenv("LC_IB_CHUNK")) chunk = strtoull(e, nullptr, 10);      // tests: tiny chunks force the multi-chunk path on small genomes
	if (chunk >= 0x7fffffffull) { lc_api_set_error("LC_IB_CHUNK must be below 2^31"); return LC_EARG; }
	LcIbCudaBackend be(device);
	std::string err;
	LcIbStats st;
	const auto t0 = std::chrono::steady_clock::now();
	const int rc = lc_ib_build(be, prefix, n_contigs, names, codes, lens, chunk, err, &st);
	if (rc) { lc_api_set_error(err); return rc; }
	if (getenv("LC_VERBOSE")) {
		const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "[lc_index_build] %llu symbols, %d chunks (largest %llu), at most %d refinement rounds, %.2f x n elements radix-sorted, %.1f s\n",
				(unsigned long long)st.n, st.n_chunks, (unsigned long long)st.max_chunk, st.max_rounds, (double)st.sorted_elems / (double)st.n, s);
	}
	return LC_OK;
}
