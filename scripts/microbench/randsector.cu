// randsector.cu — how many dependent random 32-byte sector reads per second does a B200 sustain, and what limits it?
// Every lane walks its own chain: two independent 32-byte loads per step (like the two occ blocks of one bwt_extend), the next
// addresses depend on the data.  Swept over table size (L2-resident .. DRAM), resident warps per SM, the shared-memory carve-out
// (= what is left for L1) and the load flavour.  Result (profiles/r01_randsector.txt): the rate is set by the L1 size, i.e. every
// in-flight sector miss holds an L1 line; with the carve-out at 100 % a 100 MB table gives 45 G sectors/s, at <= 50 % 150 G/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randsector randsector.cu ; run: ./randsector
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
enum { V_NC = 0, V_CG, V_CV, V_NOALLOC, V_CPASYNC, V_N };
static const char *vname[] = { "nc", "cg", "cv", "nc.noalloc", "cp.async.cg" };
template <int V>
__device__ __forceinline__ void ld256(const uint4 *p, uint64_t &a, uint64_t &b, uint64_t &c, uint64_t &d)
{
	if (V == V_NC) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
	else if (V == V_CG) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
	else if (V == V_CV) asm volatile("ld.global.cv.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
	else asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
template <int V>
__global__ void walk(const uint4 *tab, uint32_t n_blocks, int iters, uint64_t *out)
{
	extern __shared__ uint4 sm[];
	uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
	uint64_t acc = 0;
	for (int i = 0; i < iters; ++i) {
		uint64_t a[2], b[2], c[2], d[2];
		if (V == V_CPASYNC) {
			uint4 *slot = sm + threadIdx.x * 4;
#pragma unroll
			for (int t = 0; t < 2; ++t) {
				const uint32_t idx = (uint32_t)(((uint64_t)(s * (2 * t + 1) + t * 0x9e3779b9u) * n_blocks) >> 32);
				const uint4 *g = tab + 2 * (size_t)idx;
				const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot + 2 * t);
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(g));
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa + 16), "l"(g + 1));
			}
			asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
			for (int t = 0; t < 2; ++t) {
				const uint4 x = slot[2 * t], y = slot[2 * t + 1];
				a[t] = (uint64_t)x.y << 32 | x.x; b[t] = (uint64_t)x.w << 32 | x.z; c[t] = (uint64_t)y.y << 32 | y.x; d[t] = (uint64_t)y.w << 32 | y.z;
			}
		} else {
#pragma unroll
			for (int t = 0; t < 2; ++t) {
				const uint32_t idx = (uint32_t)(((uint64_t)(s * (2 * t + 1) + t * 0x9e3779b9u) * n_blocks) >> 32);
				ld256<V>(tab + 2 * (size_t)idx, a[t], b[t], c[t], d[t]);
			}
		}
#pragma unroll
		for (int t = 0; t < 2; ++t) { acc += a[t] ^ d[t]; s = s * 1664525u + 1013904223u + (uint32_t)(b[t] + c[t]); }
	}
	if (acc == 0x1234567ull) out[0] = acc;
}
typedef void (*kern_t)(const uint4 *, uint32_t, int, uint64_t *);
int main()
{
	int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
	const size_t maxb = (size_t)1024 << 20;
	uint4 *tab; cudaMalloc(&tab, maxb); cudaMemset(tab, 0x5a, maxb);
	uint64_t *out; cudaMalloc(&out, 8);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int sizes_mb[] = { 100, 1024 };
	const int warps[] = { 12, 16, 24, 32 };
	const int carve[] = { 25, 50, 62, 75, 88, 100 };
	kern_t ks[V_N] = { walk<V_NC>, walk<V_CG>, walk<V_CV>, walk<V_NOALLOC>, walk<V_CPASYNC> };
	printf("%-12s table_MB warps/SM carveout%% smem/CTA  Gsectors/s  ns_per_step\n", "load");
	for (int v = 0; v < V_N; ++v) for (int si = 0; si < 2; ++si) for (int ci = 0; ci < 6; ++ci) for (int wi = 0; wi < 4; ++wi) {
		kern_t k = ks[v];
		const uint32_t nb = (uint32_t)(((size_t)sizes_mb[si] << 20) / 32);
		const int w = warps[wi], tpb = 64, ctas = w / 2;
		// the carve-out is really occupied with dynamic shared memory, as a kernel with per-lane state there would
		int smem = (int)((227 * 1024) * (carve[ci] / 100.0)) / ctas - 1024; smem &= ~127;
		if (smem < 4096) continue;
		cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve[ci]);
		int per = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k, tpb, smem);
		if (per < ctas) { printf("%-12s %6d %6d %6d %8d   occupancy %d CTAs only\n", vname[v], sizes_mb[si], w, carve[ci], smem, per); continue; }
		const int iters = 2000;
		k<<<nsm * ctas, tpb, smem>>>(tab, nb, 200, out);
		cudaEventRecord(e0);
		k<<<nsm * ctas, tpb, smem>>>(tab, nb, iters, out);
		cudaEventRecord(e1); cudaEventSynchronize(e1);
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		const double sect = (double)nsm * ctas * tpb * iters * 2;
		printf("%-12s %6d %6d %6d %8d %10.1f %10.1f\n", vname[v], sizes_mb[si], w, carve[ci], smem, sect / ms / 1e6, ms * 1e6 / iters);
	}
	return 0;
}
