// ibsim.cpp — TEST-ONLY host execution of the index builder's logic (lariat_b200/csrc/lc_index_build.hpp).
// The product runs the same element-wise bodies as CUDA kernels with cub sorts/scans (lc_index_build.cu); this harness swaps in
// host loops, std::stable_sort and serial scans so that the chunk plan, the refinement rounds, the end-of-text rule and the five
// file formats can be compared byte for byte with `bwa index` output on a machine without a GPU.  Never linked into the product.
// usage: ibsim <in.fa> <out-prefix> [chunk]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <numeric>
#include <string>
#include <vector>
#include "../../lariat_b200/csrc/lc_index_build.hpp"

struct HostIbBackend {
	void *dmalloc(size_t n) { return malloc(n ? n : 8); }
	void dfree(void *p) { free(p); }
	void dzero(void *p, size_t n) { memset(p, 0, n); }
	void h2d(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void d2h(void *d, const void *s, size_t n) { memcpy(d, s, n); }
	void sync() {}
	template <class F> void launch(size_t n, const F &f) { for (size_t i = n; i-- > 0;) f(i); }      // reversed: the order of the "threads" must not matter
	void sort_pairs_u64(const uint64_t *kin, uint64_t *kout, const uint64_t *vin, uint64_t *vout, size_t n, int end_bit)
	{
		std::vector<size_t> idx(n);
		std::iota(idx.begin(), idx.end(), (size_t)0);
		const uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
		std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return (kin[a] & mask) < (kin[b] & mask); });
		for (size_t i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
	}
	void scan_max_u32(const uint32_t *in, uint32_t *out, size_t n) { uint32_t m = 0; for (size_t i = 0; i < n; ++i) { m = in[i] > m ? in[i] : m; out[i] = m; } }
	void scan_sum_u32(const uint32_t *in, uint32_t *out, size_t n) { uint32_t s = 0; for (size_t i = 0; i < n; ++i) { const uint32_t v = in[i]; out[i] = s; s += v; } }
	void scan_c4(const LcIbC4 *in, LcIbC4 *out, size_t n)
	{
		LcIbC4 s; s.c[0] = s.c[1] = s.c[2] = s.c[3] = 0;
		for (size_t i = 0; i < n; ++i) { const LcIbC4 v = in[i]; out[i] = s; s = LcIbC4Add()(s, v); }
	}
	bool ok(std::string &) { return true; }
};

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: ibsim <in.fa> <out-prefix> [chunk]\n"); return 2; }
	FILE *f = fopen(argv[1], "r");
	if (!f) { perror(argv[1]); return 1; }
	std::vector<std::string> names;
	std::vector<std::vector<uint8_t> > seqs;
	char line[1 << 16];
	while (fgets(line, sizeof line, f)) {
		if (line[0] == '>') {
			char *e = line + 1;
			while (*e && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r') ++e;
			names.push_back(std::string(line + 1, e));
			seqs.emplace_back();
		} else {
			for (char *p = line; *p && *p != '\n' && *p != '\r'; ++p) {
				int c = *p == 'A' || *p == 'a' ? 0 : *p == 'C' || *p == 'c' ? 1 : *p == 'G' || *p == 'g' ? 2 : *p == 'T' || *p == 't' ? 3 : -1;
				if (c < 0) { fprintf(stderr, "ambiguous base in %s\n", argv[1]); return 1; }
				seqs.back().push_back((uint8_t)c);
			}
		}
	}
	fclose(f);
	std::vector<const char *> np; std::vector<const uint8_t *> cp; std::vector<int64_t> lens;
	for (size_t i = 0; i < names.size(); ++i) { np.push_back(names[i].c_str()); cp.push_back(seqs[i].data()); lens.push_back((int64_t)seqs[i].size()); }
	HostIbBackend be;
	std::string err;
	LcIbStats st;
	const int rc = lc_ib_build(be, argv[2], (int)names.size(), np.data(), cp.data(), lens.data(), argc > 3 ? strtoull(argv[3], 0, 10) : 0, err, &st);
	if (rc) { fprintf(stderr, "ibsim: error %d: %s\n", rc, err.c_str()); return 1; }
	fprintf(stderr, "ibsim: %llu symbols, %d chunks (largest %llu), at most %d rounds, %.2f x n sorted\n", (unsigned long long)st.n, st.n_chunks,
			(unsigned long long)st.max_chunk, st.max_rounds, (double)st.sorted_elems / (double)st.n);
	return 0;
}
