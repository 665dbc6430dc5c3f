/* lc_percall_driver.c — libcuda_lariat.so driven the way Lariat's Go loop drives BWA: T worker threads, each handing the library
 * ONE barcode group (~100 read pairs) per call and keeping the result until it is done with that barcode, while the other workers
 * keep calling (go/src/inference/lariat.go:348-376 worker pool, :461-547 DoRFAForOneBarcode, :546 arena.Free; go/src/gobwa/gobwa.go:
 * 107-126 Arena).  Plain C over the C ABI of include/cuda_lariat.h — no Python, no torch: this is what the cgo shim of INTEGRATION.md
 * does, minus Go.  Used by tests/ (results held across calls must stay intact; output must equal the oracle's) and by
 * `bench.py --mode percall` (pairs/s and call latency in the drop-in regime).
 *
 * usage: lc_percall_driver <index-prefix> <pairs.bin> <out.bin|-> [-t threads] [-c pairs_per_call] [-n max_pairs] [-r repeat]
 *                          [-g device] [-G n_devices] [-k clones] [-h hold]
 *   -G n      : ONE process driving n GPUs (devices 0..n-1), the way a single Lariat process would: one lc_init per device (the index
 *               is replicated per GPU), worker t uses device t % n.  No process-wide state is shared between the devices' contexts.
 *   -k clones : number of contexts (lc_clone of the first) the threads are spread over; the index is in HBM once either way
 *   -h hold   : results a worker keeps alive before freeing the oldest (default 1: the previous barcode's result is still
 *               alive while the next call runs).  With an output file every result is kept until the dump is written.
 * Prints one JSON line on stderr. */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "../../include/cuda_lariat.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int g_threads = 32, g_call = 100, g_repeat = 1, g_hold = 1, g_clones = 1, g_ndev = 1, g_n, g_groups, g_keep;
static const uint8_t *g_bases; static const int64_t *g_off;
static lc_ctx **g_ctx;
static lc_result_t **g_res;          /* per group (kept when an output file is wanted) */
static uint64_t *g_sum;              /* per group: checksum taken right after the call ... */
static float *g_lat;                 /* per call latency (ms) */
static double g_dev[8];              /* sums over calls of the device-side stage times of the (merged) batch each call rode in */
static volatile int g_next, g_fail;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;

static uint64_t checksum(const lc_result_t *r)
{
	uint64_t h = 1469598103934665603ull;
	const uint8_t *p = (const uint8_t *)r->regs;
	for (int64_t i = 0; i < r->n_regs * (int64_t)sizeof(lc_alnreg_t); ++i) h = (h ^ p[i]) * 1099511628211ull;
	for (int64_t i = 0; i < r->n_regs; ++i) {
		const lc_aln_t *a = &r->aln[i];
		h = (h ^ (uint64_t)a->pos) * 1099511628211ull; h = (h ^ (uint64_t)a->NM) * 1099511628211ull;
		for (int k = 0; k < a->n_cigar; ++k) h = (h ^ r->cigars[a->cigar_off + k]) * 1099511628211ull;
	}
	for (int i = 0; i <= r->n_reads; ++i) h = (h ^ (uint64_t)r->reg_off[i]) * 1099511628211ull;
	return h;
}

static void *worker(void *arg)
{
	const int tid = (int)(intptr_t)arg;
	lc_ctx *ctx = g_ctx[(tid % g_ndev) * g_clones + (tid / g_ndev) % g_clones];
	lc_result_t **held = (lc_result_t **)calloc((size_t)g_hold + 1, sizeof(*held));
	int *held_g = (int *)calloc((size_t)g_hold + 1, sizeof(int));
	int nh = 0;
	for (;;) {
		int g;
		pthread_mutex_lock(&g_mu); g = g_next++; pthread_mutex_unlock(&g_mu);
		if (g >= g_groups * g_repeat || g_fail) break;
		const int gi = g % g_groups, p0 = gi * g_call, np = p0 + g_call <= g_n ? g_call : g_n - p0;
		lc_result_t *res = 0;
		const double t0 = now_s();
		const int rc = lc_align_packed(ctx, np, g_bases, g_off + 2 * (size_t)p0, 25, &res);
		g_lat[g] = (float)(1e3 * (now_s() - t0));
		if (rc) { fprintf(stderr, "lc_align_packed failed (%d): %s\n", rc, lc_last_error()); g_fail = 1; break; }
		if (res->n_reads != 2 * np) { fprintf(stderr, "wrong n_reads\n"); g_fail = 1; break; }
		pthread_mutex_lock(&g_mu);
		g_dev[0] += res->ms_total; g_dev[1] += res->ms_seed; g_dev[2] += res->ms_chain; g_dev[3] += res->ms_extend; g_dev[4] += res->ms_pair;
		g_dev[5] += res->ms_cigar; g_dev[6] += res->ms_h2d; g_dev[7] += res->ms_d2h;
		pthread_mutex_unlock(&g_mu);
		if (g_keep) { g_res[gi] = res; g_sum[gi] = checksum(res); continue; }
		/* keep it while the next calls run; before freeing the oldest, check that nothing has touched it in the meantime */
		if (g < g_groups) g_sum[gi] = checksum(res);
		if (nh == g_hold) {
			if (held_g[0] < g_groups && checksum(held[0]) != g_sum[held_g[0] % g_groups]) { fprintf(stderr, "a held result changed\n"); g_fail = 1; }
			lc_free_result(ctx, held[0]);
			memmove(held, held + 1, sizeof(*held) * (size_t)(nh - 1)); memmove(held_g, held_g + 1, sizeof(int) * (size_t)(nh - 1));
			--nh;
		}
		held[nh] = res; held_g[nh] = g; ++nh;
	}
	for (int i = 0; i < nh; ++i) lc_free_result(ctx, held[i]);
	free(held); free(held_g);
	return 0;
}

static int cmp_f(const void *a, const void *b) { const float x = *(const float *)a, y = *(const float *)b; return x < y ? -1 : x > y; }

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s <idx-prefix> <pairs.bin> <out.bin|-> [-t threads] [-c pairs_per_call] [-n max_pairs] [-r repeat] [-g device] [-k clones] [-h hold]\n", argv[0]); return 2; }
	int maxp = -1, device = 0;
	for (int i = 4; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "-t")) g_threads = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-c")) g_call = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-n")) maxp = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-r")) g_repeat = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-g")) device = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-k")) g_clones = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-G")) g_ndev = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "-h")) g_hold = atoi(argv[i + 1]);
	}
	if (g_threads < 1 || g_call < 1 || g_repeat < 1 || g_clones < 1 || g_hold < 0 || g_ndev < 1) return 2;
	FILE *f = fopen(argv[2], "rb");
	if (!f) { perror(argv[2]); return 1; }
	char magic[4]; int32_t n;
	if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "LRP1", 4) || fread(&n, 4, 1, f) != 1) { fprintf(stderr, "bad pairs file\n"); return 1; }
	int32_t *len = (int32_t *)malloc(8 * (size_t)n);
	if (fread(len, 4, 2 * (size_t)n, f) != 2 * (size_t)n) return 1;
	if (maxp >= 0 && maxp < n) n = maxp;
	int64_t *off = (int64_t *)malloc(8 * (2 * (size_t)n + 1));
	off[0] = 0;
	for (int i = 0; i < 2 * n; ++i) off[i + 1] = off[i] + len[i];
	uint8_t *bases = (uint8_t *)malloc((size_t)off[2 * n] + 16);      /* ordinary (pageable) memory, like a Go slice */
	if (fread(bases, 1, (size_t)off[2 * n], f) != (size_t)off[2 * n]) return 1;
	fclose(f);
	g_n = n; g_bases = bases; g_off = off; g_groups = (n + g_call - 1) / g_call;
	g_keep = strcmp(argv[3], "-") != 0;
	if (g_keep) g_repeat = 1;
	g_ctx = (lc_ctx **)calloc((size_t)g_clones * g_ndev, sizeof(*g_ctx));
	for (int d = 0; d < g_ndev; ++d) {
		lc_ctx **c = g_ctx + (size_t)d * g_clones;
		c[0] = lc_init(argv[1], g_ndev > 1 ? d : device, 0);
		if (!c[0]) { fprintf(stderr, "lc_init failed on device %d: %s\n", d, lc_last_error()); return 1; }
		for (int k = 1; k < g_clones; ++k) if (!(c[k] = lc_clone(c[0]))) { fprintf(stderr, "lc_clone failed: %s\n", lc_last_error()); return 1; }
	}
	g_res = (lc_result_t **)calloc((size_t)g_groups, sizeof(*g_res));
	g_sum = (uint64_t *)calloc((size_t)g_groups, 8);
	g_lat = (float *)calloc((size_t)g_groups * g_repeat, 4);
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)g_threads);
	{   /* warm-up: one pass over a few groups per thread (pool growth, lane creation), not timed */
		const int save_groups = g_groups, save_rep = g_repeat, save_keep = g_keep;
		g_groups = g_groups < 4 * g_threads ? g_groups : 4 * g_threads; g_repeat = 1; g_keep = 0; g_next = 0;
		for (int t = 0; t < g_threads; ++t) pthread_create(&th[t], 0, worker, (void *)(intptr_t)t);
		for (int t = 0; t < g_threads; ++t) pthread_join(th[t], 0);
		g_groups = save_groups; g_repeat = save_rep; g_keep = save_keep; g_next = 0;
		memset(g_dev, 0, sizeof(g_dev));
	}
	const double t0 = now_s();
	for (int t = 0; t < g_threads; ++t) pthread_create(&th[t], 0, worker, (void *)(intptr_t)t);
	for (int t = 0; t < g_threads; ++t) pthread_join(th[t], 0);
	const double dt = now_s() - t0;
	if (g_fail) return 1;
	if (g_keep) {      /* every result was held to the end: check none changed, then dump in barcode order (LRO1, lariat_b200/iofmt.py) */
		FILE *o = fopen(argv[3], "wb");
		if (!o) { perror(argv[3]); return 1; }
		fwrite("LRO1", 1, 4, o); fwrite(&n, 4, 1, o);
		for (int g = 0; g < g_groups; ++g) {
			const lc_result_t *r = g_res[g];
			if (checksum(r) != g_sum[g]) { fprintf(stderr, "result of group %d changed while it was held\n", g); return 1; }
			for (int rd = 0; rd < r->n_reads; ++rd) {
				const int32_t k = (int32_t)(r->reg_off[rd + 1] - r->reg_off[rd]);
				fwrite(&k, 4, 1, o);
				if (k) fwrite(&r->regs[r->reg_off[rd]], sizeof(lc_alnreg_t), (size_t)k, o);
				for (int i = 0; i < k; ++i) {
					const lc_aln_t *a = &r->aln[r->reg_off[rd] + i];
					int64_t pos = a->pos; int32_t h[4] = { a->rid, a->is_rev, a->NM, a->n_cigar };
					fwrite(&pos, 8, 1, o); fwrite(h, 4, 4, o);
					if (a->n_cigar) fwrite(&r->cigars[a->cigar_off], 4, (size_t)a->n_cigar, o);
				}
			}
		}
		fclose(o);
		for (int g = 0; g < g_groups; ++g) lc_free_result(g_ctx[0], g_res[g]);      /* (a result may be freed through any context) */
	}
	const int calls = g_groups * g_repeat;
	qsort(g_lat, (size_t)calls, 4, cmp_f);
	fprintf(stderr, "{\"pairs\": %d, \"calls\": %d, \"threads\": %d, \"pairs_per_call\": %d, \"clones\": %d, \"devices\": %d, \"hold\": %d, \"seconds\": %.4f, \"pairs_per_s\": %.1f, "
			"\"call_ms_p50\": %.3f, \"call_ms_p90\": %.3f, \"call_ms_p99\": %.3f, \"call_ms_max\": %.3f, "
			"\"batch_device_ms_mean\": {\"total\": %.3f, \"seed\": %.3f, \"chain\": %.3f, \"extend\": %.3f, \"pair\": %.3f, \"cigar\": %.3f, \"h2d\": %.3f, \"d2h\": %.3f}}\n",
			n, calls, g_threads, g_call, g_clones, g_ndev, g_hold, dt,
			(double)n * g_repeat / dt, g_lat[calls / 2], g_lat[(int)(calls * 0.9)], g_lat[(int)(calls * 0.99)], g_lat[calls - 1],
			g_dev[0] / calls, g_dev[1] / calls, g_dev[2] / calls, g_dev[3] / calls, g_dev[4] / calls, g_dev[5] / calls, g_dev[6] / calls, g_dev[7] / calls);
	for (int k = g_clones * g_ndev - 1; k >= 0; --k) lc_destroy(g_ctx[k]);
	return 0;
}
