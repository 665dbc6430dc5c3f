// lc_sort.cuh — the reference's UNSTABLE introsort, reproduced comparison for comparison.
//
// BWA sorts chains by weight, regions by end and by (score, rb, qb) with ks_introsort (B/ksort.h:176-226).
// Keys tie often in repeats and the permutation left among equal keys changes which chains survive
// mem_chain_flt and which regions survive mem_sort_dedup_patch, so the device code runs the very same
// algorithm: median-of-three quicksort leaving runs <= 16 for one final insertion sort (B/ksort.h:146-153),
// comb sort when the depth budget 2*ceil(log2 n) runs out (B/ksort.h:154-175).  It sorts an array of small
// handles (indices or 64-bit keys) through a caller-supplied strict-weak `lt`, which sees the same operand
// pairs in the same order as the reference's macro would.
#pragma once
#include "lc_fmindex.cuh"

template <typename T, typename Lt> LC_HD void lc_insertion_sort(T *a, long n, Lt lt)
{
	for (long i = 1; i < n; ++i)
		for (long j = i; j > 0 && lt(a[j], a[j - 1]); --j) { T t = a[j]; a[j] = a[j - 1]; a[j - 1] = t; }
}

template <typename T, typename Lt> LC_HD void lc_comb_sort(T *a, long n, Lt lt)
{
	const double shrink = 1.2473309501039786540366528676643;
	long gap = n;
	bool swapped;
	do {
		if (gap > 2) {
			gap = (long)((double)gap / shrink);
			if (gap == 9 || gap == 10) gap = 11;
		}
		swapped = false;
		for (long i = 0; i + gap < n; ++i)
			if (lt(a[i + gap], a[i])) { T t = a[i]; a[i] = a[i + gap]; a[i + gap] = t; swapped = true; }
	} while (swapped || gap > 2);
	if (gap != 1) lc_insertion_sort(a, n, lt);
}

template <typename T, typename Lt> LC_HDN void lc_introsort(T *a, long n, Lt lt)
{
	struct Frame { int lo, hi, depth; };
	Frame stack[64];
	int top = 0, d;
	if (n < 1) return;
	if (n == 2) { if (lt(a[1], a[0])) { T t = a[0]; a[0] = a[1]; a[1] = t; } return; }
	for (d = 2; (1l << d) < n; ++d) ;
	d <<= 1;
	long s = 0, e = n - 1;
	for (;;) {
		if (s < e) {
			if (--d == 0) { lc_comb_sort(a + s, e - s + 1, lt); e = s; continue; }
			long i = s, j = e, k = i + ((j - i) >> 1) + 1;
			if (lt(a[k], a[i])) { if (lt(a[k], a[j])) k = j; }
			else k = lt(a[j], a[i]) ? i : j;
			T pv = a[k];
			if (k != e) { a[k] = a[e]; a[e] = pv; }
			for (;;) {
				do ++i; while (lt(a[i], pv));
				do --j; while (i <= j && lt(pv, a[j]));
				if (j <= i) break;
				T t = a[i]; a[i] = a[j]; a[j] = t;
			}
			{ T t = a[i]; a[i] = a[e]; a[e] = t; }
			if (i - s > e - i) {
				if (i - s > 16) { stack[top].lo = (int)s; stack[top].hi = (int)(i - 1); stack[top].depth = d; ++top; }
				s = e - i > 16 ? i + 1 : e;
			} else {
				if (e - i > 16) { stack[top].lo = (int)(i + 1); stack[top].hi = (int)e; stack[top].depth = d; ++top; }
				e = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == 0) { lc_insertion_sort(a, n, lt); return; }
			--top; s = stack[top].lo; e = stack[top].hi; d = stack[top].depth;
		}
	}
}
