// ksort_dev.cuh -- device-side klib ks_introsort.  bwa sorts chains and regions with
// this unstable procedure, so the order of equal keys is part of the result; the device version works
// on an array through indices and reproduces the same sequence of comparisons and swaps.
#pragma once
#include "common.cuh"

template <class T, class LT> __device__ inline void ks_insertsort_dev(T* a, int s, int t, LT lt) {
    for (int i = s + 1; i < t; ++i)
        for (int j = i; j > s && lt(a[j], a[j - 1]); --j) { T tmp = a[j]; a[j] = a[j - 1]; a[j - 1] = tmp; }
}

template <class T, class LT> __device__ inline void ks_combsort_dev(int n, T* a, LT lt) {
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    int gap = n;
    do {
        if (gap > 2) {
            gap = (int)(gap / shrink_factor);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (int i = 0; i < n - gap; ++i) {
            int j = i + gap;
            if (lt(a[j], a[i])) { T tmp = a[i]; a[i] = a[j]; a[j] = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) ks_insertsort_dev(a, 0, n, lt);
}

template <class T, class LT> __device__ inline void ks_introsort_dev(int n, T* a, LT lt) {
    struct Frame { int left, right, depth; };
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) { T tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
    int d;
    for (d = 2; (1u << d) < (unsigned)n; ++d) {}
    Frame stack[66];
    int top = 0;
    int s = 0, t = n - 1, i, j, k;
    d <<= 1;
    while (true) {
        if (s < t) {
            if (--d == 0) { ks_combsort_dev(t - s + 1, a + s, lt); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (lt(a[k], a[i])) { if (lt(a[k], a[j])) k = j; }
            else k = lt(a[j], a[i]) ? i : j;
            T rp = a[k];
            if (k != t) { T tmp = a[k]; a[k] = a[t]; a[t] = tmp; }
            for (;;) {
                do ++i; while (lt(a[i], rp));
                do --j; while (i <= j && lt(rp, a[j]));
                if (j <= i) break;
                T tmp = a[i]; a[i] = a[j]; a[j] = tmp;
            }
            { T tmp = a[i]; a[i] = a[t]; a[t] = tmp; }
            if (i - s > t - i) {
                if (i - s > 16) { stack[top].left = s; stack[top].right = i - 1; stack[top].depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { stack[top].left = i + 1; stack[top].right = t; stack[top].depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) { ks_insertsort_dev(a, 0, n, lt); return; }
            --top; s = stack[top].left; t = stack[top].right; d = stack[top].depth;
        }
    }
}
