// oracle/ksort.inl -- TEST INFRASTRUCTURE. klib ksort.h's ks_introsort / ks_combsort restated
// (SURVEY.md A.13).  The instability pattern of exactly this procedure decides the order of
// equal-weight chains and equal-key regions, so it is reproduced step by step.
#pragma once
#include <cstddef>
#include <utility>
#include <vector>

namespace orc {

template <class T, class LT> static inline void ks_insertsort(T* s, T* t, LT lt) {
    for (T* i = s + 1; i < t; ++i)
        for (T* j = i; j > s && lt(*j, *(j - 1)); --j) std::swap(*j, *(j - 1));
}

template <class T, class LT> static inline void ks_combsort(size_t n, T* a, LT lt) {
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    size_t gap = n;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink_factor);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (T* i = a; i < a + n - gap; ++i) {
            T* j = i + gap;
            if (lt(*j, *i)) { std::swap(*i, *j); do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) ks_insertsort(a, a + n, lt);
}

template <class T, class LT> void ks_introsort(size_t n, T* a, LT lt) {
    struct Frame { T *left, *right; int depth; };
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) std::swap(a[0], a[1]); return; }
    int d;
    for (d = 2; (1ul << d) < n; ++d) {}
    std::vector<Frame> stack(sizeof(size_t) * d + 2);
    Frame* top = stack.data();
    T *s = a, *t = a + (n - 1), *i, *j, *k;
    d <<= 1;
    while (true) {
        if (s < t) {
            if (--d == 0) { ks_combsort((size_t)(t - s + 1), s, lt); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
            else k = lt(*j, *i) ? i : j;
            T rp = *k;
            if (k != t) std::swap(*k, *t);
            for (;;) {
                do ++i; while (lt(*i, rp));
                do --j; while (i <= j && lt(rp, *j));
                if (j <= i) break;
                std::swap(*i, *j);
            }
            std::swap(*i, *t);
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack.data()) { ks_insertsort(a, a + n, lt); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

}  // namespace orc
