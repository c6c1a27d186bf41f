// ksw_warp.cuh -- warp-cooperative banded DP kernels with libbwa ksw.c semantics.
//
// Both ksw_extend2 and ksw_global2 open gaps from M only, so within one target row M and E depend on the
// previous row alone and F is a max-plus prefix scan along the row.  A warp therefore sweeps the row in
// 32-column chunks: lanes compute M/E in parallel, F comes from a 5-step shuffle scan, H = max3(M,E,F)
// (DPX __vimax3_s32 / __viaddmax_s32*), and the per-row band trimming, z-drop and tie rules of the scalar
// code are applied between rows on warp-uniform registers -- bit-exact, including the right-hand trim
// that a static anti-diagonal wavefront cannot reproduce ("hard parts").
#pragma once
#include "common.cuh"

#define KSW_NEG_INF (-0x40000000)

struct ExtOut { int score, qle, tle, gtle, gscore, max_off; };

// ---- address-space helpers: the row sweeps run on warp-private scratch that lives either in shared memory (short reads)
// or in global memory (long reads).  With SH the 32-bit shared-window address is used directly (ld/st.shared, 32-bit
// address arithmetic); otherwise plain generic pointers.
__device__ __forceinline__ int2 sp_ld2(uint32_t a) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int2 sp_ld2(const uint8_t* a) { return *reinterpret_cast<const int2*>(a); }
__device__ __forceinline__ void sp_st2(uint32_t a, int x, int y) { asm volatile("st.shared.v2.s32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sp_st2(const uint8_t* a, int x, int y) { *reinterpret_cast<int2*>(const_cast<uint8_t*>(a)) = make_int2(x, y); }
__device__ __forceinline__ int sp_ldb(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return (int)v; }
__device__ __forceinline__ int sp_ldb(const uint8_t* a) { return (int)*a; }
template <bool SH> struct SpBase;
template <> struct SpBase<true> { typedef uint32_t type; static __device__ __forceinline__ type of(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); } };
template <> struct SpBase<false> { typedef const uint8_t* type; static __device__ __forceinline__ type of(const void* p) { return reinterpret_cast<const uint8_t*>(p); } };

// ksw_extend2.  q[j*qs], t[i*ts]; ehh: 2 (qlen + 1) ints of warp-private scratch, used as one {h, e} pair per column
// (both callers carve ehe right behind ehh); smat: 25 ints.  SH: q, t and ehh are in shared memory.
template <bool SH>
static __device__ __noinline__ ExtOut ksw_extend_warp_t(const DevOpts& o, int qlen, const uint8_t* q, int qs, int tlen, const uint8_t* t, int ts,
                                                 int w, int end_bonus, int h0, int* ehh, const int* smat,
                                                 unsigned long long& cells, unsigned long long& rows) {
    const int lane = threadIdx.x & 31;
    const int o_del = o.o_del, e_del = o.e_del, o_ins = o.o_ins, e_ins = o.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const typename SpBase<SH>::type eh = SpBase<SH>::of(ehh), qa = SpBase<SH>::of(q), ta = SpBase<SH>::of(t);
    // bwa_fill_scmat matrices have three distinct values (match, mismatch, ambiguous): scores by comparison, no table walk
    const int sA = smat[0], sB = smat[1], sN = smat[4];
    bool simple;
    { const int a = lane / 5, b = lane - a * 5; simple = __all_sync(FULL, lane >= 25 || smat[lane] == ((a == 4 || b == 4) ? sN : (a == b ? sA : sB))); }
    // first row (closed form of the scalar fill loop): eh[0].h = h0, eh[j].h = max(h0 - o_ins - j e_ins, 0)
    for (int j = lane; j <= qlen; j += 32) {
        int v = j == 0 ? h0 : h0 - o_ins - j * e_ins;
        sp_st2(eh + j * 8, v > 0 ? v : 0, 0);
    }
    {
        int max_ins = (int)((double)(qlen * o.mat_max + end_bonus - o_ins) / e_ins + 1.);
        max_ins = max_ins > 1 ? max_ins : 1;
        w = w < max_ins ? w : max_ins;
        int max_del = (int)((double)(qlen * o.mat_max + end_bonus - o_del) / e_del + 1.);
        max_del = max_del > 1 ? max_del : 1;
        w = w < max_del ? w : max_del;
    }
    int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    uint32_t ncell = 0, nrow = 0;
    __syncwarp();
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) { h1_init = h0 - (o_del + e_del * (i + 1)); if (h1_init < 0) h1_init = 0; }
        const int tb = sp_ldb(ta + i * ts);
        int carryU = (beg - 1) * e_ins, carryH = h1_init;
        int best_h = -1, best_j = -1, first_nz = -1, last_nz = -1;
        ncell += (uint32_t)(end > beg ? end - beg : 0); ++nrow;
        if (end - beg <= 32 && end > beg) {
            // the common case: the whole row is one chunk -- same arithmetic as the loop below with its carries folded in
            const int j = beg + lane;
            const bool act = j < end;
            int M = 0, e = 0;
            if (act) {
                const int2 v = sp_ld2(eh + j * 8);
                const int qb = sp_ldb(qa + j * qs);
                const int sc = simple ? ((qb | tb) > 3 ? sN : (qb == tb ? sA : sB)) : smat[tb * 5 + qb];
                e = v.y; M = v.x ? v.x + sc : 0;
            }
            int inc = act ? __viaddmax_s32(M, -oe_ins, 0) + j * e_ins : KSW_NEG_INF;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int ov = __shfl_up_sync(FULL, inc, d); inc = lane >= d ? ::max(inc, ov) : inc; }
            int ex = __shfl_up_sync(FULL, inc, 1);
            ex = lane == 0 ? carryU : ::max(ex, carryU);
            const int f = ex - (j - 1) * e_ins;
            const int h = __vimax3_s32(M, e, f);
            int hl = __shfl_up_sync(FULL, h, 1);
            if (lane == 0) hl = h1_init;
            const int e2 = __viaddmax_s32_relu(e, -e_del, M - oe_del);
            if (act) sp_st2(eh + j * 8, hl, e2);
            const unsigned bal = __ballot_sync(FULL, act && (hl | e2) != 0);
            if (bal) { first_nz = beg + __ffs(bal) - 1; last_nz = beg + 31 - __clz(bal); }
            carryH = __shfl_sync(FULL, h, end - beg - 1);
            best_h = act ? h : -1; best_j = act ? j : -1;
        } else
        for (int j0 = beg; j0 < end; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j < end;
            int M = 0, e = 0;
            if (act) {
                const int2 v = sp_ld2(eh + j * 8);
                const int qb = sp_ldb(qa + j * qs);
                const int sc = simple ? ((qb | tb) > 3 ? sN : (qb == tb ? sA : sB)) : smat[tb * 5 + qb];
                e = v.y; M = v.x ? v.x + sc : 0;
            }
            int u = act ? __viaddmax_s32(M, -oe_ins, 0) + j * e_ins : KSW_NEG_INF;
            int inc = u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int ov = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc = ::max(inc, ov); }
            int ex = __shfl_up_sync(FULL, inc, 1);
            ex = lane == 0 ? carryU : ::max(ex, carryU);
            const int f = ex - (j - 1) * e_ins;
            const int h = __vimax3_s32(M, e, f);
            int hl = __shfl_up_sync(FULL, h, 1);
            if (lane == 0) hl = carryH;
            bool nz = false;
            if (act) {
                const int e2 = __viaddmax_s32_relu(e, -e_del, M - oe_del);
                sp_st2(eh + j * 8, hl, e2);
                if (h >= best_h) { best_h = h; best_j = j; }
                nz = (hl | e2) != 0;
            }
            const unsigned bal = __ballot_sync(FULL, nz);
            if (bal) { if (first_nz < 0) first_nz = j0 + __ffs(bal) - 1; last_nz = j0 + 31 - __clz(bal); }
            const int last = (end - j0 < 32 ? end - j0 : 32) - 1;
            carryU = ::max(carryU, __shfl_sync(FULL, inc, 31));
            carryH = __shfl_sync(FULL, h, last);
        }
        const int h1 = carryH;  // H(i, end-1), or the first-column value when the row is empty
        if (lane == 0) sp_st2(eh + end * 8, h1, 0);
        __syncwarp();
        const int jfin = beg < end ? end : beg;
        if (jfin == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        int m = __reduce_max_sync(FULL, best_h);
        int mj = __reduce_max_sync(FULL, best_h == m ? best_j : -1);
        if (m < 0) { m = 0; mj = -1; }
        if (m == 0) break;
        if (m > max) {
            max = m; max_i = i; max_j = mj;
            int off = mj - i; off = off < 0 ? -off : off;
            max_off = max_off > off ? max_off : off;
        } else if (o.zdrop > 0) {
            if (i - max_i > mj - max_j) {
                if (max - m - ((i - max_i) - (mj - max_j)) * e_del > o.zdrop) break;
            } else {
                if (max - m - ((mj - max_j) - (i - max_i)) * e_ins > o.zdrop) break;
            }
        }
        // band trimming for the next row (literal: eh[end] takes part in the downward scan only)
        int nbeg = first_nz >= 0 ? first_nz : end;
        int jj;
        if (h1 != 0) jj = end;
        else if (last_nz >= 0) jj = last_nz;
        else jj = nbeg - 1;
        beg = nbeg;
        end = jj + 2 < qlen ? jj + 2 : qlen;
    }
    cells += ncell; rows += nrow;
    ExtOut r;
    r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
    return r;
}

// ksw_global2.  z: traceback bytes, n_col per row (nullptr => score only).
static __device__ __noinline__ int ksw_global_warp(const DevOpts& o, int qlen, const uint8_t* q, int qs, int tlen, const uint8_t* t, int ts,
                                            int w, int* ehh, int* ehe, const int* smat, uint8_t* z, int n_col, unsigned long long& cells) {
    const int lane = lane_id();
    const int o_del = o.o_del, e_del = o.e_del, o_ins = o.o_ins, e_ins = o.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    for (int j = lane; j <= qlen; j += 32) {
        int hv = j == 0 ? 0 : (j <= w ? -(o_ins + e_ins * j) : KSW_NEG_INF);
        ehh[j] = hv; ehe[j] = KSW_NEG_INF;
    }
    __syncwarp();
    for (int i = 0; i < tlen; ++i) {
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        const int h1_init = beg == 0 ? -(o_del + e_del * (i + 1)) : KSW_NEG_INF;
        const int* mrow = smat + (int)t[(long)i * ts] * 5;
        uint8_t* zi = z ? z + (size_t)i * n_col : nullptr;
        // F(i, beg) = -inf: carry in "u-space" so that f_j = U_{j-1} - (j-1) e_ins reproduces it exactly
        int carryU = KSW_NEG_INF + (beg - 1) * e_ins, carryH = h1_init;
        cells += (unsigned long long)(end > beg ? end - beg : 0);
        for (int j0 = beg; j0 < end; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j < end;
            int m = KSW_NEG_INF, e = KSW_NEG_INF;
            if (act) { m = ehh[j] + mrow[q[(long)j * qs]]; e = ehe[j]; }
            int u = act ? (m - oe_ins) + j * e_ins : 2 * KSW_NEG_INF + 1;
            int inc = u;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int ov = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc = ::max(inc, ov); }
            int ex = __shfl_up_sync(FULL, inc, 1);
            ex = lane == 0 ? carryU : ::max(ex, carryU);
            const int f = ex - (j - 1) * e_ins;        // F(i, j)
            int d = m >= e ? 0 : 1;
            int h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            int hl = __shfl_up_sync(FULL, h, 1);
            if (lane == 0) hl = carryH;
            if (act) {
                ehh[j] = hl;
                int tt = m - oe_del, e2 = e - e_del;
                d |= e2 > tt ? 1 << 2 : 0;
                ehe[j] = e2 > tt ? e2 : tt;
                tt = m - oe_ins;
                d |= (f - e_ins) > tt ? 2 << 4 : 0;
                if (zi) zi[j - beg] = (uint8_t)d;
            }
            const int last = (end - j0 < 32 ? end - j0 : 32) - 1;
            carryU = ::max(carryU, __shfl_sync(FULL, inc, 31));
            carryH = __shfl_sync(FULL, h, last);
        }
        if (lane == 0) { ehh[end] = carryH; ehe[end] = KSW_NEG_INF; }
        __syncwarp();
    }
    int score = ehh[qlen];
    __syncwarp();
    return score;
}
