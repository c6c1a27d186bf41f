// ksw_thread.cuh -- thread-per-extension ksw_extend2 (libbwa ksw.c as reached from mem_chain2aln,
// reference bioseqdb/bwa.cpp:149).
//
// For short reads one extension is a few hundred to a few thousand cells in rows of ~16 cells: a warp that sweeps
// such a row leaves half of its lanes idle and pays the per-row control work (trimming, z-drop, maxima) 32 times
// over.  Here every THREAD runs one extension with the scalar recurrence of the reference, literally, and the 32
// lanes of a warp run 32 extensions of (nearly) the same query length side by side (the job list is sorted by
// length, extend_plan.cu).  The {h, e} column state lives in shared memory laid out [column][thread] -- bank ==
// lane, conflict-free whatever column each lane is at -- packed as two unsigned 16-bit halves (all stored values
// are >= 0 and bounded by l_query * (a + 1), checked by the host); the query sits beside it as 4-bit codes, 8 per
// word, and slides through a register.  H = __vimax3_s32(M, E, F); E', F' = __viaddmax_s32_relu (DPX).
#pragma once
#include "common.cuh"
#include "ksw_warp.cuh"

// byte i of v -> nibble i of the result (every byte < 16)
__device__ __forceinline__ uint32_t nib_pack8(uint64_t v) {
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    v = v | (v >> 16);
    return (uint32_t)v;
}
__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
    const uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return (uint64_t)__byte_perm(lo, 0, 0x0123) << 32 | (uint64_t)__byte_perm(hi, 0, 0x0123);
}
// eight bytes at an arbitrarily aligned address; may touch up to 7 bytes beyond p + 8 (inside the same 8-byte granule
// sequence), never anything below p & ~7
__device__ __forceinline__ uint64_t load8_unaligned(const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint64_t* w = reinterpret_cast<const uint64_t*>(a & ~(uintptr_t)7);
    const unsigned sh = (unsigned)(a & 7) * 8;
    const uint64_t lo = w[0];
    if (sh == 0) return lo;
    return lo >> sh | w[1] << (64 - sh);
}

// Fills the caller's query words: qn[k * NT] holds the codes of extension columns 8k .. 8k+7.
//   dir > 0: column j is query[j]  (right extension, query points at the first base after the seed)
//   dir < 0: column j is query[-1 - j] (left extension, query points just past the last base before the seed);
//            `avail` = number of bases that exist below `query` (== qlen for the left extension)
// `query` must come from a buffer with >= 15 readable bytes behind its end (the batch read buffer has them).
template <int NT>
__device__ __forceinline__ void ksw_thread_load_query(uint32_t* qn, const uint8_t* query, int qlen, int dir) {
    const int nw = (qlen + 7) >> 3;
    if (dir > 0) {
        for (int k = 0; k < nw; ++k) qn[k * NT] = nib_pack8(load8_unaligned(query + 8 * k));
    } else {
        for (int k = 0; k < nw; ++k) {
            const int start = qlen - 8 * (k + 1);          // bytes [start, start + 8) below `query - qlen`'s origin, reversed
            uint64_t v;
            if (start >= 0) v = load8_unaligned(query - qlen + start);
            else v = load8_unaligned(query - qlen) << (unsigned)(-start * 8);   // never read below the read's first base
            qn[k * NT] = nib_pack8(bswap64(v));
        }
    }
}

// One extension.  eh / qn point at THIS thread's column 0 / word 0 (stride NT words); eh must hold qlen + 1 columns.
// tbase(i) returns the target code of row i.  Returns exactly what ksw_extend2 returns.
template <int NT, class TBase>
__device__ __forceinline__ ExtOut ksw_extend_thread(const DevOpts& o, uint32_t* eh, const uint32_t* qn, int qlen, int tlen, TBase tbase,
                                                    int w, int end_bonus, int h0, uint32_t& cells, uint32_t& rows) {
    const int e_del = o.e_del, e_ins = o.e_ins, o_del = o.o_del, o_ins = o.o_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    // bwa_fill_scmat: match / mismatch / ambiguous.  The score of a column is ONE byte-permute: the row's five scores (biased to be
    // non-negative) sit in the bytes of a 64-bit register pair and the query code selects one (bytes 5..7 are zero).
    const int sA = o.mat[0], sB = o.mat[1], sN = o.mat[4];
    const int bias = -::min(::min(sA, sB), ::min(sN, 0));
    const uint32_t lut_mis = (uint32_t)(sB + bias) * 0x01010101u, lut_amb = (uint32_t)(sN + bias) * 0x01010101u, lut_hi = (uint32_t)(sN + bias);
    // first row in closed form: eh[0].h = h0, eh[j].h = max(h0 - o_ins - j e_ins, 0), e = 0
    eh[0] = (uint32_t)h0;
    for (int j = 1; j <= qlen; ++j) { const int v = h0 - o_ins - j * e_ins; eh[j * NT] = (uint32_t)(v > 0 ? v : 0); }
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
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1 = 0;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        const int tb = tbase(i);
        const uint32_t lut_lo = tb > 3 ? lut_amb : lut_mis + ((uint32_t)(sA - sB) << (tb << 3));
        int f = 0, m = 0, mj = -1, first_nz = 0x7fffffff, last_nz = -1;
        cells += (uint32_t)(end > beg ? end - beg : 0); ++rows;
        int j = beg;
        if (beg < end) {
            uint32_t qw = qn[(beg >> 3) * NT] >> ((beg & 7) << 2);
            uint32_t* p = eh + beg * NT;
            for (; j < end; ++j, p += NT) {
                if ((j & 7) == 0) qw = qn[(j >> 3) * NT];
                const int sc = (int)__byte_perm(lut_lo, lut_hi, (qw & 7u) | 0x7770u) - bias;
                qw >>= 4;
                const uint32_t v = *p;
                int M = (int)(v & 0xffffu);
                const int e = (int)(v >> 16);
                M = M ? M + sc : 0;
                const int h = __vimax3_s32(M, e, f);
                const int e2 = __viaddmax_s32_relu(e, -e_del, M - oe_del);
                const uint32_t st = (uint32_t)h1 | (uint32_t)e2 << 16;
                *p = st;
                first_nz = st ? ::min(first_nz, j) : first_nz;
                last_nz = st ? j : last_nz;
                mj = m > h ? mj : j;
                m = ::max(m, h);
                f = __viaddmax_s32_relu(f, -e_ins, M - oe_ins);
                h1 = h;
            }
        }
        eh[end * NT] = (uint32_t)h1;          // eh[end] = {h1, 0}
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
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
        // band trimming for the next row: the two scans of the reference over eh[beg..end], tracked while the row was written
        const int nbeg = first_nz != 0x7fffffff ? first_nz : end;
        int jj;
        if (h1 != 0) jj = end;
        else if (last_nz >= 0) jj = last_nz;
        else jj = nbeg - 1;
        beg = nbeg;
        end = jj + 2 < qlen ? jj + 2 : qlen;
    }
    ExtOut r;
    r.score = max; r.qle = max_j + 1; r.tle = max_i + 1; r.gtle = max_ie + 1; r.gscore = gscore; r.max_off = max_off;
    return r;
}
