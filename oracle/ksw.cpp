// oracle/ksw.cpp -- TEST INFRASTRUCTURE (see oracle.h header; parity unpinned).
// Scalar restatement of libbwa ksw.c: ksw_extend2, ksw_global2 (A.11) and the
// score of ksw_align2 (local SW used by mem_seed_sw, A.6).
#include "oracle.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>

namespace orc {

struct eh_t { int32_t h, e; };

int ksw_extend2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int* _qle, int* _tle, int* _gtle, int* _gscore, int* _max_off, Counters* ctr) {
    int i, j, k, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, beg, end, max, max_i, max_j, max_ins, max_del,
        max_ie, gscore, max_off;
    if (qlen < 0) qlen = 0;
    std::vector<eh_t> eh((size_t)qlen + 1, eh_t{0, 0});
    // first row
    eh[0].h = h0;
    if (qlen >= 1) eh[1].h = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && eh[j - 1].h > e_ins; ++j) eh[j].h = eh[j - 1].h - e_ins;
    // clamp the band
    k = m * m;
    for (i = 0, max = 0; i < k; ++i) max = max > mat[i] ? max : mat[i];
    max_ins = (int)((double)(qlen * max + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    max_del = (int)((double)(qlen * max + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    // DP
    max = h0; max_i = max_j = -1; max_ie = -1; gscore = -1; max_off = 0;
    beg = 0; end = qlen;
    if (ctr) ++ctr->ext_calls;
    for (i = 0; i < tlen; ++i) {
        int t, f = 0, h1, mm = 0, mj = -1;
        const int8_t* q = &mat[target[i] * m];
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        if (beg == 0) { h1 = h0 - (o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        if (ctr) { ctr->ext_cells += (uint64_t)(end > beg ? end - beg : 0); ++ctr->ext_rows; }
        for (j = beg; j < end; ++j) {
            eh_t* p = &eh[j];
            int h, M = p->h, e = p->e;
            p->h = h1;
            M = M ? M + q[query[j]] : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            mj = mm > h ? mj : j;
            mm = mm > h ? mm : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            p->e = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
        }
        eh[end].h = h1; eh[end].e = 0;
        if (j == qlen) {
            max_ie = gscore > h1 ? max_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (mm == 0) break;
        if (mm > max) {
            max = mm; max_i = i; max_j = mj;
            max_off = max_off > abs(mj - i) ? max_off : abs(mj - i);
        } else if (zdrop > 0) {
            if (i - max_i > mj - max_j) {
                if (max - mm - ((i - max_i) - (mj - max_j)) * e_del > zdrop) break;
            } else {
                if (max - mm - ((mj - max_j) - (i - max_i)) * e_ins > zdrop) break;
            }
        }
        for (j = beg; j < end && eh[j].h == 0 && eh[j].e == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && eh[j].h == 0 && eh[j].e == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    if (_qle) *_qle = max_j + 1;
    if (_tle) *_tle = max_i + 1;
    if (_gtle) *_gtle = max_ie + 1;
    if (_gscore) *_gscore = gscore;
    if (_max_off) *_max_off = max_off;
    return max;
}

#define MINUS_INF (-0x40000000)

int ksw_global2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, std::vector<uint32_t>* cigar_out, Counters* ctr) {
    int i, j, oe_del = o_del + e_del, oe_ins = o_ins + e_ins, score, n_col;
    if (cigar_out) cigar_out->clear();
    n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    std::vector<uint8_t> z;
    if (cigar_out) z.assign((size_t)n_col * (size_t)(tlen > 0 ? tlen : 0), 0);
    std::vector<eh_t> eh((size_t)qlen + 1);
    eh[0].h = 0; eh[0].e = MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = MINUS_INF; }
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = MINUS_INF;
    if (ctr) ++ctr->glb_calls;
    for (i = 0; i < tlen; ++i) {
        int32_t f = MINUS_INF, h1, beg, end, t;
        const int8_t* q = &mat[target[i] * m];
        beg = i > w ? i - w : 0;
        end = i + w + 1 < qlen ? i + w + 1 : qlen;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : MINUS_INF;
        if (ctr) ctr->glb_cells += (uint64_t)(end > beg ? end - beg : 0);
        uint8_t* zi = cigar_out ? &z[(size_t)i * n_col] : nullptr;
        for (j = beg; j < end; ++j) {
            eh_t* p = &eh[j];
            int32_t h, mm = p->h, e = p->e;
            uint8_t d;
            p->h = h1;
            mm += q[query[j]];
            d = mm >= e ? 0 : 1;
            h = mm >= e ? mm : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = mm - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = mm - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            if (zi) zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = MINUS_INF;
    }
    score = eh[qlen].h;
    if (cigar_out) {
        std::vector<uint32_t>& cg = *cigar_out;
        auto push = [&](int op, int len) {
            if (cg.empty() || op != (int)(cg.back() & 0xf)) cg.push_back((uint32_t)len << 4 | (uint32_t)op);
            else cg.back() += (uint32_t)len << 4;
        };
        int which = 0, kk;
        i = tlen - 1; kk = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        while (i >= 0 && kk >= 0) {
            which = z[(size_t)i * n_col + (kk - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { push(0, 1); --i; --kk; }
            else if (which == 1) { push(2, 1); --i; }
            else { push(1, 1); --kk; }
        }
        if (i >= 0) push(2, i + 1);
        if (kk >= 0) push(1, kk + 1);
        std::reverse(cg.begin(), cg.end());
    }
    return score;
}

// Score of ksw_align2(xtra = KSW_XSTART, i.e. no minimum score): the optimal local alignment score
// with affine gaps opening from H (Farrar's striped kernels compute exactly this value; the u8 kernel
// saturates and is re-run in i16 by ksw_align2, so the returned score is the unsaturated one).
int ksw_local_score(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                    int o_del, int e_del, int o_ins, int e_ins, Counters* ctr) {
    int oe_del = o_del + e_del, oe_ins = o_ins + e_ins, best = 0;
    std::vector<int> H((size_t)qlen + 1, 0), E((size_t)qlen + 1, 0);
    if (ctr) { ++ctr->sw_calls; ctr->sw_cells += (uint64_t)qlen * (uint64_t)(tlen > 0 ? tlen : 0); }
    for (int i = 0; i < tlen; ++i) {
        const int8_t* q = &mat[target[i] * m];
        int f = 0, hdiag = 0;  // H(i-1, j-1)
        for (int j = 1; j <= qlen; ++j) {
            int h = hdiag + q[query[j - 1]];
            hdiag = H[j];
            int e = E[j];
            h = h > e ? h : e;
            h = h > f ? h : f;
            h = h > 0 ? h : 0;
            H[j] = h;
            best = best > h ? best : h;
            int t = h - oe_del; e -= e_del; E[j] = std::max(0, e > t ? e : t);
            t = h - oe_ins; f -= e_ins; f = std::max(0, f > t ? f : t);
        }
    }
    return best;
}

}  // namespace orc
