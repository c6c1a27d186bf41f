// oracle/mem.cpp -- TEST INFRASTRUCTURE (see oracle.h header; parity unpinned).
// Restatement of libbwa bwamem.c / bwa.c / bntseq.c as called from reference bioseqdb/bwa.cpp:141-181
// (mem_align1 + mem_reg2aln per region), following SURVEY.md Appendix A.4-A.12, A.14.
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace orc {

void opts_init(Opts& o) {
    o = Opts();
    o.mapQ_coef_fac = (int)log((double)o.mapQ_coef_len);   // bwamem.h declares `int mapQ_coef_fac`: mem_opt_init's log(50) = 3.912 is truncated to 3
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) o.mat[i * 5 + j] = i == j ? o.a : -o.b;
        o.mat[i * 5 + 4] = -1;
    }
    for (int j = 0; j < 5; ++j) o.mat[20 + j] = -1;
}

// ---------------------------------------------------------------- bntseq helpers
static inline int64_t bns_depos(const Index& ix, int64_t pos, int* is_rev) {
    return (*is_rev = (pos >= ix.l_pac)) ? (ix.l_pac << 1) - 1 - pos : pos;
}

static int bns_pos2rid(const Index& ix, int64_t pos_f) {
    int left, mid, right, n = (int)ix.anns.size();
    if (pos_f >= ix.l_pac) return -1;
    left = 0; mid = 0; right = n;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= ix.anns[mid].offset) {
            if (mid == n - 1) break;
            if (pos_f < ix.anns[mid + 1].offset) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

static int bns_intv2rid(const Index& ix, int64_t rb, int64_t re) {
    int is_rev, rid_b, rid_e;
    if (rb < ix.l_pac && re > ix.l_pac) return -2;
    rid_b = bns_pos2rid(ix, bns_depos(ix, rb, &is_rev));
    rid_e = rb < re ? bns_pos2rid(ix, bns_depos(ix, re - 1, &is_rev)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

static void bns_get_seq(const Index& ix, int64_t beg, int64_t end, std::vector<uint8_t>& seq) {
    int64_t l_pac = ix.l_pac;
    seq.clear();
    if (end < beg) std::swap(beg, end);
    if (end > (l_pac << 1)) end = l_pac << 1;
    if (beg < 0) beg = 0;
    if (beg >= l_pac || end <= l_pac) {
        if (end > beg) seq.reserve((size_t)(end - beg));
        if (beg >= l_pac) {
            int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg;
            for (int64_t k = end_f; k > beg_f; --k) seq.push_back(3 - pac_get(ix.pac.data(), (uint64_t)k));
        } else {
            for (int64_t k = beg; k < end; ++k) seq.push_back(pac_get(ix.pac.data(), (uint64_t)k));
        }
    }
}

static void bns_fetch_seq(const Index& ix, int64_t* beg, int64_t mid, int64_t* end, int* rid, std::vector<uint8_t>& seq) {
    int64_t far_beg, far_end;
    int is_rev;
    if (*end < *beg) std::swap(*beg, *end);
    *rid = bns_pos2rid(ix, bns_depos(ix, mid, &is_rev));
    far_beg = ix.anns[*rid].offset;
    far_end = far_beg + ix.anns[*rid].len;
    if (is_rev) {
        int64_t tmp = far_beg;
        far_beg = (ix.l_pac << 1) - far_end;
        far_end = (ix.l_pac << 1) - tmp;
    }
    *beg = *beg > far_beg ? *beg : far_beg;
    *end = *end < far_end ? *end : far_end;
    // libbwa asserts *beg <= *end here; a seed lying entirely in inter-row filler could violate it.
    // Restated as an empty fetch (see DESIGN.md "filler corner").
    if (*end < *beg) { seq.clear(); *end = *beg; return; }
    bns_get_seq(ix, *beg, *end, seq);
}

// ---------------------------------------------------------------- seeding
static int bwt_smem1(const Index& ix, int len, const uint8_t* q, int x, int min_intv, std::vector<Intv>& mem,
                     std::vector<Intv>& va, std::vector<Intv>& vb, Counters* ctr) {
    int i, c, ret;
    Intv ik, ok[4];
    std::vector<Intv>*prev = &va, *curr = &vb;
    mem.clear();
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    bwt_set_intv(ix, q[x], ik);
    ik.info = (uint64_t)(x + 1);
    curr->clear();
    for (i = x + 1; i < len; ++i) {  // forward search (max_intv == 0 on this path)
        if (q[i] < 4) {
            c = 3 - q[i];
            bwt_extend(ix, ik, ok, 0, ctr);
            if (ok[c].x2 != ik.x2) {
                curr->push_back(ik);
                if (ok[c].x2 < (uint64_t)min_intv) break;
            }
            ik = ok[c]; ik.info = (uint64_t)(i + 1);
        } else {
            curr->push_back(ik);
            break;
        }
    }
    if (i == len) curr->push_back(ik);
    std::reverse(curr->begin(), curr->end());
    ret = (int)(*curr)[0].info;
    std::swap(curr, prev);
    for (i = x - 1; i >= -1; --i) {  // backward search for MEMs
        c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
        curr->clear();
        for (size_t j = 0; j < prev->size(); ++j) {
            const Intv& p = (*prev)[j];
            if (c >= 0) bwt_extend(ix, p, ok, 1, ctr);
            if (c < 0 || ok[c].x2 < (uint64_t)min_intv) {
                if (curr->empty()) {
                    if (mem.empty() || (uint64_t)(i + 1) < mem.back().info >> 32) {
                        ik = p; ik.info |= (uint64_t)(i + 1) << 32;
                        mem.push_back(ik);
                    }
                }
            } else if (curr->empty() || ok[c].x2 != curr->back().x2) {
                ok[c].info = p.info;
                curr->push_back(ok[c]);
            }
        }
        if (curr->empty()) break;
        std::swap(curr, prev);
    }
    std::reverse(mem.begin(), mem.end());
    return ret;
}

static int bwt_seed_strategy1(const Index& ix, int len, const uint8_t* q, int x, int min_len, int max_intv, Intv* mem, Counters* ctr) {
    int i, c;
    Intv ik, ok[4];
    memset(mem, 0, sizeof(Intv));
    if (q[x] > 3) return x + 1;
    bwt_set_intv(ix, q[x], ik);
    for (i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            c = 3 - q[i];
            bwt_extend(ix, ik, ok, 0, ctr);
            if (ok[c].x2 < (uint64_t)max_intv && i - x >= min_len) {
                *mem = ok[c];
                mem->info = (uint64_t)x << 32 | (uint64_t)(i + 1);
                return i + 1;
            }
            ik = ok[c];
        } else return i + 1;
    }
    return len;
}

void collect_intv(const Opts& opt, const Index& ix, int len, const uint8_t* seq, std::vector<Intv>& mem, Counters* ctr) {
    int x = 0, start_width = 1;
    int split_len = (int)(opt.min_seed_len * opt.split_factor + .499);
    std::vector<Intv> mem1, va, vb;
    mem.clear();
    while (x < len) {  // pass 1: all SMEMs
        if (seq[x] < 4) {
            x = bwt_smem1(ix, len, seq, x, start_width, mem1, va, vb, ctr);
            for (const Intv& p : mem1) {
                int slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
                if (slen >= opt.min_seed_len) mem.push_back(p);
            }
        } else ++x;
    }
    size_t old_n = mem.size();  // pass 2: re-seeding inside long SMEMs
    for (size_t k = 0; k < old_n; ++k) {
        Intv p = mem[k];
        int start = (int)(p.info >> 32), end = (int32_t)p.info;
        if (end - start < split_len || p.x2 > (uint64_t)opt.split_width) continue;
        bwt_smem1(ix, len, seq, (start + end) >> 1, (int)(p.x2 + 1), mem1, va, vb, ctr);
        for (const Intv& r : mem1)
            if ((int)((uint32_t)r.info - (uint32_t)(r.info >> 32)) >= opt.min_seed_len) mem.push_back(r);
    }
    if (opt.max_mem_intv > 0) {  // pass 3: LAST-like
        x = 0;
        while (x < len) {
            if (seq[x] < 4) {
                Intv m;
                x = bwt_seed_strategy1(ix, len, seq, x, opt.min_seed_len, opt.max_mem_intv, &m, ctr);
                if (m.x2 > 0) mem.push_back(m);
            } else ++x;
        }
    }
    ks_introsort(mem.size(), mem.data(), [](const Intv& a, const Intv& b) { return a.info < b.info; });
}

// ---------------------------------------------------------------- chaining
static int test_and_merge(const Opts& opt, int64_t l_pac, Chain& c, const Seed& p, int seed_rid) {
    int64_t qend, rend, x, y;
    const Seed& last = c.seeds.back();
    qend = last.qbeg + last.len;
    rend = last.rbeg + last.len;
    if (seed_rid != c.rid) return 0;
    if (p.qbeg >= c.seeds[0].qbeg && p.qbeg + p.len <= qend && p.rbeg >= c.seeds[0].rbeg && p.rbeg + p.len <= rend)
        return 1;
    if ((last.rbeg < l_pac || c.seeds[0].rbeg < l_pac) && p.rbeg >= l_pac) return 0;
    x = p.qbeg - last.qbeg;
    y = p.rbeg - last.rbeg;
    if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - last.len < opt.max_chain_gap && y - last.len < opt.max_chain_gap) {
        c.seeds.push_back(p);
        return 1;
    }
    return 0;
}

void mem_chain(const Opts& opt, const Index& ix, int len, const uint8_t* seq, std::vector<Chain>& chains,
               std::vector<Intv>* intv_out, std::vector<Seed>* seeds_out, Counters* ctr) {
    chains.clear();
    if (intv_out) intv_out->clear();
    if (seeds_out) seeds_out->clear();
    if (len < opt.min_seed_len) return;
    std::vector<Intv> mem;
    collect_intv(opt, ix, len, seq, mem, ctr);
    if (intv_out) *intv_out = mem;
    int b = 0, e = 0, l_rep = 0;
    for (const Intv& p : mem) {
        int sb = (int)(p.info >> 32), se = (int)(uint32_t)p.info;
        if (p.x2 <= (uint64_t)opt.max_occ) continue;
        if (sb > e) { l_rep += e - b; b = sb; e = se; }
        else e = e > se ? e : se;
    }
    l_rep += e - b;
    // The ordered container: libbwa keeps chains in a B-tree keyed by pos.  Restated as a vector kept
    // sorted by pos; for equal keys the newcomer goes after its equals and lookups return the last
    // element with pos <= key (SURVEY A.5 corner; counted in dup_chain_pos).
    for (const Intv& p : mem) {
        int step, count, slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
        int64_t k;
        step = p.x2 > (uint64_t)opt.max_occ ? (int)(p.x2 / (uint64_t)opt.max_occ) : 1;
        for (k = count = 0; (uint64_t)k < p.x2 && count < opt.max_occ; k += step, ++count) {
            Seed s;
            s.rbeg = (int64_t)bwt_sa(ix, p.x0 + (uint64_t)k, ctr);
            s.qbeg = (int)(p.info >> 32);
            s.score = s.len = slen;
            int rid = bns_intv2rid(ix, s.rbeg, s.rbeg + s.len);
            if (rid < 0) continue;
            if (seeds_out) seeds_out->push_back(s);
            bool to_add = false;
            if (!chains.empty()) {
                // upper_bound by pos, then step back
                size_t lo = 0, hi = chains.size();
                while (lo < hi) { size_t md = (lo + hi) >> 1; if (chains[md].pos <= s.rbeg) lo = md + 1; else hi = md; }
                if (lo == 0) to_add = true;
                else {
                    Chain& lower = chains[lo - 1];
                    if (!test_and_merge(opt, ix.l_pac, lower, s, rid)) {
                        to_add = true;
                        if (lower.pos == s.rbeg && ctr) ++ctr->dup_chain_pos;
                    }
                }
                if (to_add) {
                    Chain c; c.pos = s.rbeg; c.rid = rid; c.first = -1; c.w = 0; c.kept = 0; c.is_alt = 0; c.frac_rep = 0;
                    c.seeds.push_back(s);
                    chains.insert(chains.begin() + lo, std::move(c));
                }
            } else {
                Chain c; c.pos = s.rbeg; c.rid = rid; c.first = -1; c.w = 0; c.kept = 0; c.is_alt = 0; c.frac_rep = 0;
                c.seeds.push_back(s);
                chains.push_back(std::move(c));
            }
        }
    }
    for (Chain& c : chains) c.frac_rep = (float)l_rep / len;
}

// ---------------------------------------------------------------- chain filter
static int mem_chain_weight(const Chain& c) {
    int64_t end;
    int w = 0, tmp;
    size_t j;
    for (j = 0, end = 0; j < c.seeds.size(); ++j) {
        const Seed& s = c.seeds[j];
        if (s.qbeg >= end) w += s.len;
        else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
        end = end > s.qbeg + s.len ? end : s.qbeg + s.len;
    }
    tmp = w; w = 0;
    for (j = 0, end = 0; j < c.seeds.size(); ++j) {
        const Seed& s = c.seeds[j];
        if (s.rbeg >= end) w += s.len;
        else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
        end = end > s.rbeg + s.len ? end : s.rbeg + s.len;
    }
    w = w < tmp ? w : tmp;
    return w < 1 << 30 ? w : (1 << 30) - 1;
}

#define chn_beg(ch) ((ch).seeds.front().qbeg)
#define chn_end(ch) ((ch).seeds.back().qbeg + (ch).seeds.back().len)

void mem_chain_flt(const Opts& opt, std::vector<Chain>& a) {
    int i, k, n_chn = (int)a.size();
    if (n_chn == 0) return;
    std::vector<int> kept_idx;
    for (i = k = 0; i < n_chn; ++i) {
        Chain& c = a[i];
        c.first = -1; c.kept = 0;
        c.w = (uint32_t)mem_chain_weight(c);
        if ((int)c.w < opt.min_chain_weight) continue;
        if (k != i) a[k] = std::move(a[i]);
        ++k;
    }
    n_chn = k; a.resize(n_chn);
    if (n_chn == 0) return;
    ks_introsort((size_t)n_chn, a.data(), [](const Chain& x, const Chain& y) { return x.w > y.w; });
    a[0].kept = 3;
    kept_idx.push_back(0);
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        size_t kk;
        for (kk = 0; kk < kept_idx.size(); ++kk) {
            int j = kept_idx[kk];
            int b_max = chn_beg(a[j]) > chn_beg(a[i]) ? chn_beg(a[j]) : chn_beg(a[i]);
            int e_min = chn_end(a[j]) < chn_end(a[i]) ? chn_end(a[j]) : chn_end(a[i]);
            if (e_min > b_max && (!a[j].is_alt || a[i].is_alt)) {
                int li = chn_end(a[i]) - chn_beg(a[i]);
                int lj = chn_end(a[j]) - chn_beg(a[j]);
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
                    large_ovlp = 1;
                    if (a[j].first < 0) a[j].first = i;
                    if ((int)a[i].w < (int)a[j].w * opt.drop_ratio && (int)a[j].w - (int)a[i].w >= opt.min_seed_len << 1) break;
                }
            }
        }
        if (kk == kept_idx.size()) {
            kept_idx.push_back(i);
            a[i].kept = large_ovlp ? 2 : 3;
        }
    }
    for (size_t q = 0; q < kept_idx.size(); ++q) {
        Chain& c = a[kept_idx[q]];
        if (c.first >= 0) a[c.first].kept = 1;
    }
    for (i = k = 0; i < n_chn; ++i) {
        if (a[i].kept == 0 || a[i].kept == 3) continue;
        if (++k >= opt.max_chain_extend) break;
    }
    for (; i < n_chn; ++i) if (a[i].kept < 3) a[i].kept = 0;
    for (i = k = 0; i < n_chn; ++i) {
        if (a[i].kept == 0) continue;
        if (k != i) a[k] = std::move(a[i]);
        ++k;
    }
    a.resize(k);
}

#define MEM_SHORT_EXT 50
#define MEM_SHORT_LEN 200
#define MEM_HSP_COEF 1.1f
#define MEM_MINSC_COEF 5.5f
#define MEM_SEEDSW_COEF 0.05f

static int mem_seed_sw(const Opts& opt, const Index& ix, int l_query, const uint8_t* query, const Seed& s, Counters* ctr) {
    int qb, qe, rid;
    int64_t rb, re, mid, l_pac = ix.l_pac;
    if (s.len >= MEM_SHORT_LEN) return -1;
    qb = s.qbeg; qe = s.qbeg + s.len;
    rb = s.rbeg; re = s.rbeg + s.len;
    mid = (rb + re) >> 1;
    qb -= MEM_SHORT_EXT; qb = qb > 0 ? qb : 0;
    qe += MEM_SHORT_EXT; qe = qe < l_query ? qe : l_query;
    rb -= MEM_SHORT_EXT; rb = rb > 0 ? rb : 0;
    re += MEM_SHORT_EXT; re = re < (l_pac << 1) ? re : (l_pac << 1);
    if (rb < l_pac && l_pac < re) {
        if (mid < l_pac) re = l_pac;
        else rb = l_pac;
    }
    if (qe - qb >= MEM_SHORT_LEN || re - rb >= MEM_SHORT_LEN) return -1;
    std::vector<uint8_t> rseq;
    bns_fetch_seq(ix, &rb, mid, &re, &rid, rseq);
    return ksw_local_score(qe - qb, query + qb, (int)(re - rb), rseq.data(), 5, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, ctr);
}

void mem_flt_chained_seeds(const Opts& opt, const Index& ix, int l_query, const uint8_t* query, std::vector<Chain>& a, Counters* ctr) {
    double min_l = opt.min_chain_weight ? MEM_HSP_COEF * opt.min_chain_weight : MEM_MINSC_COEF * log(l_query);
    int min_HSP_score = (int)(opt.a * min_l + .499);
    if (min_l > MEM_SEEDSW_COEF * l_query) return;
    for (Chain& c : a) {
        size_t k = 0;
        for (size_t j = 0; j < c.seeds.size(); ++j) {
            Seed& s = c.seeds[j];
            s.score = mem_seed_sw(opt, ix, l_query, query, s, ctr);
            if (s.score < 0 || s.score >= min_HSP_score) {
                s.score = s.score < 0 ? s.len * opt.a : s.score;
                c.seeds[k++] = s;
            }
        }
        c.seeds.resize(k);
    }
}

// ---------------------------------------------------------------- extension
static inline int cal_max_gap(const Opts& opt, int qlen) {
    int l_del = (int)((double)(qlen * opt.a - opt.o_del) / opt.e_del + 1.);
    int l_ins = (int)((double)(qlen * opt.a - opt.o_ins) / opt.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < opt.w << 1 ? l : opt.w << 1;
}

#define MAX_BAND_TRY 2

void mem_chain2aln(const Opts& opt, const Index& ix, int l_query, const uint8_t* query, const Chain& c, std::vector<Reg>& av, Counters* ctr) {
    int i, k, rid, max_off[2], aw[2];
    int64_t l_pac = ix.l_pac, rmax[2], tmp, max = 0;
    int n = (int)c.seeds.size();
    if (n == 0) return;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (i = 0; i < n; ++i) {
        int64_t b, e;
        const Seed& t = c.seeds[i];
        b = t.rbeg - (t.qbeg + cal_max_gap(opt, t.qbeg));
        e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(opt, l_query - t.qbeg - t.len));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
        if (t.len > max) max = t.len;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < (l_pac << 1) ? rmax[1] : (l_pac << 1);
    if (rmax[0] < l_pac && l_pac < rmax[1]) {
        if (c.seeds[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    std::vector<uint8_t> rseq;
    bns_fetch_seq(ix, &rmax[0], c.seeds[0].rbeg, &rmax[1], &rid, rseq);
    int64_t rlen = rmax[1] - rmax[0];

    std::vector<uint64_t> srt((size_t)n);
    for (i = 0; i < n; ++i) srt[i] = (uint64_t)c.seeds[i].score << 32 | (uint64_t)i;
    ks_introsort((size_t)n, srt.data(), [](uint64_t x, uint64_t y) { return x < y; });

    for (k = n - 1; k >= 0; --k) {
        const Seed* s = &c.seeds[(uint32_t)srt[k]];
        for (i = 0; i < (int)av.size(); ++i) {
            const Reg* p = &av[i];
            int64_t rd;
            int qd, w, max_gap;
            if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
            if (s->len - p->seedlen0 > .1 * l_query) continue;
            qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
            qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
            max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
            w = max_gap < p->w ? max_gap : p->w;
            if (qd - rd < w && rd - qd < w) break;
        }
        if (i < (int)av.size()) {
            for (i = k + 1; i < n; ++i) {
                if (srt[i] == 0) continue;
                const Seed* t = &c.seeds[(uint32_t)srt[i]];
                if (t->len < s->len * .95) continue;
                if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
            }
            if (i == n) { srt[k] = 0; continue; }
        }
        av.emplace_back();
        Reg* a = &av.back();
        memset(a, 0, sizeof(Reg));
        a->w = aw[0] = aw[1] = opt.w;
        a->score = a->truesc = -1;
        a->rid = c.rid;

        if (s->qbeg) {  // left extension
            int qle, tle, gtle, gscore;
            std::vector<uint8_t> qs((size_t)s->qbeg), rs;
            for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
            tmp = s->rbeg - rmax[0];
            if (tmp > 0) { rs.resize((size_t)tmp); for (int64_t ii = 0; ii < tmp; ++ii) rs[ii] = rseq[tmp - 1 - ii]; }
            for (i = 0; i < MAX_BAND_TRY; ++i) {
                int prev = a->score;
                aw[0] = opt.w << i;
                a->score = ksw_extend2(s->qbeg, qs.data(), (int)tmp, rs.data(), 5, opt.mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins,
                                       aw[0], opt.pen_clip5, opt.zdrop, s->len * opt.a, &qle, &tle, &gtle, &gscore, &max_off[0], ctr);
                if (a->score == prev || max_off[0] < (aw[0] >> 1) + (aw[0] >> 2)) break;
            }
            if (gscore <= 0 || gscore <= a->score - opt.pen_clip5) {
                a->qb = s->qbeg - qle; a->rb = s->rbeg - tle;
                a->truesc = a->score;
            } else {
                a->qb = 0; a->rb = s->rbeg - gtle;
                a->truesc = gscore;
            }
        } else { a->score = a->truesc = s->len * opt.a; a->qb = 0; a->rb = s->rbeg; }

        if (s->qbeg + s->len != l_query) {  // right extension
            int qle, tle, qe, gtle, gscore, sc0 = a->score;
            int64_t re;
            qe = s->qbeg + s->len;
            re = s->rbeg + s->len - rmax[0];
            for (i = 0; i < MAX_BAND_TRY; ++i) {
                int prev = a->score;
                aw[1] = opt.w << i;
                int64_t tl = rlen - re;
                a->score = ksw_extend2(l_query - qe, query + qe, (int)tl, tl > 0 ? rseq.data() + re : nullptr, 5, opt.mat, opt.o_del, opt.e_del,
                                       opt.o_ins, opt.e_ins, aw[1], opt.pen_clip3, opt.zdrop, sc0, &qle, &tle, &gtle, &gscore, &max_off[1], ctr);
                if (a->score == prev || max_off[1] < (aw[1] >> 1) + (aw[1] >> 2)) break;
            }
            if (gscore <= 0 || gscore <= a->score - opt.pen_clip3) {
                a->qe = qe + qle; a->re = rmax[0] + re + tle;
                a->truesc += a->score - sc0;
            } else {
                a->qe = l_query; a->re = rmax[0] + re + gtle;
                a->truesc += gscore - sc0;
            }
        } else { a->qe = l_query; a->re = s->rbeg + s->len; }

        for (i = 0, a->seedcov = 0; i < n; ++i) {
            const Seed& t = c.seeds[i];
            if (t.qbeg >= a->qb && t.qbeg + t.len <= a->qe && t.rbeg >= a->rb && t.rbeg + t.len <= a->re) a->seedcov += t.len;
        }
        a->w = aw[0] > aw[1] ? aw[0] : aw[1];
        a->seedlen0 = s->len;
        a->frac_rep = c.frac_rep;
    }
}

// ---------------------------------------------------------------- global alignment wrapper
// returns false when libbwa would return without touching *score (rejected input)
static bool bwa_gen_cigar2(const Opts& opt, const Index& ix, int w_, int l_query, uint8_t* query, int64_t rb, int64_t re,
                           int* score, std::vector<uint32_t>* cigar, int* NM, Counters* ctr) {
    int64_t l_pac = ix.l_pac;
    const int8_t* mat = opt.mat;
    if (cigar) cigar->clear();
    if (NM) *NM = -1;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return false;
    std::vector<uint8_t> rseq;
    bns_get_seq(ix, rb, re, rseq);
    int64_t rlen = (int64_t)rseq.size();
    if (re - rb != rlen) return false;
    if (rb >= l_pac) {
        std::reverse(query, query + l_query);
        std::reverse(rseq.begin(), rseq.end());
    }
    if (l_query == re - rb && w_ == 0) {
        if (cigar) cigar->push_back((uint32_t)l_query << 4 | 0);
        *score = 0;
        for (int i = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - opt.o_ins) / opt.e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - opt.o_del) / opt.e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        max_gap = max_gap > 1 ? max_gap : 1;
        w = (max_gap + abs((int)rlen - l_query) + 1) >> 1;
        w = w < w_ ? w : w_;
        min_w = abs((int)rlen - l_query) + 3;
        w = w > min_w ? w : min_w;
        *score = ksw_global2(l_query, query, (int)rlen, rseq.data(), 5, mat, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, w, cigar, ctr);
    }
    if (NM && cigar) {
        int x = 0, y = 0, n_mm = 0, n_gap = 0, nc = (int)cigar->size();
        for (int k = 0; k < nc; ++k) {
            int op = (*cigar)[k] & 0xf, len = (int)((*cigar)[k] >> 4);
            if (op == 0) {
                for (int i = 0; i < len; ++i) if (query[x + i] != rseq[y + i]) ++n_mm;
                x += len; y += len;
            } else if (op == 2) {
                if (k > 0 && k < nc - 1) n_gap += len;
                y += len;
            } else if (op == 1) { x += len; n_gap += len; }
        }
        *NM = n_mm + n_gap;
    }
    if (rb >= l_pac) std::reverse(query, query + l_query);
    return true;
}

// ---------------------------------------------------------------- dedup / patch
#define PATCH_MAX_R_BW 0.05f
#define PATCH_MIN_SC_RATIO 0.90f

static int mem_patch_reg(const Opts& opt, const Index& ix, uint8_t* query, const Reg* a, const Reg* b, int* _w, Counters* ctr) {
    int w, score = 0, q_s, r_s;
    double r;
    if (a->rb < ix.l_pac && b->rb >= ix.l_pac) return 0;
    if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;
    w = (int)((a->re - b->rb) - (a->qe - b->qb));
    w = w > 0 ? w : -w;
    r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);
    r = r > 0. ? r : -r;
    if (a->re < b->rb || a->qe < b->qb) {
        if (w > opt.w << 1 || r >= PATCH_MAX_R_BW) return 0;
    } else if (w > opt.w << 2 || r >= PATCH_MAX_R_BW * 2) return 0;
    w += a->w + b->w;
    w = w < opt.w << 2 ? w : opt.w << 2;
    bwa_gen_cigar2(opt, ix, w, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, nullptr, nullptr, ctr);
    q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
    r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
    if ((double)score / (q_s > r_s ? q_s : r_s) < PATCH_MIN_SC_RATIO) return 0;
    *_w = w;
    return score;
}

void mem_sort_dedup_patch(const Opts& opt, const Index& ix, uint8_t* query, std::vector<Reg>& av, Counters* ctr) {
    int m, i, j, n = (int)av.size();
    if (n <= 1) return;
    Reg* a = av.data();
    ks_introsort((size_t)n, a, [](const Reg& x, const Reg& y) { return x.re < y.re; });
    for (i = 0; i < n; ++i) a[i].n_comp = 1;
    for (i = 1; i < n; ++i) {
        Reg* p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + opt.max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + opt.max_chain_gap; --j) {
            Reg* q = &a[j];
            int64_t orr, oq, mr, mq;
            int score, w;
            if (q->qe == q->qb) continue;
            orr = q->re - p->rb;
            oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
            mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if (orr > opt.mask_level_redun * mr && oq > opt.mask_level_redun * mq) {
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (q->rb < p->rb && (score = mem_patch_reg(opt, ix, query, q, p, &w, ctr)) > 0) {
                p->n_comp += q->n_comp + 1;
                p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
                p->sub = p->sub > q->sub ? p->sub : q->sub;
                p->csub = p->csub > q->csub ? p->csub : q->csub;
                p->qb = q->qb; p->rb = q->rb;
                p->truesc = p->score = score;
                p->w = w;
                q->qb = q->qe;
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    n = m;
    ks_introsort((size_t)n, a, [](const Reg& x, const Reg& y) {
        return x.score > y.score || (x.score == y.score && (x.rb < y.rb || (x.rb == y.rb && x.qb < y.qb)));
    });
    for (i = 1; i < n; ++i)
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    av.resize((size_t)(n < 1 ? n : m));
}

void mem_mark_primary_se(const Opts& opt, std::vector<Reg>& av, int64_t id) {
    int i, n = (int)av.size();
    if (n == 0) return;
    Reg* a = av.data();
    for (i = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1;
        a[i].hash = hash_64((uint64_t)(id + i));
    }
    ks_introsort((size_t)n, a, [](const Reg& x, const Reg& y) {
        return x.score > y.score || (x.score == y.score && (x.is_alt < y.is_alt || (x.is_alt == y.is_alt && x.hash < y.hash)));
    });
    int tmp = opt.a + opt.b;
    tmp = opt.o_del + opt.e_del > tmp ? opt.o_del + opt.e_del : tmp;
    tmp = opt.o_ins + opt.e_ins > tmp ? opt.o_ins + opt.e_ins : tmp;
    std::vector<int> z;
    z.push_back(0);
    for (i = 1; i < n; ++i) {
        size_t k;
        for (k = 0; k < z.size(); ++k) {
            int j = z[k];
            int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * opt.mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == z.size()) z.push_back(i);
        else a[i].secondary = z[k];
    }
    for (i = 0; i < n; ++i) a[i].secondary_all = i;
}

// ---------------------------------------------------------------- mem_reg2aln, MAPQ
int mem_approx_mapq_se(const Opts& opt, const Reg& a) {
    int mapq, l, sub = a.sub ? a.sub : opt.min_seed_len * opt.a;
    double identity;
    sub = a.csub > sub ? a.csub : sub;
    if (sub >= a.score) return 0;
    l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
    identity = 1. - (double)(l * opt.a - a.score) / (opt.a + opt.b) / l;
    if (a.score == 0) mapq = 0;
    else if (opt.mapQ_coef_len > 0) {
        double tmp;
        tmp = l < opt.mapQ_coef_len ? 1. : opt.mapQ_coef_fac / log(l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a.score - sub) / opt.a * tmp * tmp + .499);
    } else {
        mapq = (int)(30.0 * (1. - (double)sub / a.score) * log(a.seedcov) + .499);
        mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
    }
    if (a.sub_n > 0) mapq -= (int)(4.343 * log(a.sub_n + 1) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - a.frac_rep) + .499);
    return mapq;
}

static inline int infer_bw(int l1, int l2, int score, int a, int q, int r) {
    int w;
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    if (w < abs(l1 - l2)) w = abs(l1 - l2);
    return w;
}

Aln mem_reg2aln(const Opts& opt, const Index& ix, int l_query, const char* query_, const Reg& ar, Counters* ctr) {
    Aln a;
    int i, w2, tmp, qb, qe, NM = -1, score = 0, is_rev, last_sc = -(1 << 30);
    int64_t pos, rb, re;
    a.pos = -1; a.rid = -1; a.flag = 0; a.is_rev = 0; a.mapq = 0; a.NM = 0; a.score = 0; a.sub = 0;
    if (ar.rb < 0 || ar.re < 0) { a.flag |= 0x4; return a; }
    qb = ar.qb; qe = ar.qe; rb = ar.rb; re = ar.re;
    std::vector<uint8_t> query((size_t)l_query);
    for (i = 0; i < l_query; ++i) query[i] = (uint8_t)(query_[i] < 5 ? query_[i] : nt4(query_[i]));
    a.mapq = ar.secondary < 0 ? mem_approx_mapq_se(opt, ar) : 0;
    if (ar.secondary >= 0) a.flag |= 0x100;
    tmp = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt.a, opt.o_del, opt.e_del);
    w2 = infer_bw(qe - qb, (int)(re - rb), ar.truesc, opt.a, opt.o_ins, opt.e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > opt.w) w2 = w2 < ar.w ? w2 : ar.w;
    i = 0;
    do {
        w2 = w2 < opt.w << 2 ? w2 : opt.w << 2;
        bwa_gen_cigar2(opt, ix, w2, qe - qb, query.data() + qb, rb, re, &score, &a.cigar, &NM, ctr);
        if (score == last_sc || w2 == opt.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar.truesc - opt.a);
    a.NM = NM;
    pos = bns_depos(ix, rb < ix.l_pac ? rb : re - 1, &is_rev);
    a.is_rev = is_rev;
    if (!a.cigar.empty()) {
        if ((a.cigar[0] & 0xf) == 2) { pos += a.cigar[0] >> 4; a.cigar.erase(a.cigar.begin()); }
        else if ((a.cigar.back() & 0xf) == 2) a.cigar.pop_back();
    }
    if (qb != 0 || qe != l_query) {
        int clip5, clip3;
        clip5 = is_rev ? l_query - qe : qb;
        clip3 = is_rev ? qb : l_query - qe;
        if (clip5) a.cigar.insert(a.cigar.begin(), (uint32_t)clip5 << 4 | 3);
        if (clip3) a.cigar.push_back((uint32_t)clip3 << 4 | 3);
    }
    a.rid = bns_pos2rid(ix, pos);
    a.pos = pos - ix.anns[a.rid].offset;
    a.score = ar.score; a.sub = ar.sub > ar.csub ? ar.sub : ar.csub;
    return a;
}

// ---------------------------------------------------------------- drivers
void mem_align1(const Opts& opt, const Index& ix, int l_seq, const char* seq_, int64_t id, std::vector<Reg>& regs, Counters* ctr) {
    std::vector<uint8_t> seq((size_t)l_seq);
    for (int i = 0; i < l_seq; ++i) seq[i] = (uint8_t)(seq_[i] < 4 ? seq_[i] : nt4(seq_[i]));
    std::vector<Chain> chn;
    mem_chain(opt, ix, l_seq, seq.data(), chn, nullptr, nullptr, ctr);
    mem_chain_flt(opt, chn);
    mem_flt_chained_seeds(opt, ix, l_seq, seq.data(), chn, ctr);
    regs.clear();
    for (const Chain& c : chn) mem_chain2aln(opt, ix, l_seq, seq.data(), c, regs, ctr);
    mem_sort_dedup_patch(opt, ix, seq.data(), regs, ctr);
    mem_mark_primary_se(opt, regs, id);
}

static std::string extract_reference_subseq(const Index& ix, int64_t rb, int64_t re) {
    // bwa.cpp:55-68.  Forward-strand hits only read pac_forward in range; for reverse-strand hits the
    // reference indexes past the vector (UB, SURVEY B#3) -- defined here as the reverse-strand text.
    std::string s((size_t)(re - rb), '?');
    for (int64_t i = 0; i < re - rb; ++i) {
        int64_t p = rb + i;
        int c = p < ix.l_pac ? pac_get(ix.pac.data(), (uint64_t)p) : 3 - pac_get(ix.pac.data(), (uint64_t)((ix.l_pac << 1) - 1 - p));
        s[(size_t)i] = "ACGT"[c];
    }
    for (const Hole& h : ix.holes) {
        int64_t l = std::max<int64_t>(h.offset, rb), r = std::min<int64_t>(h.offset + h.len, re);
        for (int64_t i = l; i < r; ++i) s[(size_t)(i - rb)] = h.amb;
    }
    return s;
}

void align_sequence(const Opts& opt, const Index& ix, const std::string& query, int64_t id,
                    std::vector<Reg>& regs, std::vector<Aln>& alns, std::vector<Row>* rows, Counters* ctr) {
    regs.clear(); alns.clear();
    if (rows) rows->clear();
    if (ix.pac.empty()) return;
    mem_align1(opt, ix, (int)query.size(), query.data(), id, regs, ctr);
    for (const Reg& r : regs) {
        Aln d = mem_reg2aln(opt, ix, (int)query.size(), query.data(), r, ctr);
        if (rows) {
            int64_t ref_offset = ix.anns[r.rid].offset;
            Row row;
            row.ref_id = ix.anns[r.rid].id;
            row.ref_subseq = extract_reference_subseq(ix, r.rb, r.re);
            row.ref_match_begin = (int32_t)(r.rb - ref_offset);
            row.ref_match_end = (int32_t)(r.re - ref_offset);
            row.ref_match_len = (int32_t)(r.re - r.rb);
            row.query_subseq = query.substr((size_t)r.qb, (size_t)(r.qe - r.qb));
            row.query_match_begin = r.qb; row.query_match_end = r.qe; row.query_match_len = r.qe - r.qb;
            row.is_primary = (d.flag & 0x100) == 0;
            row.is_secondary = (d.flag & 0x100) != 0;
            row.is_reverse = d.is_rev != 0;
            for (uint32_t c : d.cigar) { row.cigar += std::to_string(c >> 4); row.cigar += "MIDNSHP=XB"[c & 0xf]; }
            row.score = d.score;
            rows->push_back(std::move(row));
        }
        alns.push_back(std::move(d));
    }
}

}  // namespace orc
