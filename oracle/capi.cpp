// oracle/capi.cpp -- TEST INFRASTRUCTURE (see oracle.h header; parity unpinned).
// Flat C entry points over the oracle so that tests/ and bench.py's cpu_baseline leg can drive it
// through ctypes.  The row layout matches include/bioseqdb_gpu.h's bsq_row so results compare
// byte for byte; nothing here is linked into the product library.
#include "oracle.h"
#include <cstring>
#include <thread>
#include <chrono>

using namespace orc;

extern "C" {

struct orc_row {  // == bsq_row (include/bioseqdb_gpu.h)
    int64_t rb, re, pos; uint64_t hash;
    int32_t qb, qe, rid, score, truesc, sub, csub, sub_n, w, seedcov, secondary, seedlen0, n_comp;
    float frac_rep;
    int32_t is_rev, mapq, NM, flag;
    uint32_t cigar_off, n_cigar;
    int64_t ref_id;
};

struct orc_opts {  // == bsq_opts
    int32_t min_seed_len, max_occ, a, b, pen_clip3, pen_clip5, zdrop, w, o_del, e_del, o_ins, e_ins;
};

struct orc_handle {
    Opts opt;
    Index ix;
};

struct orc_result {
    std::vector<uint64_t> row_off;
    std::vector<orc_row> rows;
    std::vector<uint32_t> cigar;
    Counters ctr;
    double seconds = 0;
};

static void apply(Opts& o, const orc_opts* p) {
    opts_init(o);  // note: mat is NOT refreshed after a/b change (SURVEY B#5)
    if (!p) return;
    o.min_seed_len = p->min_seed_len; o.max_occ = p->max_occ; o.a = p->a; o.b = p->b;
    o.pen_clip3 = p->pen_clip3; o.pen_clip5 = p->pen_clip5; o.zdrop = p->zdrop; o.w = p->w;
    o.o_del = p->o_del; o.e_del = p->e_del; o.o_ins = p->o_ins; o.e_ins = p->e_ins;
}

orc_handle* orc_new(const orc_opts* p) { auto* h = new orc_handle; apply(h->opt, p); return h; }
void orc_free(orc_handle* h) { delete h; }
void orc_set_opts(orc_handle* h, const orc_opts* p) { apply(h->opt, p); }

int orc_add_ref_text(orc_handle* h, int64_t id, const char* text, uint64_t len) {
    Nuclseq s; char bad = 0;
    if (!nuclseq_from_text(std::string(text, len), s, &bad)) return (int)(unsigned char)bad ? (int)(unsigned char)bad : -1;
    h->ix.add_ref(id, s);
    return 0;
}
int orc_add_ref_packed(orc_handle* h, int64_t id, const uint8_t* pac, uint32_t len, const Hole* holes, uint32_t n_holes) {
    Nuclseq s; s.len = len; s.pac.assign(pac, pac + (len + 3) / 4); s.holes.assign(holes, holes + n_holes);
    h->ix.add_ref(id, s);
    return 0;
}
double orc_build(orc_handle* h) {
    auto t0 = std::chrono::steady_clock::now();
    h->ix.build();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void orc_adopt(orc_handle* h, const uint32_t* plain_bwt, uint64_t primary, const uint64_t* sa) { h->ix.adopt(plain_bwt, primary, sa); }

// index inspection: what[0]=l_pac, 1=seq_len, 2=primary, 3..7=L2, 8=bwt_size(u32 words), 9=n_sa, 10=n_anns, 11=n_holes
void orc_index_info(const orc_handle* h, uint64_t* what) {
    const Index& ix = h->ix;
    what[0] = (uint64_t)ix.l_pac; what[1] = ix.seq_len; what[2] = ix.primary;
    for (int i = 0; i < 5; ++i) what[3 + i] = ix.L2[i];
    what[8] = ix.bwt_size; what[9] = ix.sa.size(); what[10] = ix.anns.size(); what[11] = ix.holes.size();
}
const uint32_t* orc_index_bwt(const orc_handle* h) { return h->ix.bwt.data(); }
const uint64_t* orc_index_sa(const orc_handle* h) { return h->ix.sa.data(); }
const uint8_t* orc_index_pac(const orc_handle* h) { return h->ix.pac.data(); }
void orc_index_anns(const orc_handle* h, int64_t* offset, int32_t* len, int64_t* id) {
    for (size_t i = 0; i < h->ix.anns.size(); ++i) { offset[i] = h->ix.anns[i].offset; len[i] = h->ix.anns[i].len; id[i] = h->ix.anns[i].id; }
}
// plain (non-interleaved) BWT, 16 symbols per u32 MSB-first, as produced at bwa.cpp:48-50
void orc_index_bwt_plain(const orc_handle* h, uint32_t* out) {
    const Index& ix = h->ix;
    uint64_t nw = (ix.seq_len + 15) / 16;
    for (uint64_t i = 0; i < nw; ++i) out[i] = ix.bwt[((i * 16) >> 7 << 4) + 8 + (((i * 16) & 0x7f) >> 4)];
}

void orc_suffix_array(const uint8_t* T, int64_t n, int64_t* out) {
    std::vector<int64_t> sa; suffix_array(T, n, sa); memcpy(out, sa.data(), sizeof(int64_t) * (size_t)(n + 1));
}
uint64_t orc_bwt_sa(const orc_handle* h, uint64_t k) { return bwt_sa(h->ix, k, nullptr); }
void orc_bwt_occ4(const orc_handle* h, uint64_t k, uint64_t* cnt) { bwt_occ4(h->ix, k, cnt); }

// ---- codec
int orc_nuclseq_from_text(const char* text, uint64_t len, uint8_t* pac_out, Hole* holes_out, uint32_t holes_cap, uint32_t* n_holes) {
    Nuclseq s; char bad = 0;
    if (!nuclseq_from_text(std::string(text, len), s, &bad)) return (int)(unsigned char)bad ? (int)(unsigned char)bad : -1;
    memcpy(pac_out, s.pac.data(), s.pac.size());
    *n_holes = (uint32_t)s.holes.size();
    for (uint32_t i = 0; i < s.holes.size() && i < holes_cap; ++i) { memset(&holes_out[i], 0, sizeof(Hole)); holes_out[i] = s.holes[i]; }
    return 0;
}
void orc_nuclseq_to_text(const uint8_t* pac, uint32_t len, const Hole* holes, uint32_t n_holes, char* out) {
    Nuclseq s; s.len = len; s.pac.assign(pac, pac + (len + 3) / 4); s.holes.assign(holes, holes + n_holes);
    std::string t = nuclseq_to_text(s); memcpy(out, t.data(), t.size());
}

// ---- generators
void orc_lrand48(int n, int64_t* out) { Lrand48 g; for (int i = 0; i < n; ++i) out[i] = g.next(); }
void orc_minstd(uint32_t seed, int n, uint32_t* out) { MinstdRand g(seed); for (int i = 0; i < n; ++i) out[i] = g.next(); }
uint64_t orc_hash64(uint64_t k) { return hash_64(k); }
void orc_introsort_u64(uint64_t n, uint64_t* keys_hi32_payload_lo32) {
    // sorts by the high 32 bits only (payload in the low 32 bits shows the instability pattern)
    ks_introsort((size_t)n, keys_hi32_payload_lo32, [](uint64_t a, uint64_t b) { return (a >> 32) < (b >> 32); });
}

// ---- DP kernels alone
int orc_ksw_extend2(int qlen, const uint8_t* q, int tlen, const uint8_t* t, const orc_opts* p, int w, int end_bonus, int h0, int* out5) {
    Opts o; apply(o, p);
    return ksw_extend2(qlen, q, tlen, t, 5, o.mat, o.o_del, o.e_del, o.o_ins, o.e_ins, w, end_bonus, o.zdrop, h0,
                       &out5[0], &out5[1], &out5[2], &out5[3], &out5[4], nullptr);
}
int orc_ksw_global2(int qlen, const uint8_t* q, int tlen, const uint8_t* t, const orc_opts* p, int w, uint32_t* cigar, int cap, int* n_cigar) {
    Opts o; apply(o, p);
    std::vector<uint32_t> cg;
    int sc = ksw_global2(qlen, q, tlen, t, 5, o.mat, o.o_del, o.e_del, o.o_ins, o.e_ins, w, cigar ? &cg : nullptr, nullptr);
    if (cigar) { *n_cigar = (int)cg.size(); for (int i = 0; i < (int)cg.size() && i < cap; ++i) cigar[i] = cg[i]; }
    return sc;
}
int orc_ksw_local(int qlen, const uint8_t* q, int tlen, const uint8_t* t, const orc_opts* p) {
    Opts o; apply(o, p);
    return ksw_local_score(qlen, q, tlen, t, 5, o.mat, o.o_del, o.e_del, o.o_ins, o.e_ins, nullptr);
}

// ---- stage dumps for one read (tests): intervals / seeds / post-filter chains
// intv: 4 x u64 each; seeds: rbeg,qbeg,len as 3 x i64; chains: per chain (pos, rid, n_seeds, w, kept) 5 x i64 then seeds 3 x i64
int64_t orc_stage_dump(const orc_handle* h, const char* seq, int len, uint64_t* intv, int64_t intv_cap, int64_t* n_intv,
                       int64_t* seeds, int64_t seeds_cap, int64_t* n_seeds, int64_t* chains, int64_t chains_cap) {
    std::vector<uint8_t> s((size_t)len);
    for (int i = 0; i < len; ++i) s[i] = (uint8_t)(seq[i] < 4 ? seq[i] : nt4(seq[i]));
    std::vector<Chain> chn; std::vector<Intv> iv; std::vector<Seed> sd;
    mem_chain(h->opt, h->ix, len, s.data(), chn, &iv, &sd, nullptr);
    *n_intv = (int64_t)iv.size();
    for (size_t i = 0; i < iv.size() && (int64_t)i < intv_cap; ++i) memcpy(intv + 4 * i, &iv[i], 32);
    *n_seeds = (int64_t)sd.size();
    for (size_t i = 0; i < sd.size() && (int64_t)i < seeds_cap; ++i) { seeds[3 * i] = sd[i].rbeg; seeds[3 * i + 1] = sd[i].qbeg; seeds[3 * i + 2] = sd[i].len; }
    mem_chain_flt(h->opt, chn);
    mem_flt_chained_seeds(h->opt, h->ix, len, s.data(), chn, nullptr);
    int64_t k = 0;
    for (const Chain& c : chn) {
        if (k + 5 + 3 * (int64_t)c.seeds.size() > chains_cap) return -1;
        chains[k++] = c.pos; chains[k++] = c.rid; chains[k++] = (int64_t)c.seeds.size(); chains[k++] = c.w; chains[k++] = c.kept;
        for (const Seed& q : c.seeds) { chains[k++] = q.rbeg; chains[k++] = q.qbeg; chains[k++] = q.len; }
    }
    return k;
}

// ---- batch alignment (the CPU baseline): reads = concatenated ASCII, offs[n+1], ids[n]
orc_result* orc_align_batch(const orc_handle* h, const char* seqs, const uint64_t* offs, const int64_t* ids, uint64_t n, int n_threads) {
    auto* res = new orc_result;
    if (n_threads < 1) n_threads = 1;
    struct Part { std::vector<uint64_t> nrows; std::vector<orc_row> rows; std::vector<uint32_t> cigar; Counters ctr; };
    std::vector<Part> parts((size_t)n_threads);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int t) {
        Part& P = parts[(size_t)t];
        uint64_t lo = n * (uint64_t)t / (uint64_t)n_threads, hi = n * (uint64_t)(t + 1) / (uint64_t)n_threads;
        std::vector<Reg> regs; std::vector<Aln> alns;
        for (uint64_t r = lo; r < hi; ++r) {
            std::string q(seqs + offs[r], (size_t)(offs[r + 1] - offs[r]));
            align_sequence(h->opt, h->ix, q, ids[r], regs, alns, nullptr, &P.ctr);
            P.nrows.push_back(regs.size());
            for (size_t i = 0; i < regs.size(); ++i) {
                const Reg& g = regs[i]; const Aln& a = alns[i];
                orc_row w; memset(&w, 0, sizeof(w));
                w.rb = g.rb; w.re = g.re; w.pos = a.pos; w.hash = g.hash; w.qb = g.qb; w.qe = g.qe; w.rid = g.rid; w.score = g.score;
                w.truesc = g.truesc; w.sub = g.sub; w.csub = g.csub; w.sub_n = g.sub_n; w.w = g.w; w.seedcov = g.seedcov;
                w.secondary = g.secondary; w.seedlen0 = g.seedlen0; w.n_comp = g.n_comp; w.frac_rep = g.frac_rep;
                w.is_rev = a.is_rev; w.mapq = a.mapq; w.NM = a.NM; w.flag = a.flag;
                w.cigar_off = (uint32_t)P.cigar.size(); w.n_cigar = (uint32_t)a.cigar.size();
                w.ref_id = h->ix.anns[(size_t)g.rid].id;
                P.cigar.insert(P.cigar.end(), a.cigar.begin(), a.cigar.end());
                P.rows.push_back(w);
            }
        }
    };
    if (n_threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    res->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res->row_off.push_back(0);
    for (Part& P : parts) {
        uint32_t cbase = (uint32_t)res->cigar.size();
        for (uint64_t c : P.nrows) res->row_off.push_back(res->row_off.back() + c);
        for (orc_row w : P.rows) { w.cigar_off += cbase; res->rows.push_back(w); }
        res->cigar.insert(res->cigar.end(), P.cigar.begin(), P.cigar.end());
        res->ctr.add(P.ctr);
    }
    return res;
}
void orc_result_info(const orc_result* r, uint64_t* out) {  // n_rows, n_cigar, then 12 counters
    out[0] = r->rows.size(); out[1] = r->cigar.size();
    const Counters& c = r->ctr;
    uint64_t v[12] = {c.n_extend, c.n_lf, c.n_sa, c.ext_cells, c.ext_calls, c.ext_rows, c.glb_cells, c.glb_calls, c.sw_cells, c.sw_calls, c.dup_chain_pos, 0};
    memcpy(out + 2, v, sizeof(v));
}
double orc_result_seconds(const orc_result* r) { return r->seconds; }
const uint64_t* orc_result_row_off(const orc_result* r) { return r->row_off.data(); }
const orc_row* orc_result_rows(const orc_result* r) { return r->rows.data(); }
const uint32_t* orc_result_cigar(const orc_result* r) { return r->cigar.data(); }
void orc_result_free(orc_result* r) { delete r; }

// reference-shaped rows for one read (BwaMatch, bwa.h:15-30) rendered as text lines for golden files
int64_t orc_align_rows_text(const orc_handle* h, const char* seq, uint64_t len, int64_t id, char* out, int64_t cap) {
    std::vector<Reg> regs; std::vector<Aln> alns; std::vector<Row> rows;
    align_sequence(h->opt, h->ix, std::string(seq, len), id, regs, alns, &rows, nullptr);
    std::string s;
    for (const Row& r : rows) {
        s += std::to_string(r.ref_id) + "\t" + r.ref_subseq + "\t" + std::to_string(r.ref_match_begin) + "\t" + std::to_string(r.ref_match_end) + "\t" +
             std::to_string(r.ref_match_len) + "\t" + r.query_subseq + "\t" + std::to_string(r.query_match_begin) + "\t" + std::to_string(r.query_match_end) +
             "\t" + std::to_string(r.query_match_len) + "\t" + (r.is_primary ? "t" : "f") + "\t" + (r.is_secondary ? "t" : "f") + "\t" +
             (r.is_reverse ? "t" : "f") + "\t" + r.cigar + "\t" + std::to_string(r.score) + "\n";
    }
    if ((int64_t)s.size() + 1 > cap) return -(int64_t)s.size();
    memcpy(out, s.c_str(), s.size() + 1);
    return (int64_t)s.size();
}

}  // extern "C"
