// oracle/oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; PARITY UNPINNED).
//
// A plain C++17 restatement of the read-alignment path that bioseqdb reaches through
// bioseqdb/bwa.cpp (reference file) and, below it, the un-vendored third-party library
// lh3/bwa (branch Apache2, unpinned HEAD; reference Dockerfile:6) -- restated from the published
// algorithm.  Neither libbwa nor PostgreSQL exist in the build container,
// and the reference's own tests never call a bwa function (reference test/run.py:25-171), so this
// oracle is pinned only by the self-generated known answers listed in SURVEY.md section 8(c)
// (lrand48 / minstd_rand streams, NUCLSEQ payload bytes, brute-force suffix arrays, naive DP).
// => "parity unpinned" with respect to a real libbwa binary.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this code.  The product (bioseqdb_b200/) never includes or links anything from oracle/.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <string>

namespace orc {

// ---- options: every mem_opt_t field the path reads (reference extension.cpp:220-231)
struct Opts {
    int a = 1, b = 4;
    int o_del = 6, e_del = 1, o_ins = 6, e_ins = 1;
    int pen_clip5 = 5, pen_clip3 = 5;
    int w = 100, zdrop = 100;
    int min_seed_len = 19, max_occ = 500;
    int max_mem_intv = 20, split_width = 10;
    int max_chain_gap = 10000, max_chain_extend = 1 << 30, min_chain_weight = 0;
    float split_factor = 1.5f, mask_level = 0.50f, drop_ratio = 0.50f, mask_level_redun = 0.95f;
    float mapQ_coef_len = 50.f; int mapQ_coef_fac = 0;  // fac = (int)log(len) = 3 (an int field in bwamem.h), filled by opts_init()
    int8_t mat[25];
};
void opts_init(Opts& o);  // mem_opt_init() defaults incl. bwa_fill_scmat(a=1,b=4)

// ---- packed nucleotide type (reference sequence.h:18-38)
struct Hole { int64_t offset; int32_t len; char amb; };  // == bntamb1_t, 16 bytes
struct Nuclseq {
    uint32_t len = 0;
    std::vector<Hole> holes;
    std::vector<uint8_t> pac;  // ceil(len/4) bytes, MSB-first
};
int nt4(char c);                                   // nst_nt4_table
bool nuclseq_from_text(const std::string& s, Nuclseq& out, char* bad);  // sequence.cpp:209-245
std::string nuclseq_to_text(const Nuclseq& s);     // sequence.cpp:71-81
static inline uint8_t pac_get(const uint8_t* pac, uint64_t i) { return pac[i >> 2] >> ((~i & 3) << 1) & 3; }

// ---- index (bwa.cpp:82-128 + libbwa bwt_t/bntseq_t)
struct Ann { int64_t offset; int32_t len; int32_t n_ambs; int64_t id; };
struct Counters {
    uint64_t n_extend = 0, n_lf = 0, n_sa = 0;          // seeding roofline units
    uint64_t ext_cells = 0, ext_calls = 0, ext_rows = 0; // ksw_extend2
    uint64_t glb_cells = 0, glb_calls = 0;               // ksw_global2
    uint64_t sw_cells = 0, sw_calls = 0;                 // ksw_align2 (mem_seed_sw)
    uint64_t dup_chain_pos = 0;                          // A.5 equal-key corner events
    void add(const Counters& o);
};
struct Index {
    std::vector<uint8_t> pac;      // forward pac bytes, byte-rounded row concatenation
    std::vector<Ann> anns;
    std::vector<Hole> holes;       // NOT rebased (reference bug, SURVEY B#2)
    int64_t l_pac = 0;
    // bwt_t
    uint64_t seq_len = 0, primary = 0, L2[5] = {0, 0, 0, 0, 0};
    uint64_t bwt_size = 0;         // in u32 words, after Occ interleave
    std::vector<uint32_t> bwt;     // bwa layout: per 128 symbols {4 x u64 counts, 8 x u32 symbols}
    int sa_intv = 32;
    std::vector<uint64_t> sa;      // sampled, sa[0] = (uint64_t)-1
    bool built = false;
    void add_ref(int64_t id, const Nuclseq& s);   // bwa.cpp:82-105
    void build();                                  // bwa.cpp:107-128
    // alternative: adopt index arrays computed elsewhere (bench cpu_baseline reuses the GPU-built
    // index for the alignment timing; the arrays are mathematically unique)
    void adopt(const uint32_t* packed_bwt_plain, uint64_t primary, const uint64_t* sa_sampled);
};

// suffix array of T$ ($ smallest) by SA-IS; out has n+1 entries, out[0] = n
void suffix_array(const uint8_t* T, int64_t n, std::vector<int64_t>& out);

// ---- FM-index primitives
struct Intv { uint64_t x0, x1, x2, info; };
void bwt_occ4(const Index& ix, uint64_t k, uint64_t cnt[4]);
uint64_t bwt_occ(const Index& ix, uint64_t k, int c);
void bwt_extend(const Index& ix, const Intv& ik, Intv ok[4], int is_back, Counters* ctr);
void bwt_set_intv(const Index& ix, int c, Intv& ik);
uint64_t bwt_sa(const Index& ix, uint64_t k, Counters* ctr);

// ---- pipeline stages
struct Seed { int64_t rbeg; int32_t qbeg, len, score; };
struct Chain {
    int64_t pos; int rid; int first; uint32_t w; int kept; int is_alt; float frac_rep;
    std::vector<Seed> seeds;
};
struct Reg {  // mem_alnreg_t
    int64_t rb, re; int qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary,
        secondary_all, seedlen0; int n_comp, is_alt; float frac_rep; uint64_t hash;
};
struct Aln {  // mem_aln_t (fields the adapter or the parity gate reads)
    int64_t pos; int rid, flag, is_rev, mapq, NM, score, sub; std::vector<uint32_t> cigar;
};
struct Row {  // BwaMatch, reference bwa.h:15-30
    int64_t ref_id; std::string ref_subseq; int32_t ref_match_begin, ref_match_end, ref_match_len;
    std::string query_subseq; int32_t query_match_begin, query_match_end, query_match_len;
    bool is_primary, is_secondary, is_reverse; std::string cigar; int score;
};

void collect_intv(const Opts& o, const Index& ix, int len, const uint8_t* seq, std::vector<Intv>& mem, Counters* ctr);
void mem_chain(const Opts& o, const Index& ix, int len, const uint8_t* seq, std::vector<Chain>& chains,
               std::vector<Intv>* intv_out, std::vector<Seed>* seeds_out, Counters* ctr);
void mem_chain_flt(const Opts& o, std::vector<Chain>& a);
void mem_flt_chained_seeds(const Opts& o, const Index& ix, int l_query, const uint8_t* query, std::vector<Chain>& a, Counters* ctr);
void mem_chain2aln(const Opts& o, const Index& ix, int l_query, const uint8_t* query, const Chain& c, std::vector<Reg>& av, Counters* ctr);
void mem_sort_dedup_patch(const Opts& o, const Index& ix, uint8_t* query, std::vector<Reg>& a, Counters* ctr);
void mem_mark_primary_se(const Opts& o, std::vector<Reg>& a, int64_t id);
Aln mem_reg2aln(const Opts& o, const Index& ix, int l_query, const char* query_ascii, const Reg& ar, Counters* ctr);
int mem_approx_mapq_se(const Opts& o, const Reg& a);

// mem_align1 on an ASCII read; id = the value lrand48() would have returned
void mem_align1(const Opts& o, const Index& ix, int l_seq, const char* seq, int64_t id, std::vector<Reg>& regs, Counters* ctr);
// BwaIndex::align_sequence (bwa.cpp:141-181): rows in reference order
void align_sequence(const Opts& o, const Index& ix, const std::string& query_text, int64_t id,
                    std::vector<Reg>& regs, std::vector<Aln>& alns, std::vector<Row>* rows, Counters* ctr);

// ---- DP kernels -- also exported alone for kernel parity tests
int ksw_extend2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int* qle, int* tle, int* gtle, int* gscore, int* max_off, Counters* ctr);
int ksw_global2(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, std::vector<uint32_t>* cigar, Counters* ctr);
int ksw_local_score(int qlen, const uint8_t* query, int tlen, const uint8_t* target, int m, const int8_t* mat,
                    int o_del, int e_del, int o_ins, int e_ins, Counters* ctr);

// ---- libc / libstdc++ generators restated (SURVEY 8c golden values)
struct Lrand48 { uint64_t x = 0; long next() { x = (x * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL; return (long)(x >> 17); } };
struct MinstdRand { uint32_t s; explicit MinstdRand(uint32_t seed) { s = seed % 2147483647u; if (s == 0) s = 1; }
                    uint32_t next() { s = (uint32_t)((uint64_t)s * 48271u % 2147483647u); return s; } };
uint64_t hash_64(uint64_t key);

// klib ks_introsort, generic over a less-than functor
template <class T, class LT> void ks_introsort(size_t n, T* a, LT lt);

}  // namespace orc

#include "ksort.inl"
