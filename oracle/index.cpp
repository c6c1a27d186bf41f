// oracle/index.cpp -- TEST INFRASTRUCTURE (see oracle.h header; parity unpinned).
// Packed nucleotide codec (reference bioseqdb/sequence.cpp:46-81,209-245), reference concatenation
// (bioseqdb/bwa.cpp:82-105), text/BWT/Occ/SA construction (bioseqdb/bwa.cpp:20-53,107-128 and
// libbwa is.c / bwtindex.c as restated in SURVEY.md A.1-A.3) and the FM-index queries (A.2/A.3).
#include "oracle.h"
#include <algorithm>
#include <cstring>
#include <cassert>

namespace orc {

void Counters::add(const Counters& o) {
    n_extend += o.n_extend; n_lf += o.n_lf; n_sa += o.n_sa;
    ext_cells += o.ext_cells; ext_calls += o.ext_calls; ext_rows += o.ext_rows;
    glb_cells += o.glb_cells; glb_calls += o.glb_calls;
    sw_cells += o.sw_cells; sw_calls += o.sw_calls; dup_chain_pos += o.dup_chain_pos;
}

// ---------------------------------------------------------------- codec
int nt4(char ch) {  // libbwa nst_nt4_table: ACGT (either case) -> 0..3, '-' -> 5, everything else 4
    switch (ch) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

static inline void pac_or(uint8_t* pac, uint64_t i, uint8_t v) { pac[i >> 2] |= v << ((~i & 3) << 1); }

bool nuclseq_from_text(const std::string& str, Nuclseq& out, char* bad) {
    static const char* allowed = "ACGTNWSMKRYBDHV";  // sequence.h:16, checked in extension.cpp:52-57
    for (char c : str)
        if (!c || !strchr(allowed, c)) { if (bad) *bad = c; return false; }
    // calculate_num_of_holes (sequence.cpp:46-57)
    uint32_t holes_num = 0; char prev = 0;
    for (char c : str) { if (prev != c && nt4(c) >= 4) ++holes_num; prev = c; }
    out.len = (uint32_t)str.size();
    out.holes.clear(); out.holes.reserve(holes_num);
    out.pac.assign((str.size() + 3) / 4, 0);
    MinstdRand rng(holes_num ^ (uint32_t)str.size());
    prev = 0;
    for (uint32_t i = 0; i < str.size(); ++i) {
        char c = str[i]; int code = nt4(c);
        if (code >= 4) {
            if (prev == c) out.holes.back().len++;
            else out.holes.push_back(Hole{(int64_t)i, 1, c});
            pac_or(out.pac.data(), i, rng.next() & 3);
        } else pac_or(out.pac.data(), i, (uint8_t)code);
        prev = c;
    }
    for (uint64_t i = out.len; i < out.pac.size() * 4; ++i) pac_or(out.pac.data(), i, rng.next() & 3);
    return true;
}

std::string nuclseq_to_text(const Nuclseq& s) {
    std::string t(s.len, '?');
    for (uint32_t i = 0; i < s.len; ++i) t[i] = "ACGT"[pac_get(s.pac.data(), i)];
    for (const Hole& h : s.holes) std::fill(t.begin() + h.offset, t.begin() + h.offset + h.len, h.amb);
    return t;
}

// ---------------------------------------------------------------- SA-IS (Nong, Zhang, Chan 2009)
namespace {
template <class Ch, class Ix>
void sais_core(const Ch* s, Ix* SA, Ix n, Ix K) {
    // s[n-1] must be the unique smallest symbol
    std::vector<uint8_t> t((size_t)n);  // 1 = S-type
    t[n - 1] = 1;
    for (Ix i = n - 2; i >= 0; --i) t[i] = (s[i] < s[i + 1] || (s[i] == s[i + 1] && t[i + 1])) ? 1 : 0;
    auto is_lms = [&](Ix i) { return i > 0 && t[i] && !t[i - 1]; };
    std::vector<Ix> cnt((size_t)K, 0), bkt((size_t)K);
    for (Ix i = 0; i < n; ++i) ++cnt[s[i]];
    auto bucket_ends = [&]() { Ix sum = 0; for (Ix c = 0; c < K; ++c) { sum += cnt[c]; bkt[c] = sum; } };
    auto bucket_starts = [&]() { Ix sum = 0; for (Ix c = 0; c < K; ++c) { bkt[c] = sum; sum += cnt[c]; } };
    auto induce = [&]() {
        bucket_starts();
        for (Ix i = 0; i < n; ++i) { Ix j = SA[i] - 1; if (SA[i] > 0 && !t[j]) SA[bkt[s[j]]++] = j; }
        bucket_ends();
        for (Ix i = n - 1; i >= 0; --i) { Ix j = SA[i] - 1; if (SA[i] > 0 && t[j]) SA[--bkt[s[j]]] = j; }
    };
    // stage 1: sort LMS substrings
    std::fill(SA, SA + n, (Ix)-1);
    bucket_ends();
    for (Ix i = 1; i < n; ++i) if (is_lms(i)) SA[--bkt[s[i]]] = i;
    induce();
    Ix n1 = 0;
    for (Ix i = 0; i < n; ++i) if (is_lms(SA[i])) SA[n1++] = SA[i];
    std::fill(SA + n1, SA + n, (Ix)-1);
    Ix name = 0, prev = -1;
    for (Ix i = 0; i < n1; ++i) {
        Ix pos = SA[i]; bool diff = false;
        if (prev < 0) diff = true;
        else for (Ix d = 0;; ++d) {
            if (s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d]) { diff = true; break; }
            if (d > 0 && (is_lms(pos + d) || is_lms(prev + d))) break;
        }
        if (diff) { ++name; prev = pos; }
        SA[n1 + (pos >> 1)] = name - 1;
    }
    for (Ix i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];
    Ix* SA1 = SA; Ix* s1 = SA + n - n1;
    if (name < n1) sais_core<Ix, Ix>(s1, SA1, n1, name);
    else for (Ix i = 0; i < n1; ++i) SA1[s1[i]] = i;
    // stage 3: induce the final SA from the sorted LMS suffixes
    for (Ix i = 1, j = 0; i < n; ++i) if (is_lms(i)) s1[j++] = i;
    for (Ix i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    std::fill(SA + n1, SA + n, (Ix)-1);
    bucket_ends();
    for (Ix i = n1 - 1; i >= 0; --i) { Ix j = SA[i]; SA[i] = -1; SA[--bkt[s[j]]] = j; }
    induce();
}
}  // namespace

void suffix_array(const uint8_t* T, int64_t n, std::vector<int64_t>& out) {
    // shift symbols by +1 so that the appended sentinel 0 is the unique smallest
    std::vector<uint8_t> s((size_t)n + 1);
    for (int64_t i = 0; i < n; ++i) s[i] = T[i] + 1;
    s[n] = 0;
    out.resize((size_t)n + 1);
    if (n + 1 < (int64_t)0x7fffffff) {
        std::vector<int32_t> sa((size_t)n + 1);
        sais_core<uint8_t, int32_t>(s.data(), sa.data(), (int32_t)(n + 1), 256);
        for (int64_t i = 0; i <= n; ++i) out[i] = sa[i];
    } else {
        sais_core<uint8_t, int64_t>(s.data(), out.data(), n + 1, 256);
    }
}

// ---------------------------------------------------------------- reference concatenation + build
void Index::add_ref(int64_t id, const Nuclseq& s) {
    int64_t offset = (int64_t)pac.size() * 4;
    anns.push_back(Ann{offset, (int32_t)s.len, (int32_t)s.holes.size(), id});
    pac.insert(pac.end(), s.pac.begin(), s.pac.begin() + (s.len + 3) / 4);
    for (const Hole& h : s.holes) holes.push_back(h);  // offsets stay row-relative (SURVEY B#2)
}

static void occ_interleave(Index& ix, const std::vector<uint32_t>& plain) {
    // bwt_bwtupdate_core: before every 128 symbols, 4 x u64 counts of symbols in B[0, block start)
    uint64_t n = ix.seq_len;
    uint64_t n_occ = (n + 127) / 128 + 1;
    ix.bwt_size = (n + 15) / 16 + n_occ * 8;
    ix.bwt.assign(ix.bwt_size, 0);
    uint64_t c[4] = {0, 0, 0, 0};
    uint64_t k = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i % 128 == 0) { memcpy(&ix.bwt[k], c, 32); k += 8; }
        if (i % 16 == 0) ix.bwt[k++] = plain[i / 16];
        ++c[plain[i >> 4] >> ((~i & 0xf) << 1) & 3];
    }
    memcpy(&ix.bwt[k], c, 32); k += 8;
    assert(k == ix.bwt_size);
}

void Index::build() {
    if (pac.empty()) return;
    l_pac = (int64_t)pac.size() * 4;
    seq_len = (uint64_t)l_pac * 2;
    std::vector<uint8_t> T(seq_len);
    memset(L2, 0, sizeof(L2));
    for (int64_t i = 0; i < l_pac; ++i) {           // bwa.cpp:34-38
        T[i] = pac_get(pac.data(), (uint64_t)i);
        L2[1 + T[i]]++; L2[4 - T[i]]++;
    }
    for (int64_t i = l_pac - 1; i >= 0; --i) T[2 * l_pac - 1 - i] = 3 - pac_get(pac.data(), (uint64_t)i);  // :41-42
    for (int i = 2; i <= 4; ++i) L2[i] += L2[i - 1];
    std::vector<int64_t> SA;
    suffix_array(T.data(), (int64_t)seq_len, SA);
    // BWT with the $ row dropped (is_bwt semantics)
    std::vector<uint32_t> plain((seq_len + 15) / 16, 0);
    uint64_t out = 0;
    for (uint64_t r = 0; r <= seq_len; ++r) {
        if (SA[r] == 0) { primary = r; continue; }
        uint32_t c = T[SA[r] - 1];
        plain[out >> 4] |= c << ((15 - (out & 15)) << 1);
        ++out;
    }
    occ_interleave(*this, plain);
    // bwt_cal_sa(bwt, 32): sa[k/32] = SA[k] for k % 32 == 0, sa[0] = -1; restated as the
    // literal LF walk so that invPsi is exercised, then cross-checked against SA in the tests.
    uint64_t n_sa = (seq_len + sa_intv) / sa_intv;
    sa.assign(n_sa, 0);
    {
        uint64_t isa = 0, sav = seq_len;
        for (uint64_t i = 0; i < seq_len; ++i) {
            if (isa % sa_intv == 0) sa[isa / sa_intv] = sav;
            --sav;
            // invPsi
            if (isa == primary) isa = 0;
            else {
                uint64_t kk = isa < primary ? isa : isa - 1;
                int c = bwt[((kk >> 7) << 4) + 8 + ((kk & 0x7f) >> 4)] >> ((~kk & 0xf) << 1) & 3;
                isa = L2[c] + bwt_occ(*this, isa, c);
            }
        }
        if (isa % sa_intv == 0) sa[isa / sa_intv] = sav;
        sa[0] = (uint64_t)-1;
    }
    built = true;
}

void Index::adopt(const uint32_t* plain_bwt, uint64_t prim, const uint64_t* sa_sampled) {
    l_pac = (int64_t)pac.size() * 4;
    seq_len = (uint64_t)l_pac * 2;
    memset(L2, 0, sizeof(L2));
    for (int64_t i = 0; i < l_pac; ++i) { int c = pac_get(pac.data(), (uint64_t)i); L2[1 + c]++; L2[4 - c]++; }
    for (int i = 2; i <= 4; ++i) L2[i] += L2[i - 1];
    primary = prim;
    std::vector<uint32_t> plain(plain_bwt, plain_bwt + (seq_len + 15) / 16);
    occ_interleave(*this, plain);
    uint64_t n_sa = (seq_len + sa_intv) / sa_intv;
    sa.assign(sa_sampled, sa_sampled + n_sa);
    built = true;
}

// ---------------------------------------------------------------- Occ queries
static inline void count_word(uint32_t w, uint32_t keep, uint64_t cnt[4]) {
    // keep: mask (on the 0x55555555 lattice) of the symbols to be counted
    uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
    cnt[0] += __builtin_popcount(~hi & ~lo & keep);
    cnt[1] += __builtin_popcount(~hi & lo & keep);
    cnt[2] += __builtin_popcount(hi & ~lo & keep);
    cnt[3] += __builtin_popcount(hi & lo & keep);
}

void bwt_occ4(const Index& ix, uint64_t k, uint64_t cnt[4]) {
    if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    k -= (k >= ix.primary);
    const uint32_t* p = &ix.bwt[(k >> 7) << 4];
    memcpy(cnt, p, 32);
    p += 8;
    uint64_t within = k & 0x7f;           // count symbols at block offsets [0, within]
    uint64_t full = (within + 1) >> 4;    // whole words
    for (uint64_t wd = 0; wd < full; ++wd) count_word(p[wd], 0x55555555u, cnt);
    int rem = (int)((within + 1) & 15);   // leading (MSB-first) symbols of the next word
    if (rem) count_word(p[full], 0x55555555u & (~0u << (32 - 2 * rem)), cnt);
}

uint64_t bwt_occ(const Index& ix, uint64_t k, int c) {
    if (k == ix.seq_len) return ix.L2[c + 1] - ix.L2[c];
    if (k == (uint64_t)-1) return 0;
    uint64_t cnt[4];
    bwt_occ4(ix, k, cnt);
    return cnt[c];
}

void bwt_set_intv(const Index& ix, int c, Intv& ik) {
    ik.x0 = ix.L2[c] + 1; ik.x1 = ix.L2[3 - c] + 1; ik.x2 = ix.L2[c + 1] - ix.L2[c]; ik.info = 0;
}

void bwt_extend(const Index& ix, const Intv& ik, Intv ok[4], int is_back, Counters* ctr) {
    if (ctr) ++ctr->n_extend;
    uint64_t tk[4], tl[4];
    const uint64_t* x = &ik.x0;
    int o = !is_back;
    bwt_occ4(ix, x[o] - 1, tk);
    bwt_occ4(ix, x[o] - 1 + ik.x2, tl);
    for (int i = 0; i < 4; ++i) {
        uint64_t* y = &ok[i].x0;
        y[o] = ix.L2[i] + 1 + tk[i];
        y[2] = tl[i] - tk[i];
    }
    (&ok[3].x0)[is_back] = x[is_back] + (x[o] <= ix.primary && x[o] + ik.x2 - 1 >= ix.primary);
    (&ok[2].x0)[is_back] = (&ok[3].x0)[is_back] + ok[3].x2;
    (&ok[1].x0)[is_back] = (&ok[2].x0)[is_back] + ok[2].x2;
    (&ok[0].x0)[is_back] = (&ok[1].x0)[is_back] + ok[1].x2;
}

uint64_t bwt_sa(const Index& ix, uint64_t k, Counters* ctr) {
    uint64_t sa = 0, mask = (uint64_t)ix.sa_intv - 1;
    while (k & mask) {
        ++sa;
        if (ctr) ++ctr->n_lf;
        if (k == ix.primary) k = 0;
        else {
            uint64_t kk = k < ix.primary ? k : k - 1;
            int c = ix.bwt[((kk >> 7) << 4) + 8 + ((kk & 0x7f) >> 4)] >> ((~kk & 0xf) << 1) & 3;
            k = ix.L2[c] + bwt_occ(ix, k, c);
        }
    }
    if (ctr) ++ctr->n_sa;
    return sa + ix.sa[k / ix.sa_intv];
}

uint64_t hash_64(uint64_t key) {
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

}  // namespace orc
