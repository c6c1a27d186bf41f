// tests/seed_thread_check.cpp -- TEST INFRASTRUCTURE: runs the thread-per-read seeding code of the product
// (bioseqdb_b200/csrc/seed_thread.cuh, compiled here for the HOST) on a CPU copy of the index and compares every read's interval
// list and the logical bwt_extend count with the oracle's mem_collect_intv (oracle/mem.cpp).  Reads the thread code declines
// (ambiguous bases, buffer limits) are reported as "fallback" -- on the GPU they go to the warp kernel -- and must be rare.
//   usage: seed_thread_check ref_len n_reads read_len sub_rate indel_rate repeat_copies K seed wide(0/1)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../oracle/oracle.h"
#include "../bioseqdb_b200/csrc/seed_thread.cuh"

static uint64_t rng_state = 1;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
static double rnd01() { return (rnd() >> 11) * (1.0 / 9007199254740992.0); }

template <class IdxT> static int run(int ref_len, int n_reads, int read_len, double sub, double indel, int copies, int K, int n_frac_reads) {
    using namespace orc;
    std::string ref((size_t)ref_len, 'A');
    for (auto& c : ref) c = "ACGT"[rnd() & 3];
    if (copies > 0) {   // repeat families: copies of a 100-700 bp unit with ~1 % divergence
        for (int f = 0; f < 4; ++f) {
            const int ulen = 100 + (int)(rnd() % 600);
            std::string u((size_t)ulen, 'A');
            for (auto& c : u) c = "ACGT"[rnd() & 3];
            for (int k = 0; k < copies; ++k) {
                const size_t p = rnd() % (size_t)(ref_len - ulen);
                for (int j = 0; j < ulen; ++j) ref[p + j] = rnd01() < 0.01 ? "ACGT"[rnd() & 3] : u[(size_t)j];
            }
        }
        // a low-complexity stretch: long lists in the backward phase
        const size_t p = rnd() % (size_t)(ref_len - 400);
        for (int j = 0; j < 300; ++j) ref[p + j] = "AC"[j & 1];
    }
    Opts opt; opts_init(opt);
    Index ix;
    // two rows so that the byte-rounding filler and a row boundary exist
    Nuclseq s1, s2; char bad;
    const size_t cut = (size_t)ref_len / 3 * 2 + 1;
    nuclseq_from_text(ref.substr(0, cut), s1, &bad); nuclseq_from_text(ref.substr(cut), s2, &bad);
    ix.add_ref(1, s1); ix.add_ref(2, s2);
    ix.build();
    const int64_t l_pac = ix.l_pac, n = (int64_t)ix.seq_len;
    // text, full SA, inverse SA
    std::vector<uint8_t> T((size_t)n);
    for (int64_t i = 0; i < l_pac; ++i) { T[(size_t)i] = pac_get(ix.pac.data(), (uint64_t)i); T[(size_t)(n - 1 - i)] = 3 - T[(size_t)i]; }
    std::vector<int64_t> sa64; suffix_array(T.data(), n, sa64);
    std::vector<IdxT> sa((size_t)n + 1), isa((size_t)n + 1);
    for (int64_t r = 0; r <= n; ++r) { sa[(size_t)r] = (IdxT)sa64[(size_t)r]; isa[(size_t)sa64[(size_t)r]] = (IdxT)r; }
    // prefix table, level by level (as k_kmer_level builds it on the device)
    std::vector<seedt::U4> tab(seedt::level_off(K + 1));
    auto entry = [](uint64_t x0, uint64_t x1, uint32_t x2) { seedt::U4 e; e.x = (uint32_t)x0; e.y = (uint32_t)x1; e.z = x2; e.w = (uint32_t)((x0 >> 32) & 0xff) | (uint32_t)((x1 >> 32) & 0xff) << 8; return e; };
    for (int c = 0; c < 4; ++c) { Intv ik; bwt_set_intv(ix, c, ik); tab[seedt::level_off(1) + (uint32_t)c] = entry(ik.x0, ik.x1, (uint32_t)ik.x2); }
    for (int t = 2; t <= K; ++t) {
        const uint32_t np = 1u << (2 * (t - 1));
        for (uint32_t p = 0; p < np; ++p) {
            const seedt::U4 pe = tab[seedt::level_off(t - 1) + p];
            Intv ik; ik.x0 = seedt::tab_x0<uint64_t>(pe); ik.x1 = seedt::tab_x1<uint64_t>(pe); ik.x2 = pe.z; ik.info = 0;
            Intv ok[4];
            bwt_extend(ix, ik, ok, 0, nullptr);
            for (int b = 0; b < 4; ++b) tab[seedt::level_off(t) + p * 4 + (uint32_t)b] = entry(ok[3 - b].x0, ok[3 - b].x1, pe.z ? (uint32_t)ok[3 - b].x2 : 0u);
        }
    }
    std::vector<uint8_t> pac(ix.pac); pac.resize(pac.size() + 16, 0);
    std::vector<uint32_t> occ(ix.bwt); occ.resize(occ.size() + 32, 0);
    seedt::Index<IdxT> X{};
    std::vector<uint32_t> ztab(tab.size());
    for (size_t i = 0; i < tab.size(); ++i) ztab[i] = tab[i].z;
    X.occ = occ.data(); X.tab = tab.data(); X.kk = K; X.ztab = (K & 1) ? ztab.data() : nullptr;   // odd depths run with the size table, even ones without
    X.sa = sa.data(); X.isa = isa.data(); X.pac = pac.data();
    X.l_pac = (IdxT)l_pac; X.n = (IdxT)n; X.primary = (IdxT)ix.primary;
    for (int c = 0; c < 5; ++c) X.L2[c] = (IdxT)ix.L2[c];
    seedt::Opts so; so.min_seed_len = opt.min_seed_len; so.split_len = (int)(opt.min_seed_len * opt.split_factor + .499); so.split_width = opt.split_width; so.max_mem_intv = opt.max_mem_intv;

    int bad_reads = 0, fallbacks = 0;
    unsigned long long ext_thread = 0, ext_oracle = 0;
    std::vector<uint8_t> q;
    std::vector<seedt::IntvOut> out(4096);
    for (int r = 0; r < n_reads; ++r) {
        // simulate: a window of the reference (either strand) with substitutions and indels; some reads are random or periodic
        q.clear();
        int rl = read_len;
        if (r % 97 == 0) rl = 19 + (int)(rnd() % 40);             // short reads around min_seed_len
        const int kind = r % 50;
        if (kind == 1) { for (int j = 0; j < rl; ++j) q.push_back((uint8_t)(rnd() & 3)); }
        else if (kind == 2) { for (int j = 0; j < rl; ++j) q.push_back((uint8_t)("\0\1"[j & 1])); }
        else {
            size_t p = rnd() % (size_t)(ref_len - rl - 8);
            const bool rev = rnd() & 1;
            for (size_t j = p; (int)q.size() < rl && j < (size_t)ref_len; ++j) {
                const double u = rnd01();
                if (u < indel) continue;                          // deletion
                if (u < 2 * indel) q.push_back((uint8_t)(rnd() & 3));   // insertion
                uint8_t b = (uint8_t)nt4(ref[j]);
                if (rnd01() < sub) b = (uint8_t)((b + 1 + rnd() % 3) & 3);
                q.push_back(b);
            }
            q.resize((size_t)rl, 0);
            if (rev) { std::reverse(q.begin(), q.end()); for (auto& b : q) b = 3 - b; }
        }
        if (n_frac_reads && r % n_frac_reads == 3) q[(size_t)(rnd() % q.size())] = 4;     // an ambiguous base: the thread path must decline
        const int len = (int)q.size();
        // oracle
        std::vector<Intv> mem; Counters ctr;
        collect_intv(opt, ix, len, q.data(), mem, &ctr);
        // thread code
        bool has_n = false;
        std::vector<uint32_t> pk((size_t)(len >> 4) + 3, 0);
        for (int j = 0; j < len; ++j) { if (q[(size_t)j] > 3) has_n = true; pk[(size_t)j >> 4] |= (uint32_t)(q[(size_t)j] & 3) << (30 - ((j & 15) << 1)); }
        if (has_n || len < so.min_seed_len) { ++fallbacks; continue; }
        seedt::Read R; R.pk = pk.data(); R.stride = 1; R.len = len;
        uint32_t pcbuf[seedt::PCAP];
        seedt::Work<IdxT> W; seedt::work_init(W, out.data(), 42u, pcbuf, 1);
        seedt::collect(X, so, R, W);
        if (W.fail) { ++fallbacks; continue; }
        std::sort(out.begin(), out.begin() + W.n_out, [](const seedt::IntvOut& a, const seedt::IntvOut& b) { return a.info < b.info; });
        bool same = W.n_out == mem.size() && W.n_ext == ctr.n_extend;
        for (size_t k = 0; same && k < mem.size(); ++k)
            same = out[k].x0 == mem[k].x0 && out[k].x1 == mem[k].x1 && out[k].x2 == mem[k].x2 && out[k].info == mem[k].info;
        ext_thread += W.n_ext; ext_oracle += ctr.n_extend;
        if (!same) {
            if (bad_reads < 5) {
                fprintf(stderr, "read %d (len %d): thread %u intervals / %llu extends, oracle %zu / %llu\n", r, len, W.n_out, W.n_ext, mem.size(), (unsigned long long)ctr.n_extend);
                for (size_t k = 0; k < std::max<size_t>(W.n_out, mem.size()) && k < 12; ++k) {
                    if (k < W.n_out) fprintf(stderr, "   T %llu %llu %llu [%d,%d)", (unsigned long long)out[k].x0, (unsigned long long)out[k].x1, (unsigned long long)out[k].x2, (int)(out[k].info >> 32), (int)(uint32_t)out[k].info);
                    if (k < mem.size()) fprintf(stderr, "   O %llu %llu %llu [%d,%d)", (unsigned long long)mem[k].x0, (unsigned long long)mem[k].x1, (unsigned long long)mem[k].x2, (int)(mem[k].info >> 32), (int)(uint32_t)mem[k].info);
                    fprintf(stderr, "\n");
                }
            }
            ++bad_reads;
        }
    }
    printf("reads %d mismatching %d fallback %d extends thread %llu oracle %llu\n", n_reads, bad_reads, fallbacks, ext_thread, ext_oracle);
    return bad_reads ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc < 10) { fprintf(stderr, "usage: %s ref_len n_reads read_len sub indel repeat_copies K seed wide\n", argv[0]); return 2; }
    const int ref_len = atoi(argv[1]), n_reads = atoi(argv[2]), read_len = atoi(argv[3]);
    const double sub = atof(argv[4]), indel = atof(argv[5]);
    const int copies = atoi(argv[6]), K = atoi(argv[7]);
    rng_state = (uint64_t)atoll(argv[8]);
    const int wide = atoi(argv[9]);
    return wide ? run<uint64_t>(ref_len, n_reads, read_len, sub, indel, copies, K, 40) : run<uint32_t>(ref_len, n_reads, read_len, sub, indel, copies, K, 40);
}
