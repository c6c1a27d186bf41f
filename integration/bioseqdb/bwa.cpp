// integration/bioseqdb/bwa.cpp -- DROP-IN replacement for reference bioseqdb/bwa.cpp: the same adapter (add_ref_sequence, build,
// align_sequence, extract_reference_subseq, cigar_compressed_to_string, reference lines 55-181) with every libbwa / htslib call
// replaced by the C ABI of libbioseqdb_gpu (include/bioseqdb_gpu.h).  Errors of the GPU library surface as ereport(ERROR) when built
// inside PostgreSQL (BIOSEQDB_HAVE_POSTGRES) and as std::runtime_error otherwise.
#include "bwa.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "bioseqdb_gpu.h"

// nst_nt4_table of libbwa's bntseq.c: A/a 0, C/c 1, G/g 2, T/t 3, '-' 5, everything else 4
unsigned char nst_nt4_table[256] = {
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 5 /*'-'*/, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  3, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  3, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,   4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4
};

namespace {

[[noreturn]] void fail() {
#ifdef BIOSEQDB_HAVE_POSTGRES
    ereport(ERROR, (errcode(ERRCODE_EXTERNAL_ROUTINE_EXCEPTION), errmsg("bioseqdb gpu: %s", bsq_last_error())));
    abort();   // not reached: ereport(ERROR) does not return
#else
    throw std::runtime_error(bsq_last_error());
#endif
}

bsq_opts to_bsq(const mem_opt_t& o) {
    bsq_opts b;
    b.min_seed_len = o.min_seed_len; b.max_occ = o.max_occ; b.a = o.a; b.b = o.b; b.pen_clip3 = o.pen_clip3; b.pen_clip5 = o.pen_clip5;
    b.zdrop = o.zdrop; b.w = o.w; b.o_del = o.o_del; b.e_del = o.e_del; b.o_ins = o.o_ins; b.e_ins = o.e_ins;
    return b;
}

std::string cigar_compressed_to_string(const uint32_t* raw, int len) {
    std::string cigar;
    for (int i = 0; i < len; i++) {
        cigar += std::to_string(raw[i] >> 4);      // bam_cigar_oplen
        cigar += "MIDNSHP=XB"[raw[i] & 0xf];       // bam_cigar_opchr on bwa's op codes: a soft clip (3) prints as 'N' (reference bwa.cpp:70-77)
    }
    return cigar;
}

// glibc lrand48() from a fresh process (state 0), one draw per aligned read: mem_align1's `id` (the backend's session stream)
uint64_t g_lrand48_state = 0;
int64_t next_lrand48() {
    g_lrand48_state = (g_lrand48_state * 0x5DEECE66DULL + 0xBULL) & 0xFFFFFFFFFFFFULL;
    return (int64_t)(g_lrand48_state >> 17);
}

}  // namespace

BwaIndex::BwaIndex() : options(new mem_opt_t), index(nullptr) {
    bsq_opts d; bsq_opts_init(&d);                  // mem_opt_init() defaults (reference bwa.cpp:80)
    options->a = d.a; options->b = d.b; options->o_del = d.o_del; options->e_del = d.e_del; options->o_ins = d.o_ins; options->e_ins = d.e_ins;
    options->pen_clip5 = d.pen_clip5; options->pen_clip3 = d.pen_clip3; options->w = d.w; options->zdrop = d.zdrop;
    options->min_seed_len = d.min_seed_len; options->max_occ = d.max_occ;
    const char* dev = getenv("BIOSEQDB_GPU_DEVICE");
    index = bsq_index_new(&d, dev ? atoi(dev) : 0);   // the CUDA context is created here, i.e. inside the backend, after fork
    if (!index) { delete options; options = nullptr; fail(); }
}

void BwaIndex::release() noexcept {
    if (index) bsq_index_free(index);
    delete options;
    index = nullptr; options = nullptr;
}

BwaIndex::~BwaIndex() { release(); }

BwaIndex::BwaIndex(BwaIndex&& o) noexcept
    : options(o.options), pac_forward(std::move(o.pac_forward)), holes(std::move(o.holes)), offsets(std::move(o.offsets)), index(o.index) {
    o.options = nullptr; o.index = nullptr;
}

BwaIndex& BwaIndex::operator=(BwaIndex&& o) noexcept {
    if (this != &o) {
        release();
        options = o.options; index = o.index; pac_forward = std::move(o.pac_forward); holes = std::move(o.holes); offsets = std::move(o.offsets);
        o.options = nullptr; o.index = nullptr;
    }
    return *this;
}

void BwaIndex::add_ref_sequence(int64_t id, const NucleotideSequence& seq) {
    // reference bwa.cpp:82-105: byte-rounded concatenation, holes copied without rebasing
    offsets.push_back((int64_t)pac_forward.size() * 4);
    pac_forward.insert(pac_forward.end(), seq.pac(), seq.pac() + pac_byte_size(seq.len));
    holes.insert(holes.end(), seq.holes(), seq.holes() + seq.holes_num);
    static_assert(sizeof(bntamb1_t) == sizeof(bsq_hole), "hole record layout");
    if (bsq_index_add_ref(index, id, seq.pac(), seq.len, reinterpret_cast<const bsq_hole*>(seq.holes()), seq.holes_num) != BSQ_OK) fail();
}

void BwaIndex::build() {
    if (pac_forward.empty()) return;                  // reference bwa.cpp:108-109
    const bsq_opts o = to_bsq(*options);
    if (bsq_index_set_opts(index, &o) != BSQ_OK) fail();
    if (bsq_index_build(index) != BSQ_OK) fail();
}

static std::string extract_reference_subseq(const std::vector<ubyte_t>& pac_forward, const std::vector<bntamb1_t>& holes, int64_t rb, int64_t re) {
    // reference bwa.cpp:55-68.  Forward hits: the same arithmetic incl. the un-rebased hole overlay.  Reverse hits index past
    // pac_forward in the reference (undefined behaviour, bwa.cpp:156 TODO): defined here as the reverse-strand text.
    const int64_t l_pac = (int64_t)pac_forward.size() * 4;
    std::string subseq((size_t)(re - rb), '?');
    for (int64_t i = 0; i < re - rb; ++i) {
        const int64_t p = rb + i;
        const int c = p < l_pac ? pac_raw_get(pac_forward.data(), (size_t)p) : 3 - pac_raw_get(pac_forward.data(), (size_t)((l_pac << 1) - 1 - p));
        subseq[(size_t)i] = "ACGT"[c];
    }
    for (const bntamb1_t& h : holes) {
        const int64_t l = std::max<int64_t>(h.offset, rb), r = std::min<int64_t>(h.offset + h.len, re);
        for (int64_t i = l; i < r; ++i) subseq[(size_t)(i - rb)] = h.amb;
    }
    return subseq;
}

std::vector<std::vector<BwaMatch>> BwaIndex::align_sequences(const std::vector<const NucleotideSequence*>& seqs) const {
    std::vector<std::vector<BwaMatch>> out(seqs.size());
    if (pac_forward.empty() || seqs.empty()) return out;       // reference bwa.cpp:142-143
    // options may have been written after build() (extension.cpp writes them before; this keeps the two in step either way)
    const bsq_opts o = to_bsq(*options);
    if (bsq_index_set_opts(index, &o) != BSQ_OK) fail();
    std::vector<const char*> texts(seqs.size());
    std::string cat;
    std::vector<uint64_t> offs(seqs.size() + 1, 0);
    std::vector<int64_t> ids(seqs.size());
    for (size_t i = 0; i < seqs.size(); ++i) {
        texts[i] = seqs[i]->to_text_palloc();                    // reference bwa.cpp:146: lives until the memory context resets
        cat.append(texts[i], seqs[i]->len);
        offs[i + 1] = cat.size();
        ids[i] = next_lrand48();
    }
    bsq_result* res = nullptr;
    if (bsq_align_batch(index, cat.data(), offs.data(), ids.data(), seqs.size(), &res) != BSQ_OK) fail();
    for (size_t i = 0; i < seqs.size(); ++i) {
        for (uint64_t k = res->row_off[i]; k < res->row_off[i + 1]; ++k) {
            const bsq_row& a = res->rows[k];
            const int64_t ref_offset = offsets[(size_t)a.rid];
            out[i].push_back(BwaMatch{
                /* ref_id */ a.ref_id,
                /* ref_subseq */ extract_reference_subseq(pac_forward, holes, a.rb, a.re),
                /* ref_match_begin */ static_cast<int32_t>(a.rb - ref_offset),
                /* ref_match_end */ static_cast<int32_t>(a.re - ref_offset),
                /* ref_match_len */ static_cast<int32_t>(a.re - a.rb),
                /* query_subseq */ std::string_view(texts[i] + a.qb, (size_t)(a.qe - a.qb)),
                /* query_match_begin */ a.qb,
                /* query_match_end */ a.qe,
                /* query_match_len */ a.qe - a.qb,
                /* is_primary */ (a.flag & 0x100) == 0,        // BAM_FSECONDARY
                /* is_secondary */ (a.flag & 0x100) != 0,
                /* is_reverse */ a.is_rev != 0,
                /* cigar */ cigar_compressed_to_string(res->cigar + a.cigar_off, (int)a.n_cigar),
                /* score */ a.score,
            });
        }
    }
    bsq_result_free(res);
    return out;
}

std::vector<BwaMatch> BwaIndex::align_sequence(const NucleotideSequence& seq) const {
    if (pac_forward.empty()) return {};
    return std::move(align_sequences({&seq})[0]);
}
