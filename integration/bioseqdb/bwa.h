// integration/bioseqdb/bwa.h -- DROP-IN replacement for reference bioseqdb/bwa.h (same class surface, lines 15-48): BwaMatch and
// BwaIndex{BwaIndex(), ~BwaIndex(), align_sequence() const, build(), add_ref_sequence(), mem_opt_t* options}.  extension.cpp,
// sequence.h and sequence.cpp of the reference compile against it UNCHANGED (with integration/include on the include path in place
// of libbwa's headers): bwa_index_from_query returns a BwaIndex by value and writes bwa.options->... (extension.cpp:211-236), the two
// SQL functions call bwa.align_sequence(*nucls) (extension.cpp:336,363).  Everything below the class is libbioseqdb_gpu's C ABI.
#pragma once

#include <cstdint>
#include <string>
#include <string_view>
#include <vector>

extern "C" {
#include <bwa/bwt.h>
#include <bwa/bwamem.h>
}

#include "sequence.h"

struct bsq_index;

struct BwaMatch {
    int64_t ref_id;
    std::string ref_subseq;
    int32_t ref_match_begin;
    int32_t ref_match_end;
    int32_t ref_match_len;
    std::string_view query_subseq;
    int32_t query_match_begin;
    int32_t query_match_end;
    int32_t query_match_len;
    bool is_primary;
    bool is_secondary;
    bool is_reverse;
    std::string cigar;
    int score;
};

class BwaIndex {
public:
    explicit BwaIndex();
    ~BwaIndex();
    // the reference returns a BwaIndex by value from bwa_index_from_query (extension.cpp:211-236): movable, not copyable
    BwaIndex(BwaIndex&& other) noexcept;
    BwaIndex& operator=(BwaIndex&& other) noexcept;
    BwaIndex(const BwaIndex&) = delete;
    BwaIndex& operator=(const BwaIndex&) = delete;

    std::vector<BwaMatch> align_sequence(const NucleotideSequence& seq) const;
    // the whole per-read loop of nuclseq_multi_search_bwa (extension.cpp:362-370) in one GPU pass; a maintainer who collects the
    // cursor first calls this instead of align_sequence per row
    std::vector<std::vector<BwaMatch>> align_sequences(const std::vector<const NucleotideSequence*>& seqs) const;
    void build();
    void add_ref_sequence(int64_t id, const NucleotideSequence& seq);

    mem_opt_t* options;

private:
    std::vector<ubyte_t> pac_forward;
    std::vector<bntamb1_t> holes;
    std::vector<int64_t> offsets;
    bsq_index* index;
    void release() noexcept;
};
