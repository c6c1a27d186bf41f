// sequence.h -- host-side packed nucleotide type, PostgreSQL-free mirror of reference
// bioseqdb/sequence.h:18-61.  Same payload bytes as the NUCLSEQ varlena after its 4-byte length word:
// {u32 holes_num, u32 len, holes[holes_num] (16-byte bntamb1_t records), pac[ceil(len/4)]}.
#pragma once
#include <cstdint>
#include <string>
#include <string_view>
#include <vector>
#include "../../include/bioseqdb_gpu.h"

namespace bioseqdb {

constexpr std::string_view allowed_nucleotides = "ACGTNWSMKRYBDHV";

struct NucleotideSequence {
    uint32_t len = 0;
    std::vector<bsq_hole> holes_;
    std::vector<uint8_t> pac_;

    size_t length() const { return len; }
    uint32_t holes_num() const { return (uint32_t)holes_.size(); }
    const bsq_hole* holes() const { return holes_.data(); }
    const uint8_t* pac() const { return pac_.data(); }
    std::string to_text() const;                        // to_text_palloc (sequence.cpp:162-166)
    std::vector<uint8_t> varlena_payload() const;       // bytes after vl_len in the PG datum
};

// nuclseq_in's validation (extension.cpp:46-60) + nuclseq_from_text (sequence.cpp:209-245).
// Throws std::invalid_argument("invalid nucleotide in nuclseq_in: 'x'") on a bad letter.
NucleotideSequence nuclseq_from_text(std::string_view text);

static inline size_t pac_byte_size(size_t x) { return x / 4 + (x % 4 != 0 ? 1 : 0); }
static inline uint8_t pac_raw_get(const uint8_t* pac, size_t index) { return pac[index >> 2] >> ((~index & 3) << 1) & 3; }
static inline void pac_raw_set(uint8_t* pac, size_t index, uint8_t value) { pac[index >> 2] |= value << ((~index & 3) << 1); }
int nuclcode_from_char(char chr);   // nst_nt4_table

}  // namespace bioseqdb
