// sequence.cpp -- see sequence.h.  Mirrors reference bioseqdb/sequence.cpp:46-81,209-245.
#include "sequence.h"
#include <algorithm>
#include <cstring>
#include <random>
#include <stdexcept>

namespace bioseqdb {

int nuclcode_from_char(char chr) {
    switch (chr) {
        case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3;
        case '-': return 5; default: return 4;
    }
}

NucleotideSequence nuclseq_from_text(std::string_view str) {
    if (str.length() > INT32_MAX / 4) throw std::invalid_argument("provided sequence is too long");
    for (char chr : str)
        if (allowed_nucleotides.find(chr) == std::string_view::npos || chr == 0)
            throw std::invalid_argument(std::string("invalid nucleotide in nuclseq_in: '") + chr + "'");
    NucleotideSequence s;
    s.len = (uint32_t)str.size();
    uint32_t holes_num = 0;
    {
        char prev = 0;
        for (char chr : str) { if (prev != chr && nuclcode_from_char(chr) >= 4) ++holes_num; prev = chr; }
    }
    s.holes_.reserve(holes_num);
    s.pac_.assign(pac_byte_size(s.len), 0);
    std::minstd_rand rng(holes_num ^ (uint32_t)str.size());   // deterministic filler under holes and in the tail
    char prev = 0;
    for (uint32_t idx = 0; idx < str.size(); ++idx) {
        const char chr = str[idx];
        const int code = nuclcode_from_char(chr);
        if (code >= 4) {
            if (prev == chr) s.holes_.back().len++;
            else { bsq_hole h; memset(&h, 0, sizeof(h)); h.offset = idx; h.len = 1; h.amb = chr; s.holes_.push_back(h); }
            pac_raw_set(s.pac_.data(), idx, rng() & 3);
        } else pac_raw_set(s.pac_.data(), idx, (uint8_t)code);
        prev = chr;
    }
    for (size_t i = s.len; i < s.pac_.size() * 4; ++i) pac_raw_set(s.pac_.data(), i, rng() & 3);
    return s;
}

std::string NucleotideSequence::to_text() const {
    std::string text(len, '?');
    for (uint32_t i = 0; i < len; ++i) text[i] = "ACGT"[pac_raw_get(pac_.data(), i)];
    for (const bsq_hole& h : holes_) std::fill(text.begin() + h.offset, text.begin() + h.offset + h.len, h.amb);
    return text;
}

std::vector<uint8_t> NucleotideSequence::varlena_payload() const {
    std::vector<uint8_t> out(8 + holes_.size() * sizeof(bsq_hole) + pac_.size(), 0);
    uint32_t hn = holes_num();
    memcpy(out.data(), &hn, 4); memcpy(out.data() + 4, &len, 4);
    if (!holes_.empty()) memcpy(out.data() + 8, holes_.data(), holes_.size() * sizeof(bsq_hole));
    if (!pac_.empty()) memcpy(out.data() + 8 + holes_.size() * sizeof(bsq_hole), pac_.data(), pac_.size());
    return out;
}

}  // namespace bioseqdb
