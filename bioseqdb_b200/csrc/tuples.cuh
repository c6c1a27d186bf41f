// tuples.cuh -- row materialisation kernels (tuples.cu): NUCLSEQ datum images of ref_subseq / query_subseq, CIGAR strings
// and ref_match_* integers for every row of a result.
#pragma once
#include "common.cuh"
#include "pipeline.cuh"

struct TupleHole { int64_t offset, end; uint32_t idx; int32_t amb; };   // a hole of the index: [offset, end), position in the hole list, letter

struct TupleParams {
    const RowPub* rows; uint64_t n_rows;
    const uint32_t* row_read;      // read index of every row
    const uint32_t* cigar;         // the result's CIGAR words
    const uint8_t* seqs;           // the reads as ASCII (the text BwaIndex::align_sequence works on), offs[n_reads + 1]
    const uint64_t* offs;
    const uint8_t* pac; int64_t l_pac; const int64_t* ann_offset;
    const TupleHole* holes; const int64_t* hole_maxend; uint32_t n_holes;   // sorted by offset; maxend[k] = max end of holes[0..k]
    uint32_t* nholes;              // 2 per row: holes of the ref image (bit 31: the row overlaps index holes), holes of the query image
    uint64_t* off;                 // 3 per row + 1: sizes, then (after the scan) byte offsets of ref image / query image / CIGAR string
    int32_t* ref_match;            // 3 per row: ref_match_begin, ref_match_end, ref_match_len
    uint8_t* bytes;
    bool fix_reverse;              // opt-in fix-up: reverse-strand hits are reported in forward-strand coordinates and text
};
size_t tuple_scan_tmp_elems(uint64_t n_rows);
void launch_tuple_sizes(const TupleParams& P, uint64_t* scan_tmp, cudaStream_t st, uint64_t* launches);
void launch_tuple_fill(const TupleParams& P, cudaStream_t st, uint64_t* launches);
