// tuples.cu -- row materialisation: the variable-length columns of the 15-column bwa_result tuple,
// built on the GPU for every row of a result.  Replaces, per row, reference bioseqdb/bwa.cpp:55-68
// (extract_reference_subseq), bwa.cpp:70-77 (cigar_compressed_to_string), bwa.cpp:171-173 (ref_match_*) and the two
// nuclseq_from_text calls of bioseqdb/extension.cpp:285,290 (sequence.cpp:209-245): ref_subseq and query_subseq leave the
// device as finished NUCLSEQ datum images -- 12-byte header {varlena length word, holes_num, len}, holes_num hole records
// of 16 bytes, ceil(len / 4) bytes of 2-bit codes MSB first with the bits under holes and in the tail padding drawn from
// std::minstd_rand(holes_num ^ len) in text order -- so the PostgreSQL shim copies bytes instead of formatting text and
// parsing it back twice per row.
//
// Reference quirks kept: hole offsets of the index are NOT rebased per row, later holes overwrite earlier
// ones; the CIGAR letters are htslib's table applied to bwa's op codes (soft clip prints 'N'); ref_match_begin/end wrap to
// int32.  A reverse-strand hit's ref_subseq is the reverse-strand text (the reference reads out of bounds there, B#3).
#include "pipeline.cuh"
#include "primitives.cuh"
#include "tuples.cuh"

namespace {

constexpr int TUP_THREADS = 128;

__device__ __forceinline__ int nt4_of(uint8_t c) {   // nst_nt4_table as used by nuclcode_from_char (sequence.h:51-53)
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case '-': return 5;
        default: return 4;
    }
}

// letter of the LAST hole (in index order) that covers pos, 0 when none.  holes are sorted by offset; maxend[k] = the
// largest end among holes[0..k], which bounds the walk back from the first hole starting beyond pos.
__device__ __forceinline__ int amb_letter(const TupleParams& P, int64_t pos) {
    uint32_t lo = 0, hi = P.n_holes;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (P.holes[mid].offset > pos) hi = mid; else lo = mid + 1; }
    int best = -1, letter = 0;
    for (int64_t k = (int64_t)lo - 1; k >= 0 && P.hole_maxend[k] > pos; --k) {
        const TupleHole h = P.holes[k];
        if (h.end > pos && (int)h.idx > best) { best = (int)h.idx; letter = h.amb; }
    }
    return letter;
}
__device__ __forceinline__ bool range_has_holes(const TupleParams& P, int64_t rb, int64_t re) {
    if (P.n_holes == 0 || re <= rb) return false;
    uint32_t lo = 0, hi = P.n_holes;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (P.holes[mid].offset >= re) hi = mid; else lo = mid + 1; }
    return lo > 0 && P.hole_maxend[lo - 1] > rb;
}

// character i of the two texts
struct RefText {
    const TupleParams& P; int64_t rb; bool overlay;
    __device__ __forceinline__ uint8_t operator()(int64_t i) const {
        const int64_t p = rb + i;
        if (overlay) { const int a = amb_letter(P, p); if (a) return (uint8_t)a; }
        const uint32_t b = p < P.l_pac ? pac_get(P.pac, p) : 3u - pac_get(P.pac, (P.l_pac << 1) - 1 - p);
        return (uint8_t)"ACGT"[b];
    }
};
struct QueryText {
    const uint8_t* q;
    __device__ __forceinline__ uint8_t operator()(int64_t i) const { return q[i]; }
};

template <class Text> __device__ __forceinline__ uint32_t count_holes(const Text& T, int64_t len) {   // calculate_num_of_holes, sequence.cpp:46-57
    uint32_t n = 0; uint8_t prev = 0;
    for (int64_t i = 0; i < len; ++i) { const uint8_t c = T(i); if (c != prev && nt4_of(c) >= 4) ++n; prev = c; }
    return n;
}

__device__ __forceinline__ uint64_t image_bytes(uint32_t holes, int64_t len) { return 12ull + 16ull * holes + (uint64_t)((len + 3) >> 2); }
__device__ __forceinline__ uint64_t pad8(uint64_t x) { return (x + 7) & ~7ull; }

__device__ __forceinline__ int dec_digits(uint32_t v) { int d = 1; while (v >= 10) { v /= 10; ++d; } return d; }

// The reference interval of a row.  Reference behaviour: [rb, re) of the doubled coordinate as it is.  With the reverse-strand
// fix-up a hit on the reverse strand becomes the forward-strand interval it covers, [2 l_pac - re, 2 l_pac - rb).
__device__ __forceinline__ RowPub row_for_text(const TupleParams& P, RowPub a) {
    if (P.fix_reverse && a.rb >= P.l_pac) { const int64_t b = (P.l_pac << 1) - a.re, e = (P.l_pac << 1) - a.rb; a.rb = b; a.re = e; }
    return a;
}

// pass 1: sizes of the three byte strings of every row, hole counts, the ref_match_* integers
__global__ void __launch_bounds__(TUP_THREADS) k_tuple_sizes(TupleParams P) {
    const uint64_t row = (uint64_t)blockIdx.x * TUP_THREADS + threadIdx.x;
    if (row >= P.n_rows) return;
    const RowPub a = row_for_text(P, P.rows[row]);
    const int64_t rlen = a.re > a.rb ? a.re - a.rb : 0;
    const bool overlay = range_has_holes(P, a.rb, a.re);
    uint32_t nh_ref = 0;
    if (overlay) { RefText T{P, a.rb, true}; nh_ref = count_holes(T, rlen); }
    const uint32_t rd = P.row_read[row];
    const int64_t qlen = a.qe > a.qb ? a.qe - a.qb : 0;
    QueryText Q{P.seqs + P.offs[rd] + a.qb};
    const uint32_t nh_q = count_holes(Q, qlen);
    uint64_t cig = 0;
    for (uint32_t k = 0; k < a.n_cigar; ++k) cig += (uint64_t)dec_digits(P.cigar[a.cigar_off + k] >> 4) + 1;
    P.nholes[2 * row] = nh_ref | (overlay ? 0x80000000u : 0u);
    P.nholes[2 * row + 1] = nh_q;
    P.off[3 * row] = pad8(image_bytes(nh_ref, rlen));
    P.off[3 * row + 1] = pad8(image_bytes(nh_q, qlen));
    P.off[3 * row + 2] = pad8(cig + 1);      // NUL-terminated (the block is zero-filled), padded so that the next image stays 8-byte aligned
    const int64_t ref_offset = P.ann_offset[a.rid];
    P.ref_match[3 * row] = (int32_t)(uint32_t)(uint64_t)(a.rb - ref_offset);       // bwa.cpp:171-173 store into int32 fields
    P.ref_match[3 * row + 1] = (int32_t)(uint32_t)(uint64_t)(a.re - ref_offset);
    P.ref_match[3 * row + 2] = (int32_t)(uint32_t)(uint64_t)(a.re - a.rb);
}

// nuclseq_from_text (sequence.cpp:209-245) straight into the datum image at dst (8-byte aligned, zero-filled by the caller)
template <class Text>
__device__ __forceinline__ void write_image(uint8_t* dst, const Text& T, int64_t len, uint32_t holes_num) {
    const uint64_t size = image_bytes(holes_num, len);
    uint32_t* hdr = reinterpret_cast<uint32_t*>(dst);
    hdr[0] = (uint32_t)size << 2;     // SET_VARSIZE of an uncompressed 4-byte varlena header (little endian)
    hdr[1] = holes_num; hdr[2] = (uint32_t)len;
    uint8_t* holes = dst + 12;
    uint8_t* pac = holes + 16ull * holes_num;
    // std::minstd_rand(holes_num ^ len): x <- 48271 x mod (2^31 - 1), a zero seed becomes 1
    uint64_t x = (uint64_t)((uint32_t)holes_num ^ (uint32_t)len) % 2147483647ull;
    if (x == 0) x = 1;
    int64_t hole_i = -1, h_off = 0; int32_t h_len = 0; uint8_t h_amb = 0, prev = 0;
    auto flush = [&]() {
        if (hole_i < 0) return;
        uint8_t* r = holes + 16 * hole_i;
        *reinterpret_cast<uint32_t*>(r) = (uint32_t)h_off; *reinterpret_cast<uint32_t*>(r + 4) = (uint32_t)((uint64_t)h_off >> 32);
        *reinterpret_cast<int32_t*>(r + 8) = h_len; r[12] = h_amb; r[13] = r[14] = r[15] = 0;
    };
    uint32_t cur = 0;
    const int64_t padded = ((len + 3) >> 2) << 2;
    for (int64_t i = 0; i < padded; ++i) {
        uint32_t code;
        if (i < len) {
            const uint8_t c = T(i);
            const int k = nt4_of(c);
            if (k >= 4) {
                if (prev == c) ++h_len;
                else { flush(); ++hole_i; h_amb = c; h_off = i; h_len = 1; }
                x = x * 48271ull % 2147483647ull;
                code = (uint32_t)x & 3u;
            } else code = (uint32_t)k;
            prev = c;
        } else {
            x = x * 48271ull % 2147483647ull;
            code = (uint32_t)x & 3u;
        }
        cur = cur << 2 | code;
        if ((i & 3) == 3) { pac[i >> 2] = (uint8_t)cur; cur = 0; }
    }
    flush();
}

// pass 2: one thread per (row, column)
__global__ void __launch_bounds__(TUP_THREADS) k_tuple_fill(TupleParams P) {
    const uint64_t t = (uint64_t)blockIdx.x * TUP_THREADS + threadIdx.x;
    if (t >= 3 * P.n_rows) return;
    const uint64_t row = t / 3; const int col = (int)(t - row * 3);
    const RowPub a = row_for_text(P, P.rows[row]);
    uint8_t* dst = P.bytes + P.off[t];
    if (col == 0) {
        const uint32_t nh = P.nholes[2 * row];
        RefText T{P, a.rb, (nh & 0x80000000u) != 0};
        write_image(dst, T, a.re > a.rb ? a.re - a.rb : 0, nh & 0x7fffffffu);
    } else if (col == 1) {
        QueryText Q{P.seqs + P.offs[P.row_read[row]] + a.qb};
        write_image(dst, Q, a.qe > a.qb ? a.qe - a.qb : 0, P.nholes[2 * row + 1]);
    } else {
        for (uint32_t k = 0; k < a.n_cigar; ++k) {
            const uint32_t w = P.cigar[a.cigar_off + k];
            uint32_t v = w >> 4;
            const int d = dec_digits(v);
            for (int j = d - 1; j >= 0; --j) { dst[j] = (uint8_t)('0' + v % 10); v /= 10; }
            dst[d] = (uint8_t)"MIDNSHP=XB??????"[w & 0xf];   // htslib's BAM_CIGAR_STR indexed by bwa's op code (bwa.cpp:70-77)
            dst += d + 1;
        }
    }
}

}  // namespace

size_t tuple_scan_tmp_elems(uint64_t n_rows) { return prim::scan_tmp_elems(3 * n_rows + 1) + 16; }

// off[] holds 3 n_rows sizes (+ one trailing slot) on entry of the scan and the byte offsets afterwards
void launch_tuple_sizes(const TupleParams& P, uint64_t* scan_tmp, cudaStream_t st, uint64_t* launches) {
    if (P.n_rows == 0) return;
    k_tuple_sizes<<<(unsigned)((P.n_rows + TUP_THREADS - 1) / TUP_THREADS), TUP_THREADS, 0, st>>>(P);
    if (launches) ++*launches;
    prim::device_scan<uint64_t, prim::OpSum, false>(P.off, P.off, (size_t)(3 * P.n_rows + 1), scan_tmp, prim::OpSum(), st, launches);
}
void launch_tuple_fill(const TupleParams& P, cudaStream_t st, uint64_t* launches) {
    if (P.n_rows == 0) return;
    k_tuple_fill<<<(unsigned)((3 * P.n_rows + TUP_THREADS - 1) / TUP_THREADS), TUP_THREADS, 0, st>>>(P);
    if (launches) ++*launches;
}
