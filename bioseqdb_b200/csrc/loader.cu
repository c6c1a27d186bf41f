// loader.cu -- bulk text -> NUCLSEQ conversion: nuclseq_in + nuclseq_from_text (reference
// bioseqdb/extension.cpp:46-60, bioseqdb/sequence.cpp:46-57,209-245) for a whole batch of sequences at once, as the bulk loader
// that replaces bioseqdb-import's one-INSERT-per-record loop (bioseqdb-import/main.cpp:52-72) needs it.  Input: the sequences'
// texts back to back + offsets; output: one finished NUCLSEQ datum image per sequence (tuples.cu describes the layout).
//
// The scalar code walks a sequence once, carrying three pieces of state: the index of the current hole, the previous character
// and the state of std::minstd_rand(holes_num ^ len).  Here a thread owns 16 consecutive bases of one sequence ("chunk"):
//   k_scan_chunks  counts, per chunk, the ambiguous bases and the hole starts (a base whose letter is ambiguous and differs from
//                  the previous character of the same sequence), validates the letters;
//   device scans   turn the counts into ranks;
//   k_record_sizes holes_num per sequence -> image size, RNG seed;
//   k_fill_chunks  writes the chunk's 32-bit word of 2-bit codes -- the r-th ambiguous base of a sequence (and, behind the last
//                  base, the tail padding) takes draw r of the generator, reached by a jump x_r = x_0 * 48271^r mod (2^31 - 1)
//                  and then stepped -- and the hole records: the thread that sees a run start writes offset and letter, the one
//                  that sees the run end adds the length (both contribute to `len` with atomicAdd: end + 1 - offset).
#include "pipeline.cuh"
#include "primitives.cuh"
#include "loader.cuh"

namespace {

constexpr int LD_THREADS = 256;
constexpr uint64_t MINSTD_M = 2147483647ull, MINSTD_A = 48271ull;

__device__ __forceinline__ int code_of(uint8_t c) {
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
}
__device__ __forceinline__ bool allowed(uint8_t c) {   // sequence.h:16: "ACGTNWSMKRYBDHV"
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'W': case 'S': case 'M': case 'K': case 'R': case 'Y': case 'B': case 'D': case 'H': case 'V': return true;
        default: return false;
    }
}
__device__ __forceinline__ uint32_t record_of_chunk(const LoaderParams& P, uint64_t chunk) {   // last r with chunk_off[r] <= chunk
    uint64_t lo = 0, hi = P.n_seqs;
    while (lo + 1 < hi) { const uint64_t mid = (lo + hi) >> 1; if (P.chunk_off[mid] <= chunk) lo = mid; else hi = mid; }
    return (uint32_t)lo;
}
__device__ __forceinline__ uint64_t mulmod(uint64_t a, uint64_t b) { return a * b % MINSTD_M; }   // a, b < 2^31
__device__ __forceinline__ uint64_t minstd_jump(uint64_t x0, uint64_t r) {                       // x0 * A^r mod M
    uint64_t base = MINSTD_A, acc = x0;
    while (r) { if (r & 1) acc = mulmod(acc, base); base = mulmod(base, base); r >>= 1; }
    return acc;
}

// masks of one chunk: bit k = base k of the chunk is ambiguous / starts a hole / ends a hole
struct ChunkMasks { uint32_t amb, start, end; int n; };
__device__ __forceinline__ ChunkMasks chunk_masks(const uint8_t* t, uint64_t rec_len, uint64_t j0, uint32_t* invalid_at) {
    ChunkMasks m; m.amb = m.start = m.end = 0;
    const uint64_t left = rec_len - j0;
    m.n = left < 16 ? (int)left : 16;
    uint8_t prev = j0 ? t[j0 - 1] : 0;
    for (int k = 0; k < m.n; ++k) {
        const uint8_t c = t[j0 + k];
        if (invalid_at && !allowed(c) && *invalid_at == 0xffffffffu) *invalid_at = (uint32_t)k;
        if (code_of(c) >= 4) {
            m.amb |= 1u << k;
            if (c != prev) m.start |= 1u << k;
            const bool last = j0 + k + 1 == rec_len;
            if (last || t[j0 + k + 1] != c) m.end |= 1u << k;
        }
        prev = c;
    }
    return m;
}

__global__ void __launch_bounds__(LD_THREADS) k_scan_chunks(LoaderParams P) {
    const uint64_t chunk = (uint64_t)blockIdx.x * LD_THREADS + threadIdx.x;
    if (chunk >= P.n_chunks) return;
    const uint32_t r = record_of_chunk(P, chunk);
    const uint64_t j0 = (chunk - P.chunk_off[r]) << 4;
    const uint8_t* t = P.text + P.offs[r];
    uint32_t bad = 0xffffffffu;
    const ChunkMasks m = chunk_masks(t, P.offs[r + 1] - P.offs[r], j0, &bad);
    if (bad != 0xffffffffu) atomicMin(P.first_invalid, (unsigned long long)(P.offs[r] + j0 + bad));
    P.cnt_amb[chunk] = (uint32_t)__popc(m.amb);
    P.cnt_start[chunk] = (uint32_t)__popc(m.start);
}

// rank arrays hold exclusive prefix sums over chunks (n_chunks + 1 entries)
__global__ void __launch_bounds__(LD_THREADS) k_record_sizes(LoaderParams P) {
    const uint64_t r = (uint64_t)blockIdx.x * LD_THREADS + threadIdx.x;
    if (r >= P.n_seqs) return;
    const uint64_t c0 = P.chunk_off[r], c1 = P.chunk_off[r + 1];
    const uint64_t holes = P.cnt_start[c1] - P.cnt_start[c0];
    const uint64_t len = P.offs[r + 1] - P.offs[r];
    P.holes_num[r] = (uint32_t)holes;
    const uint64_t size = 12 + 16 * holes + ((len + 3) >> 2);
    // a varlena length word holds 30 bits: the reference fails such a datum in palloc0 (MaxAllocSize, sequence.cpp:59-69)
    if (size >= 0x3fffffffull) atomicMin(P.first_invalid, (unsigned long long)P.offs[r] | (1ull << 63));
    P.img_off[r] = (size + 7) & ~7ull;
}

__global__ void __launch_bounds__(LD_THREADS) k_fill_headers(LoaderParams P) {
    const uint64_t r = (uint64_t)blockIdx.x * LD_THREADS + threadIdx.x;
    if (r >= P.n_seqs) return;
    const uint64_t len = P.offs[r + 1] - P.offs[r];
    const uint32_t holes = P.holes_num[r];
    uint32_t* hdr = reinterpret_cast<uint32_t*>(P.bytes + P.img_off[r]);
    hdr[0] = (uint32_t)(12 + 16ull * holes + ((len + 3) >> 2)) << 2;   // SET_VARSIZE, uncompressed 4-byte header
    hdr[1] = holes; hdr[2] = (uint32_t)len;
}

__global__ void __launch_bounds__(LD_THREADS) k_fill_chunks(LoaderParams P) {
    const uint64_t chunk = (uint64_t)blockIdx.x * LD_THREADS + threadIdx.x;
    if (chunk >= P.n_chunks) return;
    const uint32_t r = record_of_chunk(P, chunk);
    const uint64_t c0 = P.chunk_off[r];
    const uint64_t j0 = (chunk - c0) << 4;
    const uint8_t* t = P.text + P.offs[r];
    const uint64_t len = P.offs[r + 1] - P.offs[r];
    const ChunkMasks m = chunk_masks(t, len, j0, nullptr);
    const uint32_t holes_num = P.holes_num[r];
    uint8_t* img = P.bytes + P.img_off[r];
    uint8_t* holes = img + 12;
    // ---- hole records
    if (m.start | m.end) {
        uint32_t hole_i = P.cnt_start[chunk] - P.cnt_start[c0];       // index of the first hole that STARTS in this chunk
        for (int k = 0; k < m.n; ++k) {
            const bool st = (m.start >> k) & 1u, en = (m.end >> k) & 1u;
            if (st) {
                uint8_t* h = holes + 16ull * hole_i;
                const uint64_t off = j0 + (uint64_t)k;
                *reinterpret_cast<uint32_t*>(h) = (uint32_t)off; *reinterpret_cast<uint32_t*>(h + 4) = (uint32_t)(off >> 32);
                h[12] = t[off];
                atomicAdd(reinterpret_cast<int*>(h + 8), -(int)(uint32_t)off);
                ++hole_i;
            }
            if (en) {   // the run ending here is hole (number of starts up to and including this base) - 1
                uint8_t* h = holes + 16ull * (hole_i - 1);
                atomicAdd(reinterpret_cast<int*>(h + 8), (int)(uint32_t)(j0 + (uint64_t)k + 1));
            }
        }
    }
    // ---- codes: 16 bases -> 4 bytes (byte b = bases 4b .. 4b+3, first base in the top bits)
    const uint64_t padded = ((len + 3) >> 2) << 2;                      // the tail padding of the last byte is drawn too
    const bool tail = j0 + 16 >= len && padded > len;
    uint32_t word = 0;
    if (m.amb || tail) {
        const uint64_t amb_before = P.cnt_amb[chunk] - P.cnt_amb[c0];   // ambiguous bases of this sequence before the chunk
        uint64_t x0 = ((uint64_t)holes_num ^ len) % MINSTD_M;
        if (x0 == 0) x0 = 1;
        uint64_t x = minstd_jump(x0, amb_before);                      // state after amb_before draws
        for (int k = 0; k < 16; ++k) {
            const uint64_t j = j0 + (uint64_t)k;
            uint32_t code = 0;
            if (j < len) {
                if ((m.amb >> k) & 1u) { x = mulmod(x, MINSTD_A); code = (uint32_t)x & 3u; }
                else code = (uint32_t)code_of(t[j]);
            } else if (j < padded) { x = mulmod(x, MINSTD_A); code = (uint32_t)x & 3u; }
            word |= code << (((k & ~3) << 1) + ((~k & 3) << 1));
        }
    } else {
        for (int k = 0; k < m.n; ++k) word |= (uint32_t)code_of(t[j0 + k]) << (((k & ~3) << 1) + ((~k & 3) << 1));
    }
    *reinterpret_cast<uint32_t*>(holes + 16ull * holes_num + (j0 >> 2)) = word;
}

}  // namespace

size_t loader_scan_tmp_elems(uint64_t n) { return prim::scan_tmp_elems(n + 1) + 16; }

void launch_loader_scan(const LoaderParams& P, uint32_t* tmp32, uint64_t* tmp64, cudaStream_t st, uint64_t* launches) {
    if (P.n_chunks) {
        k_scan_chunks<<<(unsigned)((P.n_chunks + LD_THREADS - 1) / LD_THREADS), LD_THREADS, 0, st>>>(P);
        if (launches) ++*launches;
    }
    prim::device_scan<uint32_t, prim::OpSum, false>(P.cnt_amb, P.cnt_amb, (size_t)(P.n_chunks + 1), tmp32, prim::OpSum(), st, launches);
    prim::device_scan<uint32_t, prim::OpSum, false>(P.cnt_start, P.cnt_start, (size_t)(P.n_chunks + 1), tmp32, prim::OpSum(), st, launches);
    k_record_sizes<<<(unsigned)((P.n_seqs + LD_THREADS - 1) / LD_THREADS), LD_THREADS, 0, st>>>(P);
    if (launches) ++*launches;
    prim::device_scan<uint64_t, prim::OpSum, false>(P.img_off, P.img_off, (size_t)(P.n_seqs + 1), tmp64, prim::OpSum(), st, launches);
}
void launch_loader_fill(const LoaderParams& P, cudaStream_t st, uint64_t* launches) {
    k_fill_headers<<<(unsigned)((P.n_seqs + LD_THREADS - 1) / LD_THREADS), LD_THREADS, 0, st>>>(P);
    if (launches) ++*launches;
    if (P.n_chunks) {
        k_fill_chunks<<<(unsigned)((P.n_chunks + LD_THREADS - 1) / LD_THREADS), LD_THREADS, 0, st>>>(P);
        if (launches) ++*launches;
    }
}
