// index_build.cu -- GPU construction of the FM-index that reference bioseqdb/bwa.cpp:20-53,107-128
// builds with libbwa (pac2bwt -> is_bwt, bwt_bwtupdate_core, bwt_cal_sa).  SURVEY.md 7a / 8a rows a5-a7.
//
//   idx_text      pac -> T = fwd || revcomp (1 byte per symbol), symbol histogram -> L2
//   idx_keys      28-symbol prefix keys (56 bits) + 8-bit length tag (a shorter suffix sorts first)
//   radix sort    hand-written LSD sort (primitives.cuh)
//   doubling      Larsson-Sadakane style prefix doubling on the suffixes that are still tied
//   idx_bwt       L[r] = T[SA[r]-1], primary, $ row dropped; Occ checkpoints every 128 symbols in
//                 bwa's interleaved layout; the full SA is kept in HBM (sa[k/32] = SA[k] view on request)
// All kernels are HBM-stream bound; the algorithmic bytes of the radix passes are counted in
// IndexBuild::sort_pass_bytes.
#include "index_build.cuh"
#include "primitives.cuh"
#include "../../include/bioseqdb_gpu.h"

namespace {

constexpr int KSYM = 28;  // symbols in the initial key

__global__ void k_text(const uint8_t* __restrict__ pac, int64_t l_pac, uint8_t* __restrict__ T, unsigned long long* __restrict__ hist) {
    __shared__ unsigned int h[4];
    if (threadIdx.x < 4) h[threadIdx.x] = 0;
    __syncthreads();
    unsigned int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int64_t nbytes = l_pac >> 2;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += (int64_t)gridDim.x * blockDim.x) {
        uint32_t byte = pac[b];
        uint32_t s0 = byte >> 6, s1 = (byte >> 4) & 3, s2 = (byte >> 2) & 3, s3 = byte & 3;
        *reinterpret_cast<uint32_t*>(T + 4 * b) = s0 | s1 << 8 | s2 << 16 | s3 << 24;
        // reverse complement: T[2 l_pac - 1 - i] = 3 - pac[i]
        int64_t r = 2 * l_pac - 4 - 4 * b;
        *reinterpret_cast<uint32_t*>(T + r) = (3 - s3) | (3 - s2) << 8 | (3 - s1) << 16 | (3 - s0) << 24;
        c0 += (s0 == 0) + (s1 == 0) + (s2 == 0) + (s3 == 0);
        c1 += (s0 == 1) + (s1 == 1) + (s2 == 1) + (s3 == 1);
        c2 += (s0 == 2) + (s1 == 2) + (s2 == 2) + (s3 == 2);
        c3 += (s0 == 3) + (s1 == 3) + (s2 == 3) + (s3 == 3);
    }
    atomicAdd(&h[0], c0); atomicAdd(&h[1], c1); atomicAdd(&h[2], c2); atomicAdd(&h[3], c3);
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// key(i) = first KSYM symbols of suffix i (zero padded) << 8 | min(KSYM, n - i)
__global__ void k_keys(const uint8_t* __restrict__ T, uint64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = 0;
        uint64_t rem = n - i;
        int len = rem < (uint64_t)KSYM ? (int)rem : KSYM;
#pragma unroll
        for (int j = 0; j < KSYM; ++j) k = k << 2 | (j < len ? (uint64_t)T[i + j] : 0ull);
        keys[i] = k << 8 | (uint64_t)len;
        vals[i] = (uint32_t)i;
    }
}

// head flags of equal-key groups in the sorted order
__global__ void k_heads(const uint64_t* __restrict__ keys, uint64_t n, uint32_t* __restrict__ headpos) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        headpos[i] = (i == 0 || keys[i] != keys[i - 1]) ? (uint32_t)i : 0u;  // max-scan input (position 0 is a head with value 0)
}
__global__ void k_set_isa(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ grp, uint64_t n, uint32_t* __restrict__ isa) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) isa[sa[i]] = grp[i];
}
// tied[i] = 1 when the group of sorted position i has more than one member
__global__ void k_tied(const uint32_t* __restrict__ grp, uint64_t n, uint32_t* __restrict__ tied) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        bool head = grp[i] == (uint32_t)i;
        bool next_head = (i + 1 == n) || grp[i + 1] == (uint32_t)(i + 1);
        tied[i] = (head && next_head) ? 0u : 1u;
    }
}
__global__ void k_compact(const uint32_t* __restrict__ tied, const uint32_t* __restrict__ slot, const uint32_t* __restrict__ sa,
                          const uint32_t* __restrict__ grp, const uint32_t* __restrict__ isa, uint64_t n, uint64_t h,
                          uint32_t* __restrict__ pos, uint64_t* __restrict__ key2, uint32_t* __restrict__ val2) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!tied[i]) continue;
        uint32_t k = slot[i];
        uint32_t s = sa[i];
        uint64_t nx = (uint64_t)s + h;
        uint32_t rk = nx < n ? isa[nx] + 1u : 0u;  // the empty suffix ($) is the smallest
        pos[k] = (uint32_t)i;
        key2[k] = (uint64_t)grp[i] << 32 | rk;
        val2[k] = s;
    }
}
__global__ void k_writeback(const uint32_t* __restrict__ pos, const uint64_t* __restrict__ key2, const uint32_t* __restrict__ val2,
                            uint64_t m, uint32_t* __restrict__ sa, uint32_t* __restrict__ headpos) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t p = pos[k];
        sa[p] = val2[k];
        headpos[k] = (k == 0 || key2[k] != key2[k - 1]) ? p : 0u;
    }
}
__global__ void k_regroup(const uint32_t* __restrict__ pos, const uint32_t* __restrict__ newgrp, const uint32_t* __restrict__ val2,
                          uint64_t m, uint32_t* __restrict__ grp, uint32_t* __restrict__ isa) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        grp[pos[k]] = newgrp[k];
        isa[val2[k]] = newgrp[k];
    }
}

// rows r = 0..n of the full SA: SAfull[0] = n, SAfull[r] = sa[r-1].  One thread per 16 BWT symbols (one u32).
template <class SaT>
__global__ void k_find_primary(const SaT* __restrict__ sa, uint64_t n, unsigned long long* primary) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (sa[i] == 0) *primary = i + 1;
}
template <class SaT>
__global__ void k_full_sa(const uint32_t* __restrict__ sa, uint64_t n, SaT* __restrict__ out) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (uint64_t)gridDim.x * blockDim.x)
        out[r] = r == 0 ? (SaT)n : (SaT)sa[r - 1];
}
// B[j] for j in [0,n): row = j < primary ? j : j + 1 (the $ row is skipped); B[j] = T[SAfull[row] - 1].
// Each thread packs 16 symbols; each group of 8 threads covers one 128-symbol Occ block and produces its counts.
template <class SaT>
__global__ void k_bwt_blocks(const SaT* __restrict__ sa, const uint8_t* __restrict__ T, uint64_t n, uint64_t primary,
                             uint32_t* __restrict__ occ, unsigned long long* __restrict__ blk_cnt /* [4][n_blocks] */, uint64_t n_blocks) {
    uint64_t n_words = (n + 15) >> 4;
    for (uint64_t wd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; wd < ((n_words + 31) & ~31ull); wd += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t word = 0; uint32_t cnt = 0;  // byte-packed counts of A,C,G,T in this word
        if (wd < n_words) {
#pragma unroll 4
            for (int s = 0; s < 16; ++s) {
                uint64_t j = wd * 16 + s;
                if (j < n) {
                    uint64_t row = j < primary ? j : j + 1;       // row >= 1 here unless primary == ... row 0 is '$' suffix
                    uint64_t sfx = row == 0 ? n : (uint64_t)sa[row - 1];
                    uint32_t c = T[sfx - 1];                      // sfx != 0 because the primary row is skipped
                    word |= c << ((15 - s) << 1);
                    cnt += 1u << (c << 3);
                }
            }
            occ[(wd >> 3 << 4) + 8 + (wd & 7)] = word;
        }
        // reduce the counts of the 8 words of a block
        cnt += __shfl_xor_sync(FULL, cnt, 1); cnt += __shfl_xor_sync(FULL, cnt, 2); cnt += __shfl_xor_sync(FULL, cnt, 4);
        uint64_t blk = wd >> 3;
        if ((wd & 7) == 0 && blk < n_blocks) {
#pragma unroll
            for (int c = 0; c < 4; ++c) blk_cnt[(uint64_t)c * n_blocks + blk] = (cnt >> (c << 3)) & 0xff;
        }
    }
}
__global__ void k_occ_counts(const unsigned long long* __restrict__ blk_excl, uint64_t n_blocks, uint32_t* __restrict__ occ) {
    // blk_excl[c][b] = # of symbol c in B[0, 128 b); b in [0, n_blocks] (the last one is the trailing record)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n_blocks + 1) * 4; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b = i >> 2; int c = (int)(i & 3);
        unsigned long long v = blk_excl[(uint64_t)c * (n_blocks + 1) + b];
        reinterpret_cast<unsigned long long*>(occ + (b << 4))[c] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Wide build (n + 1 >= 2^32 rows, or forced for tests): 64-bit suffix ids.  Suffixes are first bucketed by
// their 4 leading symbols with ONE counting-sort pass whose keys and values are computed from the element
// index (nothing but the 8-byte ids is materialised), then every bucket is sorted on the next 28 symbols
// (+ length tag) with the same radix sort, in place inside the final SA array.  Ties that survive the
// 32-symbol prefix are resolved by prefix doubling on a compacted list with a full inverse SA.
constexpr int PFX = 4;          // symbols of the bucket prefix (256 buckets)

struct SrcPrefix {              // key(i) = 4 leading symbols of suffix i (zero padded), val(i) = i
    const uint8_t* T; uint64_t n;
    __device__ __forceinline__ uint64_t key(size_t i) const {
        uint64_t k = 0;
#pragma unroll
        for (int j = 0; j < PFX; ++j) k = k << 2 | ((uint64_t)i + j < n ? (uint64_t)T[i + j] : 0ull);
        return k;
    }
    __device__ __forceinline__ uint64_t val(size_t i) const { return (uint64_t)i; }
};

// key of suffix id within its bucket: the 28 symbols after the prefix, then a tag that orders suffixes which
// end inside the key window: tag = n - id (< 4) for suffixes shorter than the prefix, 4 + min(28, n - id - 4) else
__global__ void k_keys_wide(const uint8_t* __restrict__ T, uint64_t n, const uint64_t* __restrict__ ids, uint64_t m, uint64_t* __restrict__ keys) {
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = ids[t];
        const uint64_t rem = n - i;
        uint64_t k = 0; uint64_t tag;
        if (rem < (uint64_t)PFX) tag = rem;
        else {
            const uint64_t r2 = rem - PFX;
            const int len = r2 < (uint64_t)KSYM ? (int)r2 : KSYM;
#pragma unroll
            for (int j = 0; j < KSYM; ++j) k = k << 2 | (j < len ? (uint64_t)T[i + PFX + j] : 0ull);
            tag = (uint64_t)(PFX + len);
        }
        keys[t] = k << 8 | tag;
    }
}

// same first PFX + KSYM symbols and same length class => still tied after the bucket sorts
__device__ __forceinline__ bool same_prefix(const uint8_t* T, uint64_t n, uint64_t a, uint64_t b) {
    const uint64_t ra = n - a, rb = n - b;
    const uint64_t la = ra < (uint64_t)(PFX + KSYM) ? ra : (uint64_t)(PFX + KSYM), lb = rb < (uint64_t)(PFX + KSYM) ? rb : (uint64_t)(PFX + KSYM);
    if (la != lb) return false;
    for (uint64_t j = 0; j < la; ++j) if (T[a + j] != T[b + j]) return false;
    return true;
}
__device__ __forceinline__ bool wide_head(const uint8_t* T, uint64_t n, const uint64_t* sa, uint64_t i) {
    return i == 0 || !same_prefix(T, n, sa[i - 1], sa[i]);
}
__device__ __forceinline__ bool wide_tied(const uint8_t* T, uint64_t n, const uint64_t* sa, uint64_t i) {
    const bool h = wide_head(T, n, sa, i);
    const bool nh = (i + 1 == n) || wide_head(T, n, sa, i + 1);
    return !(h && nh);
}

constexpr int CT_THREADS = 256, CT_TILE = 2048;
// per-tile count of tied positions
__global__ void __launch_bounds__(CT_THREADS) k_wide_tie_count(const uint8_t* __restrict__ T, uint64_t n, const uint64_t* __restrict__ sa, unsigned long long* tile_cnt) {
    __shared__ unsigned long long sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * CT_TILE;
    unsigned long long c = 0;
    for (int k = threadIdx.x; k < CT_TILE; k += CT_THREADS) { const uint64_t i = base + k; if (i < n && wide_tied(T, n, sa, i)) ++c; }
    unsigned long long tot;
    prim::block_inclusive<unsigned long long, prim::OpSum>(c, prim::OpSum(), &tot, sh);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}
// compaction of tied positions (order preserved): pos[k], val[k] = sa[pos], headpos[k] = pos if group head else 0
__global__ void __launch_bounds__(CT_THREADS) k_wide_tie_compact(const uint8_t* __restrict__ T, uint64_t n, const uint64_t* __restrict__ sa,
                                                                 const unsigned long long* __restrict__ tile_off, uint64_t* __restrict__ pos,
                                                                 uint64_t* __restrict__ val, unsigned long long* __restrict__ headpos) {
    __shared__ unsigned long long sh[32];
    const uint64_t base = (uint64_t)blockIdx.x * CT_TILE;
    unsigned long long run = tile_off[blockIdx.x];
    for (int k0 = 0; k0 < CT_TILE; k0 += CT_THREADS) {     // contiguous chunks keep the order
        const uint64_t i = base + k0 + threadIdx.x;
        const bool t = i < n && wide_tied(T, n, sa, i);
        unsigned long long tot;
        const unsigned long long inc = prim::block_inclusive<unsigned long long, prim::OpSum>(t ? 1ull : 0ull, prim::OpSum(), &tot, sh);
        if (t) { const unsigned long long k = run + inc - 1; pos[k] = i; val[k] = sa[i]; headpos[k] = wide_head(T, n, sa, i) ? i : 0ull; }
        run += tot;
    }
}
__global__ void k_wide_isa_init(const uint64_t* __restrict__ sa, uint64_t n, uint64_t* __restrict__ isa) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) isa[sa[i]] = i;
}
__global__ void k_wide_isa_set(const uint64_t* __restrict__ val, const unsigned long long* __restrict__ grp, uint64_t m, uint64_t* __restrict__ isa) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) isa[val[k]] = grp[k];
}
// dense group index of each compact element: gflag[k] = 1 at group heads (headpos != 0 or k == 0 with pos 0)
__global__ void k_wide_gflag(const uint64_t* __restrict__ pos, const unsigned long long* __restrict__ grp, uint64_t m, uint32_t* __restrict__ gflag) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) gflag[k] = grp[k] == pos[k] ? 1u : 0u;
}
// composite key = (dense group index) << 33 | (rank of suffix val + h, +1; 0 = past the end)
__global__ void k_wide_key2(const uint64_t* __restrict__ val, const uint32_t* __restrict__ gidx, const uint64_t* __restrict__ isa, uint64_t n, uint64_t h,
                            uint64_t m, uint64_t* __restrict__ key2) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t nx = val[k] + h;
        const uint64_t rk = nx < n ? isa[nx] + 1ull : 0ull;
        key2[k] = (uint64_t)(gidx[k] - 1u) << 33 | rk;
    }
}
__global__ void k_wide_writeback(const uint64_t* __restrict__ pos, const uint64_t* __restrict__ key2, const uint64_t* __restrict__ val2, uint64_t m,
                                 uint64_t* __restrict__ sa, unsigned long long* __restrict__ headpos) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = pos[k];
        sa[p] = val2[k];
        headpos[k] = (k == 0 || key2[k] != key2[k - 1]) ? p : 0ull;
    }
}
// keep the elements whose (new) group still has more than one member
__global__ void k_wide_still_tied(const uint64_t* __restrict__ pos, const unsigned long long* __restrict__ grp, uint64_t m, uint32_t* __restrict__ keep) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
        const bool head = grp[k] == pos[k];
        const bool next_head = (k + 1 == m) || grp[k + 1] == pos[k + 1];
        keep[k] = (head && next_head) ? 0u : 1u;
    }
}
__global__ void k_wide_recompact(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ slot, const uint64_t* __restrict__ pos,
                                 const uint64_t* __restrict__ val, const unsigned long long* __restrict__ grp, uint64_t m,
                                 uint64_t* __restrict__ pos2, uint64_t* __restrict__ val2, unsigned long long* __restrict__ grp2) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x)
        if (keep[k]) { const uint32_t d = slot[k]; pos2[d] = pos[k]; val2[d] = val[k]; grp2[d] = grp[k]; }
}

}  // namespace

static inline unsigned grid_for(uint64_t n, int threads = 256) {
    uint64_t g = (n + threads - 1) / threads;
    uint64_t cap = 148ull * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

#define BCHECK_DEFINED 1
#define BCHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bsq_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); goto fail; } } while (0)

// ---- wide build driver (see the kernel block above)
static int build_index_wide(IndexBuild& B, cudaStream_t st) {
    const int64_t l_pac = B.l_pac;
    const uint64_t n = (uint64_t)l_pac * 2;
    B.seq_len = n; B.launches = 0; B.sort_pass_bytes = 0; B.doubling_rounds = 0;
    if (n >= (1ull << 33)) { bsq_set_error("index build: texts of 2^33 symbols or more are not supported"); return BSQ_ERR; }
    using u64 = uint64_t; using ull = unsigned long long;
    uint8_t* T = nullptr; u64* sa_full = nullptr; ull* hist = nullptr; ull* scan_tmp = nullptr; ull* d_hist = nullptr;
    u64 *kb0 = nullptr, *kb1 = nullptr, *vb1 = nullptr; ull* blk_cnt = nullptr; ull* scan_tmp64 = nullptr;
    u64 *isa = nullptr, *pos = nullptr, *val = nullptr, *pos2 = nullptr, *val2 = nullptr, *key2 = nullptr, *key2b = nullptr, *valb = nullptr;
    ull *grp = nullptr, *grp2 = nullptr, *tile_cnt = nullptr; uint32_t *gflag = nullptr, *slot = nullptr, *scan32 = nullptr;
    prim::RadixWorkspace ws;
    const u64 tiles = prim::rs_tiles(n);
    const u64 n_blocks = (n + 127) / 128;
    ull* d_primary = nullptr;
    u64 bucket_start[257];
    u64 max_bucket = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    using SrcA = prim::SrcArrays<u64, u64>;
    BCHECK((prim::rs_prepare<u64, u64, SrcA, uint32_t>()));
    BCHECK((prim::rs_prepare<u64, u64, SrcPrefix, ull>()));
    BCHECK(cudaMalloc(&T, n + 64));
    BCHECK(cudaMalloc(&sa_full, (n + 1) * 8 + 64));
    BCHECK(cudaMalloc(&d_hist, 64));
    BCHECK(cudaMemsetAsync(d_hist, 0, 64, st));
    d_primary = d_hist + 4;
    k_text<<<grid_for((u64)l_pac / 4), 256, 0, st>>>(B.d_pac, l_pac, T, d_hist); ++B.launches;
    {
        u64* sa = sa_full + 1;
        // ---- one counting-sort pass by the 4 leading symbols, ids generated on the fly
        BCHECK(cudaMalloc(&hist, tiles * prim::RS_BINS * 8));
        BCHECK(cudaMalloc(&scan_tmp, (prim::scan_tmp_elems(tiles * prim::RS_BINS) + 16) * 8));
        SrcPrefix sp{T, n};
        prim::k_rs_hist<SrcPrefix, ull><<<(unsigned)tiles, prim::RS_THREADS, 0, st>>>(sp, n, 0, hist, (uint32_t)tiles); ++B.launches;
        prim::device_scan<ull, prim::OpSum, false>(hist, hist, tiles * prim::RS_BINS, scan_tmp, prim::OpSum(), st, &B.launches);
        prim::k_rs_scatter<u64, u64, SrcPrefix, ull><<<(unsigned)tiles, prim::RS_THREADS, prim::rs_scatter_smem<u64, u64, ull>(), st>>>(
            sp, (u64*)nullptr, sa, n, 0, hist, (uint32_t)tiles); ++B.launches;
        B.sort_pass_bytes += n * (1 + 1 + 8);
        for (int b = 0; b < 256; ++b) BCHECK(cudaMemcpyAsync(&bucket_start[b], hist + (u64)b * tiles, 8, cudaMemcpyDeviceToHost, st));
        BCHECK(cudaStreamSynchronize(st));
        bucket_start[256] = n;
        for (int b = 0; b < 256; ++b) max_bucket = std::max<u64>(max_bucket, bucket_start[b + 1] - bucket_start[b]);
        cudaFree(hist); hist = nullptr; cudaFree(scan_tmp); scan_tmp = nullptr;
        if (max_bucket >= 0xffffffffull) { bsq_set_error("index build: a 4-symbol bucket holds %llu suffixes (low-complexity text); not supported at this size", (ull)max_bucket); goto fail; }
        // ---- per-bucket sort on the next 28 symbols
        BCHECK(cudaMalloc(&kb0, max_bucket * 8 + 64)); BCHECK(cudaMalloc(&kb1, max_bucket * 8 + 64)); BCHECK(cudaMalloc(&vb1, max_bucket * 8 + 64));
        BCHECK(cudaMalloc(&ws.hist, prim::rs_tiles(max_bucket) * prim::RS_BINS * 4 + 64));
        BCHECK(cudaMalloc(&ws.scan_tmp, (prim::scan_tmp_elems(prim::rs_tiles(max_bucket) * prim::RS_BINS) + 16) * 4));
        for (int b = 0; b < 256; ++b) {
            const u64 m = bucket_start[b + 1] - bucket_start[b];
            if (m < 2) continue;
            u64* ids = sa + bucket_start[b];
            k_keys_wide<<<grid_for(m), 256, 0, st>>>(T, n, ids, m, kb0); ++B.launches;
            const int r = prim::radix_sort_pairs<u64, u64>(kb0, ids, kb1, vb1, m, 0, 64, ws, st, &B.launches, &B.sort_pass_bytes);
            if (r) BCHECK(cudaMemcpyAsync(ids, vb1, m * 8, cudaMemcpyDeviceToDevice, st));
        }
        BCHECK(cudaStreamSynchronize(st));
        cudaFree(kb0); kb0 = nullptr; cudaFree(kb1); kb1 = nullptr; cudaFree(vb1); vb1 = nullptr;
        // ---- ties beyond 32 symbols: prefix doubling on the compacted tied set
        const u64 ct_tiles = (n + CT_TILE - 1) / CT_TILE;
        BCHECK(cudaMalloc(&tile_cnt, (ct_tiles + 1) * 8));
        BCHECK(cudaMalloc(&scan_tmp, (prim::scan_tmp_elems(ct_tiles + 1) + 16) * 8));
        BCHECK(cudaMemsetAsync(tile_cnt, 0, (ct_tiles + 1) * 8, st));
        k_wide_tie_count<<<(unsigned)ct_tiles, CT_THREADS, 0, st>>>(T, n, sa, tile_cnt); ++B.launches;
        prim::device_scan<ull, prim::OpSum, false>(tile_cnt, tile_cnt, ct_tiles + 1, scan_tmp, prim::OpSum(), st, &B.launches);
        ull m_ull = 0;
        BCHECK(cudaMemcpyAsync(&m_ull, tile_cnt + ct_tiles, 8, cudaMemcpyDeviceToHost, st));
        BCHECK(cudaStreamSynchronize(st));
        u64 m = m_ull;
        if (m) {
            if (m >= (1ull << 31)) { bsq_set_error("index build: %llu suffixes tie on their first 32 symbols (highly repetitive text); not supported at this size", (ull)m); goto fail; }
            BCHECK(cudaMalloc(&isa, n * 8 + 64));
            BCHECK(cudaMalloc(&pos, m * 8)); BCHECK(cudaMalloc(&val, m * 8)); BCHECK(cudaMalloc(&grp, m * 8));
            BCHECK(cudaMalloc(&pos2, m * 8)); BCHECK(cudaMalloc(&val2, m * 8)); BCHECK(cudaMalloc(&grp2, m * 8));
            BCHECK(cudaMalloc(&key2, m * 8)); BCHECK(cudaMalloc(&key2b, m * 8)); BCHECK(cudaMalloc(&valb, m * 8));
            BCHECK(cudaMalloc(&gflag, m * 4 + 64)); BCHECK(cudaMalloc(&slot, m * 4 + 64));
            BCHECK(cudaMalloc(&scan32, (prim::scan_tmp_elems(m) + 16) * 8));
            cudaFree(ws.hist); cudaFree(ws.scan_tmp); ws.hist = nullptr; ws.scan_tmp = nullptr;
            BCHECK(cudaMalloc(&ws.hist, prim::rs_tiles(m) * prim::RS_BINS * 4 + 64));
            BCHECK(cudaMalloc(&ws.scan_tmp, (prim::scan_tmp_elems(prim::rs_tiles(m) * prim::RS_BINS) + 16) * 4));
            k_wide_tie_compact<<<(unsigned)ct_tiles, CT_THREADS, 0, st>>>(T, n, sa, tile_cnt, pos, val, grp); ++B.launches;
            prim::device_scan<ull, prim::OpMax, true>(grp, grp, m, (ull*)scan32, prim::OpMax(), st, &B.launches);   // group start positions
            k_wide_isa_init<<<grid_for(n), 256, 0, st>>>(sa, n, isa); ++B.launches;
            k_wide_isa_set<<<grid_for(m), 256, 0, st>>>(val, grp, m, isa); ++B.launches;
            u64 h = PFX + KSYM;
            while (m) {
                if (h >= 2 * n) { bsq_set_error("index build: prefix doubling did not converge"); goto fail; }
                ++B.doubling_rounds;
                k_wide_gflag<<<grid_for(m), 256, 0, st>>>(pos, grp, m, gflag); ++B.launches;
                prim::device_scan<uint32_t, prim::OpSum, true>(gflag, gflag, m, (uint32_t*)scan32, prim::OpSum(), st, &B.launches);   // dense group index (1-based)
                k_wide_key2<<<grid_for(m), 256, 0, st>>>(val, gflag, isa, n, h, m, key2); ++B.launches;
                // rank occupies bits [0, 33), the dense group index bits [33, 64): sort all eight bytes (the tied set is small)
                int rr = prim::radix_sort_pairs<u64, u64>(key2, val, key2b, valb, m, 0, 40, ws, st, &B.launches, &B.sort_pass_bytes);
                u64 *kx = rr ? key2b : key2, *vx = rr ? valb : val, *ky = rr ? key2 : key2b, *vy = rr ? val : valb;
                rr = prim::radix_sort_pairs<u64, u64>(kx, vx, ky, vy, m, 40, 64, ws, st, &B.launches, &B.sort_pass_bytes);
                u64 *kf = rr ? ky : kx, *vf = rr ? vy : vx;
                k_wide_writeback<<<grid_for(m), 256, 0, st>>>(pos, kf, vf, m, sa, grp2); ++B.launches;
                prim::device_scan<ull, prim::OpMax, true>(grp2, grp2, m, (ull*)scan32, prim::OpMax(), st, &B.launches);
                k_wide_isa_set<<<grid_for(m), 256, 0, st>>>(vf, grp2, m, isa); ++B.launches;
                // drop the elements that became singletons
                k_wide_still_tied<<<grid_for(m), 256, 0, st>>>(pos, grp2, m, gflag); ++B.launches;
                prim::device_scan<uint32_t, prim::OpSum, false>(gflag, slot, m, (uint32_t*)scan32, prim::OpSum(), st, &B.launches);
                uint32_t last_slot = 0, last_keep = 0;
                BCHECK(cudaMemcpyAsync(&last_slot, slot + (m - 1), 4, cudaMemcpyDeviceToHost, st));
                BCHECK(cudaMemcpyAsync(&last_keep, gflag + (m - 1), 4, cudaMemcpyDeviceToHost, st));
                BCHECK(cudaStreamSynchronize(st));
                const u64 m2 = (u64)last_slot + last_keep;
                if (m2) {
                    k_wide_recompact<<<grid_for(m), 256, 0, st>>>(gflag, slot, pos, vf, grp2, m, pos2, val2, grp); ++B.launches;
                    BCHECK(cudaStreamSynchronize(st));
                    std::swap(pos, pos2);
                    BCHECK(cudaMemcpyAsync(val, val2, m2 * 8, cudaMemcpyDeviceToDevice, st));
                }
                m = m2;
                h <<= 1;
            }
        }
        // ---- BWT, Occ from the u64 SA
        BCHECK(cudaMemsetAsync(d_primary, 0, 8, st));
        k_find_primary<u64><<<grid_for(n), 256, 0, st>>>(sa, n, d_primary); ++B.launches;
        {
            ull hp[8];
            BCHECK(cudaMemcpyAsync(hp, d_hist, 64, cudaMemcpyDeviceToHost, st));
            BCHECK(cudaStreamSynchronize(st));
            B.L2[0] = 0;
            for (int c = 0; c < 4; ++c) B.L2[c + 1] = B.L2[c] + hp[c] + hp[3 - c];
            B.primary = hp[4];
        }
        B.occ_bytes = (n_blocks + 1) * 64;
        BCHECK(cudaMalloc(&B.d_occ, B.occ_bytes + 64));
        BCHECK(cudaMemsetAsync(B.d_occ, 0, B.occ_bytes + 64, st));
        BCHECK(cudaMalloc(&blk_cnt, (n_blocks + 1) * 4 * 8));
        BCHECK(cudaMemsetAsync(blk_cnt, 0, (n_blocks + 1) * 4 * 8, st));
        BCHECK(cudaMalloc(&scan_tmp64, (prim::scan_tmp_elems(n_blocks + 1) + 8) * 8));
        k_bwt_blocks<u64><<<grid_for((n + 15) / 16), 256, 0, st>>>(sa, T, n, B.primary, B.d_occ, blk_cnt, n_blocks + 1); ++B.launches;
        for (int c = 0; c < 4; ++c)
            prim::device_scan<ull, prim::OpSum, false>(blk_cnt + (u64)c * (n_blocks + 1), blk_cnt + (u64)c * (n_blocks + 1), n_blocks + 1, scan_tmp64,
                                                       prim::OpSum(), st, &B.launches);
        k_occ_counts<<<grid_for((n_blocks + 1) * 4), 256, 0, st>>>(blk_cnt, n_blocks, B.d_occ); ++B.launches;
        BCHECK(cudaMemcpyAsync(sa_full, &n, 8, cudaMemcpyHostToDevice, st));   // row 0 = '$'
        B.sa_bytes = 8; B.d_sa = sa_full; sa_full = nullptr;
    }
    cudaEventRecord(e1, st);
    BCHECK(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); B.build_ms = ms; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(T); cudaFree(d_hist); cudaFree(ws.hist); cudaFree(ws.scan_tmp); cudaFree(scan_tmp); cudaFree(tile_cnt); cudaFree(blk_cnt); cudaFree(scan_tmp64);
    cudaFree(isa); cudaFree(pos); cudaFree(val); cudaFree(grp); cudaFree(pos2); cudaFree(val2); cudaFree(grp2); cudaFree(key2); cudaFree(key2b); cudaFree(valb);
    cudaFree(gflag); cudaFree(slot); cudaFree(scan32);
    return BSQ_OK;
fail:
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(T); cudaFree(sa_full); cudaFree(d_hist); cudaFree(hist); cudaFree(ws.hist); cudaFree(ws.scan_tmp); cudaFree(scan_tmp); cudaFree(tile_cnt);
    cudaFree(kb0); cudaFree(kb1); cudaFree(vb1); cudaFree(blk_cnt); cudaFree(scan_tmp64);
    cudaFree(isa); cudaFree(pos); cudaFree(val); cudaFree(grp); cudaFree(pos2); cudaFree(val2); cudaFree(grp2); cudaFree(key2); cudaFree(key2b); cudaFree(valb);
    cudaFree(gflag); cudaFree(slot); cudaFree(scan32);
    if (B.d_occ) { cudaFree(B.d_occ); B.d_occ = nullptr; }
    if (B.d_sa) { cudaFree(B.d_sa); B.d_sa = nullptr; }
    return BSQ_ERR;
}

int build_index_device(IndexBuild& B, cudaStream_t st) {
    const int64_t l_pac = B.l_pac;
    const uint64_t n = (uint64_t)l_pac * 2;
    B.seq_len = n;
    B.launches = 0; B.sort_pass_bytes = 0;
    if (n + 1 >= 0xffffffffull || B.force_wide) return build_index_wide(B, st);
    uint8_t* T = nullptr; uint64_t *k0 = nullptr, *k1 = nullptr; uint32_t *v0 = nullptr, *v1 = nullptr, *grp = nullptr, *isa = nullptr,
            *tied = nullptr, *slot = nullptr, *pos = nullptr, *scan_tmp = nullptr;
    unsigned long long *d_hist = nullptr, *d_primary = nullptr, *blk_cnt = nullptr, *scan_tmp64 = nullptr;
    prim::RadixWorkspace ws;
    uint32_t* sa = nullptr;
    const uint64_t n_blocks = (n + 127) / 128;
    size_t tmp_elems = prim::scan_tmp_elems((size_t)n) + prim::scan_tmp_elems(prim::rs_tiles(n) * prim::RS_BINS) + 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);

    BCHECK((prim::rs_prepare<uint64_t, uint32_t, prim::SrcArrays<uint64_t, uint32_t>, uint32_t>()));
    BCHECK(cudaMalloc(&T, n + 64));
    BCHECK(cudaMalloc(&k0, n * 8)); BCHECK(cudaMalloc(&k1, n * 8));
    BCHECK(cudaMalloc(&v0, n * 4)); BCHECK(cudaMalloc(&v1, n * 4));
    BCHECK(cudaMalloc(&grp, n * 4)); BCHECK(cudaMalloc(&isa, n * 4));
    BCHECK(cudaMalloc(&d_hist, 8 * 8));
    BCHECK(cudaMalloc(&scan_tmp, tmp_elems * 4));
    BCHECK(cudaMalloc(&ws.hist, prim::rs_tiles(n) * prim::RS_BINS * 4));
    ws.scan_tmp = scan_tmp;
    d_primary = d_hist + 4;
    BCHECK(cudaMemsetAsync(d_hist, 0, 64, st));

    k_text<<<grid_for((uint64_t)l_pac / 4), 256, 0, st>>>(B.d_pac, l_pac, T, d_hist); ++B.launches;
    k_keys<<<grid_for(n), 256, 0, st>>>(T, n, k0, v0); ++B.launches;
    {
        int r = prim::radix_sort_pairs<uint64_t, uint32_t>(k0, v0, k1, v1, n, 0, 64, ws, st, &B.launches, &B.sort_pass_bytes);
        uint64_t* ks = r ? k1 : k0; sa = r ? v1 : v0;
        uint32_t* spare = r ? v0 : v1;
        k_heads<<<grid_for(n), 256, 0, st>>>(ks, n, grp); ++B.launches;
        prim::device_scan<uint32_t, prim::OpMax, true>(grp, grp, n, scan_tmp, prim::OpMax(), st, &B.launches);
        k_set_isa<<<grid_for(n), 256, 0, st>>>(sa, grp, n, isa); ++B.launches;
        // prefix doubling on the tied suffixes
        tied = spare;                                   // n x u32
        slot = reinterpret_cast<uint32_t*>(r ? k1 : k0);  // sorted keys are dead after k_heads: reuse as 2 x (n x u32)
        pos = slot + n;
        uint64_t* key2 = r ? k0 : k1;                   // the other key buffer: m <= n u64 ... split in halves for ping-pong
        uint64_t h = KSYM;
        B.doubling_rounds = 0;
        for (;;) {
            k_tied<<<grid_for(n), 256, 0, st>>>(grp, n, tied); ++B.launches;
            prim::device_scan<uint32_t, prim::OpSum, false>(tied, slot, n, scan_tmp, prim::OpSum(), st, &B.launches);
            uint32_t last_slot = 0, last_tied = 0;
            BCHECK(cudaMemcpyAsync(&last_slot, slot + (n - 1), 4, cudaMemcpyDeviceToHost, st));
            BCHECK(cudaMemcpyAsync(&last_tied, tied + (n - 1), 4, cudaMemcpyDeviceToHost, st));
            BCHECK(cudaStreamSynchronize(st));
            uint64_t m = (uint64_t)last_slot + last_tied;
            if (m == 0) break;
            if (h >= n) { bsq_set_error("index build: prefix doubling did not converge"); goto fail; }
            ++B.doubling_rounds;
            // ping-pong buffers for the m tied elements: keys in key2[0..m) / key2[m..2m) needs 2m <= n:
            // when more than half of the suffixes are tied, allocate a dedicated buffer.
            uint64_t *ka = key2, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr;
            bool own = false;
            if (2 * m <= n) { kb = key2 + m; }
            else { BCHECK(cudaMalloc(&kb, m * 8)); own = true; }
            BCHECK(cudaMalloc(&va, m * 4)); BCHECK(cudaMalloc(&vb, m * 4));
            k_compact<<<grid_for(n), 256, 0, st>>>(tied, slot, sa, grp, isa, n, h, pos, ka, va); ++B.launches;
            int bits = 1; while ((1ull << bits) <= n) ++bits;   // rank+1 <= n and grp < n fit in `bits` bits
            int hi_end = 32 + ((bits + 7) / 8) * 8; if (hi_end > 64) hi_end = 64;
            int rr = prim::radix_sort_pairs<uint64_t, uint32_t>(ka, va, kb, vb, m, 0, ((bits + 7) / 8) * 8, ws, st, &B.launches, &B.sort_pass_bytes);
            uint64_t* kx = rr ? kb : ka; uint32_t* vx = rr ? vb : va; uint64_t* ky = rr ? ka : kb; uint32_t* vy = rr ? va : vb;
            rr = prim::radix_sort_pairs<uint64_t, uint32_t>(kx, vx, ky, vy, m, 32, hi_end, ws, st, &B.launches, &B.sort_pass_bytes);
            uint64_t* kf = rr ? ky : kx; uint32_t* vf = rr ? vy : vx; uint32_t* newgrp = rr ? vx : vy;
            k_writeback<<<grid_for(m), 256, 0, st>>>(pos, kf, vf, m, sa, newgrp); ++B.launches;
            prim::device_scan<uint32_t, prim::OpMax, true>(newgrp, newgrp, m, scan_tmp, prim::OpMax(), st, &B.launches);
            k_regroup<<<grid_for(m), 256, 0, st>>>(pos, newgrp, vf, m, grp, isa); ++B.launches;
            BCHECK(cudaStreamSynchronize(st));
            cudaFree(va); cudaFree(vb); if (own) cudaFree(kb);
            h <<= 1;
        }
    }
    // ---- BWT, Occ, full SA
    BCHECK(cudaMemsetAsync(d_primary, 0, 8, st));
    k_find_primary<uint32_t><<<grid_for(n), 256, 0, st>>>(sa, n, d_primary); ++B.launches;
    {
        unsigned long long hp[8];
        BCHECK(cudaMemcpyAsync(hp, d_hist, 64, cudaMemcpyDeviceToHost, st));
        BCHECK(cudaStreamSynchronize(st));
        // k_text counted the forward strand only; the reverse complement adds the mirrored counts
        B.L2[0] = 0;
        for (int c = 0; c < 4; ++c) B.L2[c + 1] = B.L2[c] + hp[c] + hp[3 - c];
        B.primary = hp[4];
    }
    B.occ_bytes = (n_blocks + 1) * 64;  // 64-byte blocks; the trailing count record sits at block n_blocks
    BCHECK(cudaMalloc(&B.d_occ, B.occ_bytes + 64));
    BCHECK(cudaMemsetAsync(B.d_occ, 0, B.occ_bytes + 64, st));
    BCHECK(cudaMalloc(&blk_cnt, (n_blocks + 1) * 4 * 8));
    BCHECK(cudaMemsetAsync(blk_cnt, 0, (n_blocks + 1) * 4 * 8, st));
    BCHECK(cudaMalloc(&scan_tmp64, (prim::scan_tmp_elems(n_blocks + 1) + 8) * 8));
    k_bwt_blocks<uint32_t><<<grid_for((n + 15) / 16), 256, 0, st>>>(sa, T, n, B.primary, B.d_occ, blk_cnt, n_blocks + 1); ++B.launches;
    for (int c = 0; c < 4; ++c)
        prim::device_scan<unsigned long long, prim::OpSum, false>(blk_cnt + (uint64_t)c * (n_blocks + 1), blk_cnt + (uint64_t)c * (n_blocks + 1),
                                                                  n_blocks + 1, scan_tmp64, prim::OpSum(), st, &B.launches);
    k_occ_counts<<<grid_for((n_blocks + 1) * 4), 256, 0, st>>>(blk_cnt, n_blocks, B.d_occ); ++B.launches;
    B.sa_bytes = 4;
    BCHECK(cudaMalloc(&B.d_sa, (n + 1) * 4 + 64));
    k_full_sa<uint32_t><<<grid_for(n + 1), 256, 0, st>>>(sa, n, (uint32_t*)B.d_sa); ++B.launches;
    cudaEventRecord(e1, st);
    BCHECK(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); B.build_ms = ms; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(T); cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(grp); cudaFree(isa); cudaFree(d_hist);
    cudaFree(scan_tmp); cudaFree(ws.hist); cudaFree(blk_cnt); cudaFree(scan_tmp64);
    return BSQ_OK;
fail:
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(T); cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(grp); cudaFree(isa); cudaFree(d_hist);
    cudaFree(scan_tmp); cudaFree(ws.hist); cudaFree(blk_cnt); cudaFree(scan_tmp64);
    if (B.d_occ) { cudaFree(B.d_occ); B.d_occ = nullptr; }
    if (B.d_sa) { cudaFree(B.d_sa); B.d_sa = nullptr; }
    return BSQ_ERR;
}
