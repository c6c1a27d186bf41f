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
__global__ void k_find_primary(const uint32_t* __restrict__ sa, uint64_t n, unsigned long long* primary) {
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
__global__ void k_bwt_blocks(const uint32_t* __restrict__ sa, const uint8_t* __restrict__ T, uint64_t n, uint64_t primary,
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
                    uint32_t sfx = row == 0 ? (uint32_t)n : sa[row - 1];
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

}  // namespace

static inline unsigned grid_for(uint64_t n, int threads = 256) {
    uint64_t g = (n + threads - 1) / threads;
    uint64_t cap = 148ull * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

#define BCHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bsq_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); goto fail; } } while (0)

int build_index_device(IndexBuild& B, cudaStream_t st) {
    const int64_t l_pac = B.l_pac;
    const uint64_t n = (uint64_t)l_pac * 2;
    B.seq_len = n;
    B.launches = 0; B.sort_pass_bytes = 0;
    if (n + 1 >= 0xffffffffull) {
        bsq_set_error("index build: text of %llu symbols needs 64-bit suffix ids; this build supports n + 1 < 2^32 "
                      "(the reference itself stops at 2^31 - 1, bioseqdb/bwa.cpp:10)", (unsigned long long)n);
        return BSQ_ERR;
    }
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

    BCHECK((prim::rs_prepare<uint64_t, uint32_t>()));
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
    k_find_primary<<<grid_for(n), 256, 0, st>>>(sa, n, d_primary); ++B.launches;
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
    k_bwt_blocks<<<grid_for((n + 15) / 16), 256, 0, st>>>(sa, T, n, B.primary, B.d_occ, blk_cnt, n_blocks + 1); ++B.launches;
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
