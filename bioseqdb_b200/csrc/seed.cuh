// seed.cuh -- batched SMEM seeding by FM-index backward/forward search (replaces
// libbwa mem_collect_intv / bwt_smem1a / bwt_seed_strategy1 / bwt_extend reached from reference
// bioseqdb/bwa.cpp:149).  One warp cooperates per read: the two 64-byte Occ blocks of a bwt_extend are
// fetched by one coalesced warp load (lanes 0-15 -> block of k, lanes 16-31 -> block of l), symbol words
// are popcounted per lane and combined with shuffles.  Bound by random 64-byte HBM/L2 reads.
#pragma once
#include "common.cuh"

// Occ4 at rows k and l (n+1 row space; (uint64_t)-1 => zeros), all lanes receive all 8 values.
__device__ __forceinline__ void occ4_pair(const DevIndex& ix, uint64_t k, uint64_t l, uint64_t tk[4], uint64_t tl[4]) {
    const int lane = lane_id();
    const int half = lane >> 4, idx = lane & 15;
    uint64_t pos = half ? l : k;
    const bool none = pos == (uint64_t)-1;
    pos -= (pos >= ix.primary) && !none;
    uint32_t word = 0;
    if (!none) word = __ldg(ix.occ + ((pos >> 7) << 4) + idx);
    uint32_t packed = 0;
    if (idx >= 8 && !none) {
        int nsym = (int)(pos & 127) + 1 - ((idx - 8) << 4);
        nsym = nsym < 0 ? 0 : (nsym > 16 ? 16 : nsym);
        uint32_t keep = nsym ? (0x55555555u & (0xffffffffu << (32 - 2 * nsym))) : 0u;
        uint32_t lo = word & 0x55555555u, hi = (word >> 1) & 0x55555555u;
        packed = __popc(~hi & ~lo & keep) | __popc(~hi & lo & keep) << 8 | __popc(hi & ~lo & keep) << 16 | __popc(hi & lo & keep) << 24;
    }
    packed += __shfl_xor_sync(FULL, packed, 1);
    packed += __shfl_xor_sync(FULL, packed, 2);
    packed += __shfl_xor_sync(FULL, packed, 4);  // lanes 8-15 (and 24-31) now hold the block's in-word counts
    uint32_t pk = __shfl_sync(FULL, packed, 8), pl = __shfl_sync(FULL, packed, 24);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t klo = __shfl_sync(FULL, word, 2 * c), khi = __shfl_sync(FULL, word, 2 * c + 1);
        uint32_t llo = __shfl_sync(FULL, word, 16 + 2 * c), lhi = __shfl_sync(FULL, word, 17 + 2 * c);
        tk[c] = ((uint64_t)khi << 32 | klo) + ((pk >> (8 * c)) & 0xff);
        tl[c] = ((uint64_t)lhi << 32 | llo) + ((pl >> (8 * c)) & 0xff);
    }
}

__device__ __forceinline__ void bwt_set_intv(const DevIndex& ix, int c, Intv& ik) {
    ik.x0 = ix.L2[c] + 1; ik.x1 = ix.L2[3 - c] + 1; ik.x2 = ix.L2[c + 1] - ix.L2[c]; ik.info = 0;
}

// bwt_extend.  IS_BACK selects which of x0/x1 plays "k".
template <int IS_BACK>
__device__ __forceinline__ void bwt_extend(const DevIndex& ix, const Intv& ik, Intv ok[4]) {
    uint64_t tk[4], tl[4];
    const uint64_t xo = IS_BACK ? ik.x0 : ik.x1;   // x[!is_back]
    const uint64_t xb = IS_BACK ? ik.x1 : ik.x0;   // x[is_back]
    occ4_pair(ix, xo - 1, xo - 1 + ik.x2, tk, tl);
    uint64_t no[4], nb[4], sz[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { no[i] = ix.L2[i] + 1 + tk[i]; sz[i] = tl[i] - tk[i]; }
    nb[3] = xb + (xo <= ix.primary && xo + ik.x2 - 1 >= ix.primary);
    nb[2] = nb[3] + sz[3];
    nb[1] = nb[2] + sz[2];
    nb[0] = nb[1] + sz[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ok[i].x0 = IS_BACK ? no[i] : nb[i];
        ok[i].x1 = IS_BACK ? nb[i] : no[i];
        ok[i].x2 = sz[i];
    }
}

struct SeedParams {
    const uint8_t* seqs;      // nt4 codes, concatenated
    const uint64_t* offs;     // n + 1
    uint32_t n_reads;
    Intv* out;                // n_reads x cap
    uint32_t* out_cnt;        // n_reads
    uint32_t cap;
    Intv* scratch;            // per resident warp: 3 lists of list_cap entries
    uint32_t list_cap;
    int lists_in_smem;        // narrow path: the two interval lists of a warp and the read live in shared memory
    uint32_t read_cap;        // bytes reserved per warp for the staged read (>= longest read, multiple of 16)
    uint32_t* ticket;
    uint32_t* overflow;       // set to 1 when a read needs more than cap intervals
    unsigned long long* n_extend;  // optional counter (roofline units); nullptr in production
};
void launch_seed(const SeedParams& p, const DevIndex& ix, const DevOpts& o, cudaStream_t st, int* n_warps_out);
int seed_resident_warps();
bool seed_lists_fit_smem(uint32_t list_cap, uint32_t read_cap);
