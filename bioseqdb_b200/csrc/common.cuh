// common.cuh -- shared device-side views and helpers for the sm_100a read-alignment kernels.
// Semantics follow SURVEY.md Appendix A (libbwa as called from reference bioseqdb/bwa.cpp).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define FULL 0xffffffffu
#define WARP 32

// mem_opt_t fields the path reads; filled on the host by opts_fill().
struct DevOpts {
    int a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop, min_seed_len, max_occ;
    int max_mem_intv, split_width, split_len, max_chain_gap, max_chain_extend, min_chain_weight;
    float mask_level, drop_ratio, mask_level_redun;
    int mat[25];
    int mat_max;
};

// FM-index + reference as laid out in HBM.
//  occ : bwa's interleaved layout -- per 128 BWT symbols one 64-byte block {4 x u64 counts, 8 x u32 of
//        16 symbols MSB-first}; one trailing count record.  64-byte aligned.
//  sa  : the FULL suffix array of T$ (n+1 rows; row 0 = $), 4 bytes per row when n+1 < 2^32 else 8.
//        bwt_sa(k) == SA[k] for every k >= 1 (SURVEY A.3 allows any layout returning the same value).
//  pac : forward 2-bit text, byte-rounded row concatenation (reference bwa.cpp:95-97).
struct DevIndex {
    const uint8_t* pac;
    const uint32_t* occ;
    const void* sa;
    const int64_t* ann_offset;
    const int32_t* ann_len;
    int64_t l_pac;
    uint64_t seq_len, primary;
    uint64_t L2[5];
    int n_anns;
    int sa_bytes;
};

struct Intv { uint64_t x0, x1, x2, info; };

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t pac_get(const uint8_t* pac, int64_t i) {
    return (pac[i >> 2] >> ((~i & 3) << 1)) & 3u;
}
// base at position p of the doubled coordinate [0, 2 l_pac)
__device__ __forceinline__ uint32_t ref_base(const DevIndex& ix, int64_t p) {
    return p < ix.l_pac ? pac_get(ix.pac, p) : 3u - pac_get(ix.pac, (ix.l_pac << 1) - 1 - p);
}
__device__ __forceinline__ int64_t bns_depos(const DevIndex& ix, int64_t pos, int* is_rev) {
    return (*is_rev = (pos >= ix.l_pac)) ? (ix.l_pac << 1) - 1 - pos : pos;
}
__device__ __forceinline__ int bns_pos2rid(const DevIndex& ix, int64_t pos_f) {
    int left = 0, mid = 0, right = ix.n_anns;
    if (pos_f >= ix.l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= ix.ann_offset[mid]) {
            if (mid == ix.n_anns - 1) break;
            if (pos_f < ix.ann_offset[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
__device__ __forceinline__ int bns_intv2rid(const DevIndex& ix, int64_t rb, int64_t re) {
    int is_rev;
    if (rb < ix.l_pac && re > ix.l_pac) return -2;
    const int rid_b = bns_pos2rid(ix, bns_depos(ix, rb, &is_rev));
    if (rb >= re) return rid_b;
    // the other end lies in the same sequence iff it is inside rid_b's [offset, next offset): no second search
    const int64_t pe = bns_depos(ix, re - 1, &is_rev);
    const bool same = pe >= ix.ann_offset[rid_b] && (rid_b == ix.n_anns - 1 || pe < ix.ann_offset[rid_b + 1]);
    return same ? rid_b : -1;
}
__device__ __forceinline__ uint64_t sa_at(const DevIndex& ix, uint64_t k) {
    return ix.sa_bytes == 4 ? (uint64_t)((const uint32_t*)ix.sa)[k] : ((const uint64_t*)ix.sa)[k];
}

__device__ __forceinline__ int cal_max_gap(const DevOpts& o, int qlen) {
    int l_del = (int)((double)(qlen * o.a - o.o_del) / o.e_del + 1.);
    int l_ins = (int)((double)(qlen * o.a - o.o_ins) / o.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < o.w << 1 ? l : o.w << 1;
}

// persistent-warp work distribution: one atomic ticket per unit
__device__ __forceinline__ uint32_t next_ticket(uint32_t* counter) {
    uint32_t t = 0;
    if (lane_id() == 0) t = atomicAdd(counter, 1u);
    return __shfl_sync(FULL, t, 0);
}

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bsq_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return BSQ_ERR; } } while (0)
#define CUDA_CHECK_NULL(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bsq_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return nullptr; } } while (0)
void bsq_set_error(const char* fmt, ...);
