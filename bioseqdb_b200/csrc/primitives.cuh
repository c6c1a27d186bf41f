// primitives.cuh -- hand-written device-wide scan and LSD radix sort used by the index build
// (idx_sort).  HBM-stream bound: per radix pass each element is read twice (histogram,
// scatter) and written once; tiles are staged through shared memory so global writes leave the SM as
// per-digit runs.
#pragma once
#include "common.cuh"

namespace prim {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct OpSum { template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; } };
struct OpMax { template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a > b ? a : b; } };

template <class T, class Op> __device__ __forceinline__ T warp_inclusive(T v, Op op) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { T o = __shfl_up_sync(FULL, v, d); if (lane_id() >= d) v = op(v, o); }
    return v;
}

// block-wide inclusive scan of one value per thread; returns inclusive value, *total = block aggregate
template <class T, class Op> __device__ __forceinline__ T block_inclusive(T v, Op op, T* total, T* sh /* >= 32 */) {
    int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T inc = warp_inclusive(v, op);
    if (lane_id() == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane_id() < nw ? sh[lane_id()] : T(0);
        x = warp_inclusive(x, op);
        sh[lane_id()] = x;
    }
    __syncthreads();
    if (w > 0) inc = op(inc, sh[w - 1]);
    *total = sh[nw - 1];
    __syncthreads();
    return inc;
}

template <class T, class Op> __global__ void k_scan_reduce(const T* in, size_t n, T* sums, Op op) {
    __shared__ T sh[32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    T acc = T(0);
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) acc = op(acc, in[idx]);
    }
    T tot;
    block_inclusive(acc, op, &tot, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// inclusive==false: exclusive scan (identity 0).  offsets = scanned (exclusive) tile sums or nullptr.
template <class T, class Op, bool INCL> __global__ void k_scan_tile(const T* in, T* out, size_t n, const T* offsets, Op op) {
    __shared__ T sh[32];
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T acc = T(0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : T(0); acc = op(acc, v[i]); }
    T tot;
    T inc = block_inclusive(acc, op, &tot, sh);
    // exclusive prefix of this thread = inclusive of previous thread
    T prev = __shfl_up_sync(FULL, inc, 1);
    __shared__ T wl[32];
    if (lane_id() == 31) wl[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (lane_id() == 0) prev = (threadIdx.x >> 5) ? wl[(threadIdx.x >> 5) - 1] : T(0);
    T run = offsets ? op(offsets[blockIdx.x], prev) : prev;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        T nxt = op(run, v[i]);
        if (base + i < n) out[base + i] = INCL ? nxt : run;
        run = nxt;
    }
}

// Scan `n` elements of `in` into `out` (may alias).  `tmp` must hold scan_tmp_elems(n) elements.
inline size_t scan_tmp_elems(size_t n) {
    size_t tot = 0;
    while (n > 1) { n = (n + SCAN_TILE - 1) / SCAN_TILE; tot += n; if (n == 1) break; }
    return tot + 1;
}
template <class T, class Op, bool INCL> void device_scan(const T* in, T* out, size_t n, T* tmp, Op op, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return;
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        k_scan_tile<T, Op, INCL><<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, op);
        if (launches) ++*launches;
        return;
    }
    k_scan_reduce<T, Op><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, tmp, op);
    if (launches) ++*launches;
    device_scan<T, Op, false>(tmp, tmp, tiles, tmp + tiles, op, st, launches);
    k_scan_tile<T, Op, INCL><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, n, tmp, op);
    if (launches) ++*launches;
}

// ------------------------------------------------------------------ LSD radix sort, 8-bit digits
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 elements per CTA tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;

// Element source of a radix pass: plain (key, value) arrays, or keys/values computed from the element index
// (used by the wide index build to bucket 64-bit suffix ids by their leading symbols without materialising them).
template <class K, class V> struct SrcArrays {
    const K* k; const V* v;
    __device__ __forceinline__ K key(size_t i) const { return k[i]; }
    __device__ __forceinline__ V val(size_t i) const { return v[i]; }
};

// histogram: hist[bin * n_tiles + tile]; H = uint32_t while n < 2^32, unsigned long long beyond
template <class Src, class H> __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(Src src, size_t n, int shift, H* hist, uint32_t n_tiles) {
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(uint32_t)(src.key(idx) >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = (H)h[threadIdx.x];
}

template <class K, class V, class H> constexpr size_t rs_scatter_smem() {
    return sizeof(uint32_t) * (RS_WARPS * RS_BINS + RS_BINS + 32) + sizeof(H) * RS_BINS + (sizeof(K) + sizeof(V)) * RS_TILE;
}

template <class K, class V, class Src, class H> __global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(Src src, K* __restrict__ keys_out, V* __restrict__ vals_out, size_t n, int shift, const H* __restrict__ hist_scanned, uint32_t n_tiles) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint32_t (*wcount)[RS_BINS] = reinterpret_cast<uint32_t (*)[RS_BINS]>(rs_smem);  // per-warp digit counts, then bases
    uint32_t* tile_base = reinterpret_cast<uint32_t*>(rs_smem + sizeof(uint32_t) * RS_WARPS * RS_BINS);  // excl. scan over digits
    uint32_t* shs = tile_base + RS_BINS;
    H* glob_base = reinterpret_cast<H*>(shs + 32);
    K* skey = reinterpret_cast<K*>(glob_base + RS_BINS);
    V* sval = reinterpret_cast<V*>(skey + RS_TILE);
    const int w = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wcount[0][0])[i] = 0;
    glob_base[threadIdx.x] = hist_scanned[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    __syncthreads();
    const size_t tbase = (size_t)blockIdx.x * RS_TILE;
    const size_t wbase = tbase + (size_t)w * (32 * RS_ITEMS);
    K k[RS_ITEMS]; V v[RS_ITEMS]; uint32_t rk[RS_ITEMS];
    // each warp owns a contiguous 512-element sub-tile; round r covers [wbase + 32 r, +32)
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        size_t idx = wbase + (size_t)r * 32 + lane;
        bool ok = idx < n;
        k[r] = ok ? src.key(idx) : K(0);
        v[r] = ok ? src.val(idx) : V(0);
        uint32_t d = ok ? ((uint32_t)(k[r] >> shift) & 0xff) : 0x100u;  // invalid lanes form their own class
        uint32_t peers = __match_any_sync(FULL, d);
        uint32_t before = __popc(peers & ((1u << lane) - 1));
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) { old = wcount[w][d]; wcount[w][d] = old + __popc(peers); }
        old = __shfl_sync(FULL, old, leader);
        rk[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, tile total
    {
        uint32_t d = threadIdx.x, run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) { uint32_t c = wcount[ww][d]; wcount[ww][d] = run; run += c; }
        uint32_t tot;
        uint32_t inc = block_inclusive<uint32_t, OpSum>(run, OpSum(), &tot, shs);
        tile_base[d] = inc - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        size_t idx = wbase + (size_t)r * 32 + lane;
        if (idx < n) {
            uint32_t d = (uint32_t)(k[r] >> shift) & 0xff;
            uint32_t p = tile_base[d] + wcount[w][d] + rk[r];
            skey[p] = k[r]; sval[p] = v[r];
        }
    }
    __syncthreads();
    uint32_t cnt = (uint32_t)((n - tbase) < (size_t)RS_TILE ? (n - tbase) : (size_t)RS_TILE);
    for (uint32_t p = threadIdx.x; p < cnt; p += RS_THREADS) {
        K kk = skey[p];
        uint32_t d = (uint32_t)(kk >> shift) & 0xff;
        size_t g = (size_t)glob_base[d] + (p - tile_base[d]);
        if (keys_out) keys_out[g] = kk;
        vals_out[g] = sval[p];
    }
}

template <class K, class V, class Src, class H> cudaError_t rs_prepare() {  // opt in to > 48 KB dynamic shared memory
    return cudaFuncSetAttribute(k_rs_scatter<K, V, Src, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_scatter_smem<K, V, H>());
}
struct RadixWorkspace {
    uint32_t* hist = nullptr; uint32_t* scan_tmp = nullptr; size_t hist_elems = 0, tmp_elems = 0;
};
inline size_t rs_tiles(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts (keys, vals) by bits [begin_bit, end_bit) of the key; ping-pongs between (k0,v0) and (k1,v1).
// Returns 0 if the result is in (k0,v0), 1 if in (k1,v1).  n must be < 2^32.
template <class K, class V>
int radix_sort_pairs(K* k0, V* v0, K* k1, V* v1, size_t n, int begin_bit, int end_bit, RadixWorkspace& ws, cudaStream_t st,
                     uint64_t* launches, uint64_t* pass_bytes) {
    int cur = 0;
    if (n == 0) return 0;
    uint32_t tiles = (uint32_t)rs_tiles(n);
    using Src = SrcArrays<K, V>;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        K* ki = cur ? k1 : k0; V* vi = cur ? v1 : v0; K* ko = cur ? k0 : k1; V* vo = cur ? v0 : v1;
        Src src{ki, vi};
        k_rs_hist<Src, uint32_t><<<tiles, RS_THREADS, 0, st>>>(src, n, shift, ws.hist, tiles);
        if (launches) ++*launches;
        device_scan<uint32_t, OpSum, false>(ws.hist, ws.hist, (size_t)tiles * RS_BINS, ws.scan_tmp, OpSum(), st, launches);
        k_rs_scatter<K, V, Src, uint32_t><<<tiles, RS_THREADS, rs_scatter_smem<K, V, uint32_t>(), st>>>(src, ko, vo, n, shift, ws.hist, tiles);
        if (launches) ++*launches;
        if (pass_bytes) *pass_bytes += (uint64_t)n * (2 * sizeof(K) + sizeof(V) + sizeof(K) + sizeof(V));
        cur ^= 1;
    }
    return cur;
}

}  // namespace prim
