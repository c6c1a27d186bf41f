// seed.cu -- kernel `seed_smem`: the three seeding passes of mem_collect_intv, one warp
// per read, persistent warps with an atomic ticket.  All control flow is warp-uniform: every lane holds
// the same interval registers; the lanes split the two 64-byte Occ blocks of a bwt_extend (one coalesced
// warp load: lanes 0-15 -> block of row k, lanes 16-31 -> block of row l).
//
// Fast path (text shorter than 2^32 rows): only the child interval of the ONE base being extended is ever
// needed, and it is three sums over the 32 lanes --
//     x2' = sum_l contrib_c - sum_k contrib_c,   S = sum over symbols > c of the same,   tk[c] = sum_k contrib_c
// where a lane's contribution is its count word (lanes holding the checkpoint of symbol c) or the popcount
// of its masked symbol word.  Each sum is ONE redux.sync (warp vote/reduce hardware), so a bwt_extend costs
// one load, two popcounts and three reductions per lane.  Interval lists live in shared memory.
// The wide path (>= 2^32 rows) keeps the shuffle-based occ4_pair of seed.cuh.
#include "seed.cuh"

namespace {

constexpr int SEED_THREADS = 256;
constexpr int SEED_WARPS = SEED_THREADS / 32;

struct Iv32 { uint32_t x0, x1, x2, info; };   // info = end (forward list) ; start<<16|end packing is NOT used: lists keep end only

__device__ __forceinline__ uint32_t occ_word(const DevIndex& ix, uint32_t pos_k, uint32_t pos_l, uint32_t& pos_adj) {
    const int lane = lane_id();
    uint32_t pos = (lane & 16) ? pos_l : pos_k;
    pos -= (pos >= (uint32_t)ix.primary);
    pos_adj = pos;
    return __ldg(ix.occ + ((size_t)(pos >> 7) << 4) + (lane & 15));
}

// child interval for base c of the parent (xo = x[!is_back], xb = x[is_back], x2): see file header
__device__ __forceinline__ void occ_reduce(const DevIndex& ix, uint32_t word, uint32_t pos, int c, uint32_t xo, uint32_t xb, uint32_t x2,
                                           uint32_t& no, uint32_t& nb, uint32_t& nsz) {
    const int lane = lane_id();
    const int idx = lane & 15;
    int eq, gt;
    if (idx < 8) {
        const bool low = !(idx & 1);
        const int cc = idx >> 1;
        eq = (low && cc == c) ? (int)word : 0;
        gt = (low && cc > c) ? (int)word : 0;
    } else {
        int nsym = (int)(pos & 127) + 1 - ((idx - 8) << 4);
        nsym = nsym < 0 ? 0 : (nsym > 16 ? 16 : nsym);
        const uint32_t keep = nsym ? (0x55555555u & (0xffffffffu << (32 - 2 * nsym))) : 0u;
        const uint32_t lo = word & keep, hi = (word >> 1) & keep;      // keep is on the 0x5555 lattice
        const uint32_t nlo = ~word & keep, nhi = ~(word >> 1) & keep;
        const uint32_t mh = (c & 2) ? hi : nhi, ml = (c & 1) ? lo : nlo;
        eq = __popc(mh & ml);
        const uint32_t g = c == 0 ? (hi | lo) : (c == 1 ? hi : (c == 2 ? (hi & lo) : 0u));
        gt = __popc(g);
    }
    const bool lhalf = (lane & 16) != 0;
    const int sz = __reduce_add_sync(FULL, lhalf ? eq : -eq);
    const int S = __reduce_add_sync(FULL, lhalf ? gt : -gt);
    const int tk = __reduce_add_sync(FULL, lhalf ? 0 : eq);
    no = (uint32_t)ix.L2[c] + 1u + (uint32_t)tk;
    nsz = (uint32_t)sz;
    nb = xb + (uint32_t)(xo <= (uint32_t)ix.primary && xo + x2 - 1 >= (uint32_t)ix.primary) + (uint32_t)S;
}

template <int IS_BACK>
__device__ __forceinline__ Iv32 extend32(const DevIndex& ix, const Iv32& ik, int c) {
    const uint32_t xo = IS_BACK ? ik.x0 : ik.x1, xb = IS_BACK ? ik.x1 : ik.x0;
    uint32_t pos;
    const uint32_t word = occ_word(ix, xo - 1, xo - 1 + ik.x2, pos);
    uint32_t no, nb, nsz;
    occ_reduce(ix, word, pos, c, xo, xb, ik.x2, no, nb, nsz);
    Iv32 ok;
    ok.x0 = IS_BACK ? no : nb; ok.x1 = IS_BACK ? nb : no; ok.x2 = nsz; ok.info = ik.info;
    return ok;
}

__device__ __forceinline__ Iv32 set_intv32(const DevIndex& ix, int c) {
    Iv32 ik;
    ik.x0 = (uint32_t)ix.L2[c] + 1; ik.x1 = (uint32_t)ix.L2[3 - c] + 1; ik.x2 = (uint32_t)(ix.L2[c + 1] - ix.L2[c]); ik.info = 0;
    return ik;
}

struct Out { Intv* out; uint32_t n, cap; bool ovf; };

__device__ __forceinline__ void emit(Out& O, const Iv32& p, uint32_t start, uint32_t end) {
    if (O.n < O.cap) {
        if (lane_id() == 0) { Intv v; v.x0 = p.x0; v.x1 = p.x1; v.x2 = p.x2; v.info = (uint64_t)start << 32 | end; O.out[O.n] = v; }
    } else O.ovf = true;
    ++O.n;
}

// bwt_smem1a with max_intv == 0.  la/lb: two interval lists of list_cap entries (shared memory).
__device__ int smem1_32(const DevIndex& ix, const DevOpts& o, int len, const uint8_t* q, int x, uint32_t min_intv, Iv32* la, Iv32* lb,
                        uint32_t list_cap, Out& O, unsigned long long& n_ext) {
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    const int lane = lane_id();
    Iv32 ik = set_intv32(ix, q[x]);
    ik.info = (uint32_t)(x + 1);
    Iv32* curr = la; Iv32* prev = lb;
    uint32_t n_curr = 0;
    int i;
    for (i = x + 1; i < len; ++i) {
        const int b = q[i];
        if (b < 4) {
            Iv32 ok = extend32<0>(ix, ik, 3 - b); ++n_ext;
            if (ok.x2 != ik.x2) {
                if (n_curr < list_cap) { if (lane == 0) curr[n_curr] = ik; } else O.ovf = true;
                ++n_curr;
                if (ok.x2 < min_intv) break;
            }
            ik = ok; ik.info = (uint32_t)(i + 1);
        } else {
            if (n_curr < list_cap) { if (lane == 0) curr[n_curr] = ik; } else O.ovf = true;
            ++n_curr;
            break;
        }
    }
    if (i == len) { if (n_curr < list_cap) { if (lane == 0) curr[n_curr] = ik; } else O.ovf = true; ++n_curr; }
    if (n_curr > list_cap) n_curr = list_cap;
    __syncwarp();
    const int ret = (int)curr[n_curr - 1].info;     // the list is consumed backwards (longest match first)
    { Iv32* t = curr; curr = prev; prev = t; }
    uint32_t n_prev = n_curr; bool reversed = true;
    const uint32_t out_first = O.n;
    bool have_mem = false; uint32_t last_mem_start = 0;
    for (i = x - 1; i >= -1; --i) {
        const int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
        n_curr = 0;
        uint32_t last_x2 = 0;
        for (uint32_t j = 0; j < n_prev; ++j) {
            const Iv32 p = prev[reversed ? n_prev - 1 - j : j];
            Iv32 ok; ok.x2 = 0;
            if (c >= 0) { ok = extend32<1>(ix, p, c); ++n_ext; }
            if (c < 0 || ok.x2 < min_intv) {
                if (n_curr == 0) {
                    if (!have_mem || (uint32_t)(i + 1) < last_mem_start) {
                        const int slen = (int)p.info - (i + 1);
                        if (slen >= o.min_seed_len) emit(O, p, (uint32_t)(i + 1), p.info);
                        have_mem = true; last_mem_start = (uint32_t)(i + 1);
                    }
                }
            } else if (n_curr == 0 || ok.x2 != last_x2) {
                if (lane == 0) curr[n_curr] = ok;    // n_curr < n_prev <= list_cap
                ++n_curr; last_x2 = ok.x2;
            }
        }
        if (n_curr == 0) break;
        __syncwarp();
        { Iv32* t = curr; curr = prev; prev = t; }
        n_prev = n_curr; reversed = false;
    }
    // emitted in decreasing start order: reverse the appended segment
    __syncwarp();
    {
        const uint32_t hi = O.n < O.cap ? O.n : O.cap;
        if (hi > out_first + 1) {
            const uint32_t cnt = hi - out_first;
            for (uint32_t k = lane; k < cnt / 2; k += 32) {
                Intv a = O.out[out_first + k], b = O.out[hi - 1 - k];
                O.out[out_first + k] = b; O.out[hi - 1 - k] = a;
            }
            __syncwarp();
        }
    }
    return ret;
}

__device__ int seed_strategy1_32(const DevIndex& ix, int len, const uint8_t* q, int x, int min_len, uint32_t max_intv, Out& O, unsigned long long& n_ext) {
    if (q[x] > 3) return x + 1;
    Iv32 ik = set_intv32(ix, q[x]);
    for (int i = x + 1; i < len; ++i) {
        const int b = q[i];
        if (b < 4) {
            Iv32 ok = extend32<0>(ix, ik, 3 - b); ++n_ext;
            if (ok.x2 < max_intv && i - x >= min_len) {
                if (ok.x2 > 0) emit(O, ok, (uint32_t)x, (uint32_t)(i + 1));
                return i + 1;
            }
            ik = ok;
        } else return i + 1;
    }
    return len;
}

// ---------------------------------------------------------------- wide path (>= 2^32 rows): 64-bit records
struct WarpLists { Intv* a; Intv* b; Intv* m; };
__device__ __forceinline__ void put(Intv* dst, const Intv& v) { if (lane_id() == 0) *dst = v; }

__device__ int smem1_64(const DevIndex& ix, const DevOpts& o, int len, const uint8_t* q, int x, uint64_t min_intv, const WarpLists& L,
                        uint32_t list_cap, Intv* out, uint32_t& n_out, uint32_t cap, bool& ovf, unsigned long long& n_ext) {
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    Intv ik, ok[4];
    bwt_set_intv(ix, q[x], ik);
    ik.info = (uint64_t)(x + 1);
    Intv* curr = L.a; Intv* prev = L.b;
    uint32_t n_curr = 0;
    int i;
    for (i = x + 1; i < len; ++i) {
        int b = q[i];
        if (b < 4) {
            int c = 3 - b;
            bwt_extend<0>(ix, ik, ok); ++n_ext;
            if (ok[c].x2 != ik.x2) {
                if (n_curr < list_cap) put(curr + n_curr, ik); else ovf = true;
                ++n_curr;
                if (ok[c].x2 < min_intv) break;
            }
            ik = ok[c]; ik.info = (uint64_t)(i + 1);
        } else {
            if (n_curr < list_cap) put(curr + n_curr, ik); else ovf = true;
            ++n_curr;
            break;
        }
    }
    if (i == len) { if (n_curr < list_cap) put(curr + n_curr, ik); else ovf = true; ++n_curr; }
    if (n_curr > list_cap) n_curr = list_cap;
    __syncwarp();
    int ret = (int)(uint32_t)curr[n_curr - 1].info;
    { Intv* t = curr; curr = prev; prev = t; }
    uint32_t n_prev = n_curr; bool prev_reversed = true;
    uint32_t n_mem = 0; uint64_t last_mem_start = 0;
    for (i = x - 1; i >= -1; --i) {
        int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
        n_curr = 0;
        uint64_t last_x2 = 0;
        for (uint32_t j = 0; j < n_prev; ++j) {
            Intv p = prev[prev_reversed ? n_prev - 1 - j : j];
            if (c >= 0) { bwt_extend<1>(ix, p, ok); ++n_ext; }
            if (c < 0 || ok[c].x2 < min_intv) {
                if (n_curr == 0) {
                    if (n_mem == 0 || (uint64_t)(i + 1) < last_mem_start) {
                        p.info |= (uint64_t)(i + 1) << 32;
                        if (n_mem < list_cap) put(L.m + n_mem, p); else ovf = true;
                        ++n_mem; last_mem_start = (uint64_t)(i + 1);
                    }
                }
            } else if (n_curr == 0 || ok[c].x2 != last_x2) {
                ok[c].info = p.info;
                put(curr + n_curr, ok[c]);
                ++n_curr; last_x2 = ok[c].x2;
            }
        }
        if (n_curr == 0) break;
        __syncwarp();
        { Intv* t = curr; curr = prev; prev = t; }
        n_prev = n_curr; prev_reversed = false;
    }
    if (n_mem > list_cap) n_mem = list_cap;
    __syncwarp();
    for (uint32_t k = 0; k < n_mem; ++k) {
        Intv p = L.m[n_mem - 1 - k];
        int slen = (int)((uint32_t)p.info - (uint32_t)(p.info >> 32));
        if (slen >= o.min_seed_len) {
            if (n_out < cap) put(out + n_out, p); else ovf = true;
            ++n_out;
        }
    }
    __syncwarp();
    return ret;
}

__device__ int seed_strategy1_64(const DevIndex& ix, int len, const uint8_t* q, int x, int min_len, uint64_t max_intv, Intv& mem, unsigned long long& n_ext) {
    mem.x0 = mem.x1 = mem.x2 = mem.info = 0;
    if (q[x] > 3) return x + 1;
    Intv ik, ok[4];
    bwt_set_intv(ix, q[x], ik);
    for (int i = x + 1; i < len; ++i) {
        int b = q[i];
        if (b < 4) {
            int c = 3 - b;
            bwt_extend<0>(ix, ik, ok); ++n_ext;
            if (ok[c].x2 < max_intv && i - x >= min_len) {
                mem = ok[c];
                mem.info = (uint64_t)x << 32 | (uint64_t)(i + 1);
                return i + 1;
            }
            ik = ok[c];
        } else return i + 1;
    }
    return len;
}

// sort a read's intervals by info (ties are bit-identical records, so any correct sort equals ks_introsort's result)
__device__ void sort_by_info(Intv* out, uint32_t n_out, Intv* tmp, uint32_t tmp_cap, uint32_t* overflow) {
    const int lane = lane_id();
    if (n_out <= 1) return;
    if (n_out <= 32) {   // one record per lane, ranks by shuffle
        Intv me; me.x0 = me.x1 = me.x2 = 0; me.info = ~0ull;
        if ((uint32_t)lane < n_out) me = out[lane];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n_out; ++j) {
            uint64_t oi = __shfl_sync(FULL, me.info, (int)j);
            rank += (oi < me.info) || (oi == me.info && j < (uint32_t)lane);
        }
        __syncwarp();
        if ((uint32_t)lane < n_out) out[rank] = me;
        __syncwarp();
        return;
    }
    if (n_out > tmp_cap) { if (lane == 0) atomicExch(overflow, 1u); return; }
    for (uint32_t base = 0; base < n_out; base += 32) {
        uint32_t k = base + lane;
        if (k < n_out) {
            Intv me = out[k];
            uint32_t rank = 0;
            for (uint32_t j = 0; j < n_out; ++j) {
                uint64_t oi = out[j].info;
                rank += (oi < me.info) || (oi == me.info && j < k);
            }
            tmp[rank] = me;
        }
    }
    __syncwarp();
    for (uint32_t k = lane; k < n_out; k += 32) out[k] = tmp[k];
    __syncwarp();
}

template <bool WIDE>
__global__ void __launch_bounds__(SEED_THREADS) seed_smem(SeedParams P, DevIndex ix, DevOpts o) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const int lane = lane_id();
    const uint32_t gwarp = (blockIdx.x * SEED_THREADS + threadIdx.x) >> 5;
    Intv* gl = P.scratch + (size_t)gwarp * 3 * P.list_cap;     // global scratch: wide-path lists, big-sort buffer
    unsigned long long n_ext = 0;
    Iv32* la = nullptr; Iv32* lb = nullptr;
    if (!WIDE) {
        if (P.lists_in_smem) { la = reinterpret_cast<Iv32*>(dyn_smem) + (size_t)(threadIdx.x >> 5) * 2 * P.list_cap; lb = la + P.list_cap; }
        else { la = reinterpret_cast<Iv32*>(gl); lb = la + P.list_cap; }
    }
    for (;;) {
        uint32_t r = next_ticket(P.ticket);
        if (r >= P.n_reads) break;
        const uint8_t* q = P.seqs + P.offs[r];
        const int len = (int)(P.offs[r + 1] - P.offs[r]);
        Intv* out = P.out + (size_t)r * P.cap;
        uint32_t n_out = 0; bool ovf = false;
        if (len >= o.min_seed_len) {   // mem_chain returns before seeding otherwise
            if (!WIDE) {
                Out O; O.out = out; O.n = 0; O.cap = P.cap; O.ovf = false;
                int x = 0;
                while (x < len) {      // pass 1: all SMEMs
                    if (q[x] < 4) x = smem1_32(ix, o, len, q, x, 1, la, lb, P.list_cap, O, n_ext);
                    else ++x;
                }
                const uint32_t old_n = O.n < O.cap ? O.n : O.cap;
                for (uint32_t k = 0; k < old_n; ++k) {   // pass 2: re-seeding inside long, rare SMEMs
                    const Intv p = out[k];
                    const int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
                    if (end - start < o.split_len || p.x2 > (uint64_t)o.split_width) continue;
                    smem1_32(ix, o, len, q, (start + end) >> 1, (uint32_t)p.x2 + 1, la, lb, P.list_cap, O, n_ext);
                }
                if (o.max_mem_intv > 0) {                // pass 3: LAST-like
                    x = 0;
                    while (x < len) {
                        if (q[x] < 4) x = seed_strategy1_32(ix, len, q, x, o.min_seed_len, (uint32_t)o.max_mem_intv, O, n_ext);
                        else ++x;
                    }
                }
                n_out = O.n; ovf = O.ovf;
            } else {
                WarpLists L; L.a = gl; L.b = gl + P.list_cap; L.m = L.b + P.list_cap;
                int x = 0;
                while (x < len) {
                    if (q[x] < 4) x = smem1_64(ix, o, len, q, x, 1, L, P.list_cap, out, n_out, P.cap, ovf, n_ext);
                    else ++x;
                }
                uint32_t old_n = n_out < P.cap ? n_out : P.cap;
                for (uint32_t k = 0; k < old_n; ++k) {
                    Intv p = out[k];
                    int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
                    if (end - start < o.split_len || p.x2 > (uint64_t)o.split_width) continue;
                    smem1_64(ix, o, len, q, (start + end) >> 1, p.x2 + 1, L, P.list_cap, out, n_out, P.cap, ovf, n_ext);
                }
                if (o.max_mem_intv > 0) {
                    x = 0;
                    while (x < len) {
                        if (q[x] < 4) {
                            Intv m;
                            x = seed_strategy1_64(ix, len, q, x, o.min_seed_len, (uint64_t)o.max_mem_intv, m, n_ext);
                            if (m.x2 > 0) { if (n_out < P.cap) put(out + n_out, m); else ovf = true; ++n_out; }
                        } else ++x;
                    }
                }
            }
            __syncwarp();
        }
        if (ovf || n_out > P.cap) { if (lane == 0) atomicExch(P.overflow, 1u); n_out = n_out < P.cap ? n_out : P.cap; }
        sort_by_info(out, n_out, gl, 3 * P.list_cap, P.overflow);
        if (lane == 0) P.out_cnt[r] = n_out;
    }
    if (P.n_extend && lane == 0 && n_ext) atomicAdd(P.n_extend, n_ext);
}

}  // namespace

static size_t seed_smem_bytes(const SeedParams& p) { return p.lists_in_smem ? (size_t)SEED_WARPS * 2 * p.list_cap * sizeof(Iv32) : 0; }

int seed_resident_warps() {
    // upper bound used to size the per-warp global scratch: 64 warps per SM
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 64 * sms;
}

bool seed_lists_fit_smem(uint32_t list_cap) { return (size_t)SEED_WARPS * 2 * list_cap * sizeof(Iv32) <= 48 * 1024; }

void launch_seed(const SeedParams& p, const DevIndex& ix, const DevOpts& o, cudaStream_t st, int* n_warps_out) {
    int dev = 0, sms = 148, nb = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool wide = ix.seq_len + 1 >= 0xffffffffull;
    const size_t smem = wide ? 0 : seed_smem_bytes(p);
    if (wide) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, seed_smem<true>, SEED_THREADS, smem);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, seed_smem<false>, SEED_THREADS, smem);
    if (nb < 1) nb = 1;
    if (nb * SEED_WARPS > 64) nb = 64 / SEED_WARPS;
    if (n_warps_out) *n_warps_out = nb * sms * SEED_WARPS;
    if (wide) seed_smem<true><<<nb * sms, SEED_THREADS, smem, st>>>(p, ix, o);
    else seed_smem<false><<<nb * sms, SEED_THREADS, smem, st>>>(p, ix, o);
}
