// seed.cu -- kernel `seed_smem`: the three seeding passes of mem_collect_intv, one warp
// per read, persistent warps with an atomic ticket.  All control flow is warp-uniform: every lane holds
// the same interval registers; the lanes only split the Occ block loads and popcounts (seed.cuh).
#include "seed.cuh"

namespace {

constexpr int SEED_THREADS = 128;

struct WarpLists { Intv* a; Intv* b; Intv* m; };

__device__ __forceinline__ void put(Intv* dst, const Intv& v) {
    if (lane_id() == 0) *dst = v;
}

// bwt_smem1a with max_intv == 0 (the only way this path calls it).  Results are appended (in order of
// increasing start) to out[*n_out ...] when their length >= min_seed_len.  Returns the next x.
__device__ int smem1(const DevIndex& ix, const DevOpts& o, int len, const uint8_t* q, int x, uint64_t min_intv, const WarpLists& L,
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
    // the list is consumed in reverse (longest match first): index it backwards instead of reversing it
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
                put(curr + n_curr, ok[c]);   // n_curr < n_prev <= list_cap
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
    // mem is in decreasing start order: append reversed, filtered by length
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

__device__ int seed_strategy1(const DevIndex& ix, int len, const uint8_t* q, int x, int min_len, uint64_t max_intv, Intv& mem, unsigned long long& n_ext) {
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

__global__ void __launch_bounds__(SEED_THREADS) seed_smem(SeedParams P, DevIndex ix, DevOpts o) {
    const int lane = lane_id();
    const uint32_t gwarp = (blockIdx.x * SEED_THREADS + threadIdx.x) >> 5;
    WarpLists L;
    L.a = P.scratch + (size_t)gwarp * 3 * P.list_cap;
    L.b = L.a + P.list_cap;
    L.m = L.b + P.list_cap;
    unsigned long long n_ext = 0;
    for (;;) {
        uint32_t r = next_ticket(P.ticket);
        if (r >= P.n_reads) break;
        const uint8_t* q = P.seqs + P.offs[r];
        const int len = (int)(P.offs[r + 1] - P.offs[r]);
        Intv* out = P.out + (size_t)r * P.cap;
        uint32_t n_out = 0; bool ovf = false;
        if (len >= o.min_seed_len) {   // mem_chain returns before seeding otherwise
            // pass 1
            int x = 0;
            while (x < len) {
                if (q[x] < 4) x = smem1(ix, o, len, q, x, 1, L, P.list_cap, out, n_out, P.cap, ovf, n_ext);
                else ++x;
            }
            // pass 2: re-seeding
            uint32_t old_n = n_out < P.cap ? n_out : P.cap;
            for (uint32_t k = 0; k < old_n; ++k) {
                Intv p = out[k];
                int start = (int)(p.info >> 32), end = (int)(uint32_t)p.info;
                if (end - start < o.split_len || p.x2 > (uint64_t)o.split_width) continue;
                smem1(ix, o, len, q, (start + end) >> 1, p.x2 + 1, L, P.list_cap, out, n_out, P.cap, ovf, n_ext);
            }
            // pass 3
            if (o.max_mem_intv > 0) {
                x = 0;
                while (x < len) {
                    if (q[x] < 4) {
                        Intv m;
                        x = seed_strategy1(ix, len, q, x, o.min_seed_len, (uint64_t)o.max_mem_intv, m, n_ext);
                        if (m.x2 > 0) { if (n_out < P.cap) put(out + n_out, m); else ovf = true; ++n_out; }
                    } else ++x;
                }
            }
            __syncwarp();
        }
        if (ovf || n_out > P.cap) { if (lane == 0) atomicExch(P.overflow, 1u); n_out = n_out < P.cap ? n_out : P.cap; }
        // sort by info (ties are bit-identical records, so any correct sort equals ks_introsort's result)
        if (n_out > 1) {
            Intv* tmp = L.a;
            if (n_out <= P.list_cap) {
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
            } else if (lane == 0) atomicExch(P.overflow, 1u);
        }
        if (lane == 0) P.out_cnt[r] = n_out;
    }
    if (P.n_extend && lane == 0 && n_ext) atomicAdd(P.n_extend, n_ext);
}

}  // namespace

int seed_resident_warps() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, seed_smem, SEED_THREADS, 0);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (nb < 1) nb = 1;
    return nb * sms * (SEED_THREADS / 32);
}

void launch_seed(const SeedParams& p, const DevIndex& ix, const DevOpts& o, cudaStream_t st, int* n_warps_out) {
    int warps = seed_resident_warps();
    int blocks = warps / (SEED_THREADS / 32);
    if (n_warps_out) *n_warps_out = warps;
    seed_smem<<<blocks, SEED_THREADS, 0, st>>>(p, ix, o);
}
