// extend.cu -- kernel `sw_extend`: mem_chain2aln with the banded Smith-Waterman extension
// ksw_extend2 (A.8).  Replaces libbwa mem_chain2aln / cal_max_gap / ksw_extend2 reached from reference
// bioseqdb/bwa.cpp:149.  One warp per read walks its chains and seeds in the reference order (the skip
// rule depends on the regions produced so far), and runs every left/right extension cooperatively with
// ksw_extend_warp (row sweep, F by shuffle scan, DPX max instructions).  Integer-pipe bound.
#include "pipeline.cuh"
#include "ksw_warp.cuh"
#include "ksort_dev.cuh"
#include "launch_cache.cuh"

namespace {

constexpr int EXT_THREADS = 128;
constexpr int EXT_WARPS = EXT_THREADS / 32;
constexpr int MAX_BAND_TRY = 2;

struct Scratch { uint8_t* rseq; uint8_t* query; int* ehh; int* ehe; };

__device__ __forceinline__ Scratch carve(uint8_t* base, uint32_t max_len, uint32_t rseq_cap) {
    Scratch s;
    s.ehh = reinterpret_cast<int*>(base);
    s.ehe = s.ehh + (max_len + 2);
    s.rseq = reinterpret_cast<uint8_t*>(s.ehe + (max_len + 2));
    s.query = s.rseq + rseq_cap;
    return s;
}

// mem_chain2aln for one read.  WARPMODE: the 32 lanes of a warp work on the read together (scalar decisions are taken
// redundantly by every lane, lane 0 writes) and run the ksw_extend2 calls the pre-pass did not answer.  !WARPMODE: ONE thread
// walks the read with the same code; a ksw_extend2 call without a memo makes it give up (returns false) and the read is
// queued for the warp kernel, which starts it again from scratch.
template <bool WARPMODE, bool SMEM>
__device__ __forceinline__ bool extend_read(const ExtendParams& P, const DevIndex& ix, const DevOpts& o, const uint32_t r, const Scratch& S, const int* smat,
                                            unsigned long long& cells, unsigned long long& calls, unsigned long long& rows) {
    const int lane = WARPMODE ? lane_id() : 0;
    const int64_t l_pac = ix.l_pac;
    {
        const ReadBlock blk = P.blocks[r];
        int n_reg = 0;
        if (blk.n_chains == 0) { if (lane == 0) P.reg_cnt[r] = 0; return true; }
        const int l_query = (int)(P.offs[r + 1] - P.offs[r]);
        if (WARPMODE) {
            const uint8_t* qg = P.seqs + P.offs[r];
            for (int i = lane; i < l_query; i += 32) S.query[i] = qg[i];
        }
        const uint8_t* query = S.query;
        RegRec* av = P.regs + blk.base;
        if (WARPMODE) __syncwarp();
        for (uint32_t ci = 0; ci < blk.n_chains; ++ci) {
            const ChainRec c = P.chains[blk.base + ci];
            const SeedRec* seeds = P.seeds + blk.base + c.seed_off;
            uint64_t* srt = P.srt + blk.base + c.seed_off;
            const int n = c.n_seeds;
            if (n == 0) continue;
            // ---- maximal span of the chain
            int64_t rmax0 = l_pac << 1, rmax1 = 0;
            for (int i = 0; i < n; ++i) {
                const SeedRec t = seeds[i];
                int64_t b = t.rbeg - (t.qbeg + cal_max_gap(o, t.qbeg));
                int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(o, l_query - t.qbeg - t.len));
                rmax0 = rmax0 < b ? rmax0 : b;
                rmax1 = rmax1 > e ? rmax1 : e;
            }
            rmax0 = rmax0 > 0 ? rmax0 : 0;
            rmax1 = rmax1 < (l_pac << 1) ? rmax1 : (l_pac << 1);
            if (rmax0 < l_pac && l_pac < rmax1) {
                if (seeds[0].rbeg < l_pac) rmax1 = l_pac;
                else rmax0 = l_pac;
            }
            // ---- bns_fetch_seq: clamp to the row of seeds[0].rbeg, decode into scratch
            {
                int is_rev;
                int rid = bns_pos2rid(ix, bns_depos(ix, seeds[0].rbeg, &is_rev));
                int64_t far_beg = ix.ann_offset[rid], far_end = far_beg + ix.ann_len[rid];
                if (is_rev) { int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
                rmax0 = rmax0 > far_beg ? rmax0 : far_beg;
                rmax1 = rmax1 < far_end ? rmax1 : far_end;
                if (rmax1 < rmax0) rmax1 = rmax0;   // seed lying wholly in inter-row filler: empty fetch
            }
            const int64_t rlen = rmax1 - rmax0;
            if (rlen > (int64_t)P.rseq_cap) {
                if (!WARPMODE) return false;
                if (lane == 0) atomicMax(P.need_rseq, (uint32_t)(rlen < 0x7fffffff ? rlen : 0x7fffffff));
                continue;
            }
            // the reference window is decoded only when a ksw_extend2 call really runs here (most calls of a short-read batch
            // were answered ahead of time by the thread-per-extension pass, extend_plan.cu)
            bool rseq_ready = false;
            auto need_rseq = [&]() {
                if (rseq_ready || !WARPMODE) return;
                for (int64_t i = lane; i < rlen; i += 32) S.rseq[i] = (uint8_t)ref_base(ix, rmax0 + i);
                __syncwarp();
                rseq_ready = true;
            };
            // ---- seed order: by (score, index) ascending, processed from the top (keys are unique)
            if (lane == 0) {
                for (int i = 0; i < n; ++i) srt[i] = (uint64_t)(uint32_t)seeds[i].score << 32 | (uint64_t)i;
                ks_introsort_dev(n, srt, [](uint64_t a, uint64_t b) { return a < b; });
            }
            if (WARPMODE) __syncwarp();
            for (int k = n - 1; k >= 0; --k) {
                const SeedRec s = seeds[(uint32_t)srt[k]];
                int i;
                for (i = 0; i < n_reg; ++i) {   // was this seed already covered by an earlier region?
                    const RegRec& p = av[i];
                    int64_t rd; int qd, w, max_gap;
                    if (s.rbeg < p.rb || s.rbeg + s.len > p.re || s.qbeg < p.qb || s.qbeg + s.len > p.qe) continue;
                    if ((double)(s.len - p.seedlen0) > .1 * (double)l_query) continue;
                    qd = s.qbeg - p.qb; rd = s.rbeg - p.rb;
                    max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
                    w = max_gap < p.w ? max_gap : p.w;
                    if (qd - rd < w && rd - qd < w) break;
                    qd = p.qe - (s.qbeg + s.len); rd = p.re - (s.rbeg + s.len);
                    max_gap = cal_max_gap(o, qd < rd ? qd : (int)rd);
                    w = max_gap < p.w ? max_gap : p.w;
                    if (qd - rd < w && rd - qd < w) break;
                }
                if (i < n_reg) {
                    for (i = k + 1; i < n; ++i) {
                        if (srt[i] == 0) continue;
                        const SeedRec t = seeds[(uint32_t)srt[i]];
                        if ((double)t.len < (double)s.len * .95) continue;
                        if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                        if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
                    }
                    if (i == n) {
                        if (lane == 0) srt[k] = 0;
                        if (WARPMODE) __syncwarp();
                        continue;
                    }
                }
                RegRec a;
                a.rb = a.re = 0; a.hash = 0; a.qb = a.qe = 0; a.sub = a.csub = a.sub_n = a.seedcov = a.secondary = a.seedlen0 = a.n_comp = 0;
                a.frac_rep = 0.f;
                int aw0 = o.w, aw1 = o.w;
                a.w = o.w; a.score = a.truesc = -1; a.rid = c.rid;
                if (s.qbeg) {   // left extension: reversed query prefix vs reversed reference prefix
                    int64_t tmp = s.rbeg - rmax0;
                    int tl = tmp > 0 ? (int)tmp : 0;
                    ExtOut e;
                    for (i = 0; i < MAX_BAND_TRY; ++i) {
                        int prev = a.score;
                        aw0 = o.w << i;
                        bool hit = false;
                        if (i == 0 && P.memo && k == n - 1 && ci < (uint32_t)EXT_MEMO_CHAINS) {
                            const ExtMemo* M = P.memo + (size_t)r * EXT_MEMO_CHAINS + ci;
                            if (M->state == 2 && M->qlen == s.qbeg && M->tlen == tl && M->h0 == s.len * o.a && M->tpos == s.rbeg - 1) {
                                e.score = M->out[0]; e.qle = M->out[1]; e.tle = M->out[2]; e.gtle = M->out[3]; e.gscore = M->out[4]; e.max_off = M->out[5];
                                hit = true;
                            }
                        }
                        if (!hit) {
                            if constexpr (WARPMODE) {
                                need_rseq();
                                e = ksw_extend_warp_t<SMEM>(o, s.qbeg, query + s.qbeg - 1, -1, tl, S.rseq + tmp - 1, -1, aw0, o.pen_clip5, s.len * o.a,
                                                          S.ehh, smat, cells, rows);
                            } else return false;
                        }
                        ++calls;
                        a.score = e.score;
                        if (a.score == prev || e.max_off < (aw0 >> 1) + (aw0 >> 2)) break;
                    }
                    if (e.gscore <= 0 || e.gscore <= a.score - o.pen_clip5) { a.qb = s.qbeg - e.qle; a.rb = s.rbeg - e.tle; a.truesc = a.score; }
                    else { a.qb = 0; a.rb = s.rbeg - e.gtle; a.truesc = e.gscore; }
                } else { a.score = a.truesc = s.len * o.a; a.qb = 0; a.rb = s.rbeg; }
                if (s.qbeg + s.len != l_query) {   // right extension
                    const int sc0 = a.score, qe = s.qbeg + s.len;
                    const int64_t re = s.rbeg + s.len - rmax0;
                    int64_t tl64 = rlen - re;
                    int tl = tl64 > 0 ? (int)tl64 : 0;
                    ExtOut e;
                    for (i = 0; i < MAX_BAND_TRY; ++i) {
                        int prev = a.score;
                        aw1 = o.w << i;
                        bool hit = false;
                        if (i == 0 && P.memo && k == n - 1 && ci < (uint32_t)EXT_MEMO_CHAINS) {
                            const ExtMemo* M = P.memo + ((size_t)P.n_reads + r) * EXT_MEMO_CHAINS + ci;
                            if (M->state == 2 && M->qlen == l_query - qe && M->tlen == tl && M->h0 == sc0 && M->tpos == rmax0 + re) {
                                e.score = M->out[0]; e.qle = M->out[1]; e.tle = M->out[2]; e.gtle = M->out[3]; e.gscore = M->out[4]; e.max_off = M->out[5];
                                hit = true;
                            }
                        }
                        if (!hit) {
                            if constexpr (WARPMODE) {
                                need_rseq();
                                e = ksw_extend_warp_t<SMEM>(o, l_query - qe, query + qe, 1, tl, S.rseq + re, 1, aw1, o.pen_clip3, sc0, S.ehh, smat, cells, rows);
                            } else return false;
                        }
                        ++calls;
                        a.score = e.score;
                        if (a.score == prev || e.max_off < (aw1 >> 1) + (aw1 >> 2)) break;
                    }
                    if (e.gscore <= 0 || e.gscore <= a.score - o.pen_clip3) { a.qe = qe + e.qle; a.re = rmax0 + re + e.tle; a.truesc += a.score - sc0; }
                    else { a.qe = l_query; a.re = rmax0 + re + e.gtle; a.truesc += e.gscore - sc0; }
                } else { a.qe = l_query; a.re = s.rbeg + s.len; }
                a.seedcov = 0;
                for (i = 0; i < n; ++i) {
                    const SeedRec t = seeds[i];
                    if (t.qbeg >= a.qb && t.qbeg + t.len <= a.qe && t.rbeg >= a.rb && t.rbeg + t.len <= a.re) a.seedcov += t.len;
                }
                a.w = aw0 > aw1 ? aw0 : aw1;
                a.seedlen0 = s.len;
                a.frac_rep = c.frac_rep;
                if (lane == 0) av[n_reg] = a;   // n_reg < n_seeds of the read <= n_alloc
                ++n_reg;
                if (WARPMODE) __syncwarp();
            }
        }
        if (lane == 0) P.reg_cnt[r] = (uint32_t)n_reg;
    }
    return true;
}

template <bool SMEM>
__global__ void __launch_bounds__(EXT_THREADS, SMEM ? 8 : 4) sw_extend(ExtendParams P, DevIndex ix, DevOpts o) {
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    __shared__ int smat[25];
    if (threadIdx.x < 25) smat[threadIdx.x] = o.mat[threadIdx.x];
    __syncthreads();
    const int lane = lane_id();
    const uint32_t gwarp = (blockIdx.x * EXT_THREADS + threadIdx.x) >> 5;
    uint8_t* sbase = SMEM ? dyn_smem + (size_t)(threadIdx.x >> 5) * P.scratch_per_warp : P.scratch + (size_t)gwarp * P.scratch_per_warp;
    const Scratch S = carve(sbase, P.max_len, P.rseq_cap);
    unsigned long long cells = 0, calls = 0, rows = 0;
    // with the pre-pass on, only the reads ext_finish could not complete are left (P.todo); otherwise every read
    const uint32_t n_todo = P.todo ? *P.todo_cnt : P.n_reads;
    for (;;) {
        const uint32_t t = next_ticket(P.ticket);
        if (t >= n_todo) break;
        extend_read<true, SMEM>(P, ix, o, P.todo ? P.todo[t] : t, S, smat, cells, calls, rows);
    }
    if (P.counters && lane == 0) { atomicAdd(&P.counters[0], cells); atomicAdd(&P.counters[1], calls); atomicAdd(&P.counters[2], rows); }
}

// Thread per read: completes every read whose ksw_extend2 calls were all answered by the pre-pass; the others go to P.todo.
__global__ void __launch_bounds__(128) ext_finish(ExtendParams P, DevIndex ix, DevOpts o) {
    const uint32_t r = blockIdx.x * 128 + threadIdx.x;
    unsigned long long cells = 0, calls = 0, rows = 0;
    if (r < P.n_reads) {
        Scratch S; S.rseq = nullptr; S.query = nullptr; S.ehh = nullptr; S.ehe = nullptr;
        if (!extend_read<false, false>(P, ix, o, r, S, nullptr, cells, calls, rows)) {
            calls = 0;
            P.todo[atomicAdd(P.todo_cnt, 1u)] = r;
        }
    }
    if (P.counters) {
#pragma unroll
        for (int d = 16; d; d >>= 1) calls += __shfl_xor_sync(FULL, calls, d);
        if (lane_id() == 0 && calls) atomicAdd(&P.counters[1], calls);
    }
}

}  // namespace

size_t extend_scratch_per_warp(uint32_t max_len, uint32_t rseq_cap) {
    size_t b = (size_t)(max_len + 2) * 8 + rseq_cap + max_len;
    return (b + 15) & ~(size_t)15;
}

static bool ext_use_smem(size_t per_warp) { return per_warp * EXT_WARPS <= 40 * 1024; }

int extend_resident_warps() {
    return cached_blocks_per_sm(sw_extend<false>, EXT_THREADS, 0) * cached_sm_count() * EXT_WARPS;
}

void launch_extend_finish(const ExtendParams& p, const DevIndex& ix, const DevOpts& o, cudaStream_t st) {
    if (!p.todo || p.n_reads == 0) return;
    ext_finish<<<(p.n_reads + 127) / 128, 128, 0, st>>>(p, ix, o);
}

void launch_extend(const ExtendParams& p, const DevIndex& ix, const DevOpts& o, cudaStream_t st) {
    const int sms = cached_sm_count();
    if (ext_use_smem(p.scratch_per_warp)) {
        const size_t smem = p.scratch_per_warp * EXT_WARPS;
        const int nb = cached_blocks_per_sm(sw_extend<true>, EXT_THREADS, smem);
        sw_extend<true><<<nb * sms, EXT_THREADS, smem, st>>>(p, ix, o);
    } else {
        const int nb = cached_blocks_per_sm(sw_extend<false>, EXT_THREADS, 0);
        sw_extend<false><<<nb * sms, EXT_THREADS, 0, st>>>(p, ix, o);
    }
}
